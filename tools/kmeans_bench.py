"""KMeans anchors of BASELINE config C3 (Flickr-shape node2vec table, 256 centres): device vs scikit-learn."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from graphpope_b200 import device as dev, synth
n, d, k = 89250, 128, 256
table = synth.node2vec_table(n, d, 3)
t_d = torch.as_tensor(table).cuda()
for n_init in (1, 10):
    torch.cuda.synchronize(); t = time.perf_counter()
    c, inertia, iters = dev.kmeans(t_d, k, n_init=n_init, seed=0)
    torch.cuda.synchronize(); print(f"device KMeans n_init={n_init}: {(time.perf_counter() - t) * 1e3:.1f} ms, inertia {inertia:.6e}, {iters} iterations (best run)", flush=True)
if len(sys.argv) > 1 and sys.argv[1] == "sklearn":
    from sklearn.cluster import KMeans
    t = time.perf_counter(); ref = KMeans(n_clusters=k, n_init=1, random_state=0).fit(table)
    print(f"scikit-learn KMeans n_init=1: {(time.perf_counter() - t):.1f} s, inertia {ref.inertia_:.6e}, {ref.n_iter_} iterations")
