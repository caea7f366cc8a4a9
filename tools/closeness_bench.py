"""Closeness-centrality scores of every node of the Flickr-shaped graph (utils.py:50-54) on the device."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from graphpope_b200 import device as dev, synth
name = sys.argv[1] if len(sys.argv) > 1 else "flickr-shape"
sh = synth.SHAPES[name]; n = sh.num_nodes
ei = synth.make_graph(sh); ei_d = torch.as_tensor(ei).cuda()
csr = dev.DeviceCsr(n, ei.shape[1]).build(ei_d)
for rep in range(3):
    torch.cuda.synchronize(); t = time.perf_counter(); x = csr.closeness(); torch.cuda.synchronize()
    ms = (time.perf_counter() - t) * 1e3
    print(f"{name}: closeness of {n} nodes in {ms:.1f} ms = {n * ei.shape[1] / ms / 1e6:.1f} GTEPS (N*|E|/t)", flush=True)
idx = dev.topk_stable(x, 256).cpu().numpy()
# spot-check 8 nodes against networkx-equivalent BFS sums from the C oracle
from oracle import cbfs
sample = np.concatenate([idx[-4:], np.array([0, 1, n // 2, n - 1])])
hops = cbfs.bfs_hops(cbfs.InCsr(ei, n), sample).astype(np.int64)
for j, a in enumerate(sample):
    reach = hops[:, j] != 0xFFFF; r = int(reach.sum()); tot = int(hops[reach, j].sum())
    want = 0.0 if tot == 0 or n == 1 else ((r - 1.0) / tot) * ((r - 1.0) / (n - 1))
    assert float(x[a].item()) == want, (a, float(x[a].item()), want)
print("8 sampled scores bit-equal to the oracle; top anchor", int(idx[-1]), "score", float(x[idx[-1]].item()))
