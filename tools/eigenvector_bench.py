"""Time the device eigenvector-centrality sampler (gp_eigenvector) on a BASELINE-shaped graph (its giant
strongly connected part is what matters; the kernel itself does not need connectivity)."""
import os, sys, time
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from graphpope_b200 import device as dev, synth
for name in ("pubmed-shape", "flickr-shape"):
    shape = synth.SHAPES[name]
    ei = synth.make_graph(shape)
    csr = dev.DeviceCsr(shape.num_nodes, ei.shape[1]).build(torch.as_tensor(ei).cuda())
    csr.info()
    for rep in range(2):
        torch.cuda.synchronize(); t = time.perf_counter()
        x, it = csr.eigenvector(); torch.cuda.synchronize()
        dt = time.perf_counter() - t
    x = x.cpu().numpy()
    print(f"{name}: N={shape.num_nodes} eigenvector centrality in {dt * 1e3:.2f} ms, {it} iterations "
          f"({dt * 1e6 / max(it, 1):.1f} us per iteration), |x|={np.linalg.norm(x):.15f}, argmax {int(x.argmax())}")
