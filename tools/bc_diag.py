"""Diagnostics for gp_betweenness: phase and per-launch times (GP_BC_TRACE) over repeated calls."""
import os, sys, time
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from graphpope_b200 import device as dev, synth
os.environ["GP_BC_TRACE"] = "1"
name = "pubmed-shape"
n, ei = synth.SHAPES[name].num_nodes, synth.make_graph(synth.SHAPES[name])
csr = dev.DeviceCsr(n, ei.shape[1]).build(torch.as_tensor(ei).cuda())
csr.info()
for rep in range(5):
    torch.cuda.synchronize(); t = time.perf_counter()
    s = csr.betweenness(); torch.cuda.synchronize()
    print(f"{name} n={n} rep={rep}: {1e3 * (time.perf_counter() - t):.1f} ms", file=sys.stderr, flush=True)
