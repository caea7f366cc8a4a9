"""Diagnostics for gp_betweenness: per-launch sweep counts and times (GP_BC_TRACE) on small and large graphs."""
import os, sys, time
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from graphpope_b200 import device as dev, synth
os.environ["GP_BC_TRACE"] = "1"
cases = {"cl3000": (3000, synth.chung_lu_symmetric(3000, 18000, 2.2, seed=9)),
         "pubmed": (synth.SHAPES["pubmed-shape"].num_nodes, synth.make_graph(synth.SHAPES["pubmed-shape"]))}
for name, (n, ei) in cases.items():
    csr = dev.DeviceCsr(n, ei.shape[1]).build(torch.as_tensor(ei).cuda())
    csr.info()
    for spl in (1, 2, 4):
        os.environ["GP_BC_SOURCES"] = str(spl)
        os.environ["GP_BC_GROUP"] = "16"
        for rep in range(2):
            torch.cuda.synchronize(); t = time.perf_counter()
            s = csr.betweenness(); torch.cuda.synchronize()
            print(f"{name} n={n} spl={spl} rep={rep}: {1e3 * (time.perf_counter() - t):.1f} ms", flush=True)
