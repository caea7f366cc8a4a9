"""Time the device betweenness sampler (gp_betweenness) on a BASELINE-shaped graph and spot-check a few
scores against networkx run on the same box (single-source Brandes restated for a handful of sources is
not a check of the full sum, so the check is on a 3000-node subsample graph instead).

    python tools/betweenness_bench.py [--workload flickr-shape|pubmed-shape] [--check-n 3000]
"""
import argparse
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from graphpope_b200 import device as dev  # noqa: E402
from graphpope_b200 import synth  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--workload", default="flickr-shape")
    ap.add_argument("--check-n", type=int, default=3000)
    args = ap.parse_args()
    shape = synth.SHAPES[args.workload]
    ei = synth.make_graph(shape)
    n = shape.num_nodes
    ei_d = torch.as_tensor(ei).cuda()
    csr = dev.DeviceCsr(n, ei.shape[1]).build(ei_d)
    e = csr.info()["num_edges"]
    torch.cuda.synchronize()
    for spl in (1, 2, 4):
        os.environ["GP_BC_SOURCES"] = str(spl)
        for rep in range(2):
            t = time.perf_counter()
            score = csr.betweenness()
            torch.cuda.synchronize()
            dt = time.perf_counter() - t
            print(f"{args.workload}: N={n} E'={e} {32 * spl} sources per batch: betweenness of all nodes in {dt:.3f} s "
                  f"= {n * e / dt / 1e9:.1f} G source-edges/s ({2 * n * e / dt / 1e9:.1f} G edge visits/s, both sweeps)",
                  flush=True)
    del os.environ["GP_BC_SOURCES"]
    s = score.cpu().numpy()
    print("sum", float(s.sum()), "max", float(s.max()), "argmax", int(s.argmax()), "non-zero", int((s > 0).sum()))

    # parity at a size networkx finishes in about a minute
    import networkx as nx
    m = args.check_n
    ei2 = synth.chung_lu_symmetric(m, 6 * m, 2.2, seed=9)
    G = nx.DiGraph()
    G.add_nodes_from(range(m))
    G.add_edges_from(zip(ei2[0].tolist(), ei2[1].tolist()))
    t = time.perf_counter()
    want = nx.betweenness_centrality(G)
    t_nx = time.perf_counter() - t
    want = np.asarray([want[i] for i in range(m)])
    csr2 = dev.DeviceCsr(m, ei2.shape[1]).build(torch.as_tensor(ei2).cuda())
    t = time.perf_counter()
    got = csr2.betweenness().cpu().numpy()
    t_dev = time.perf_counter() - t
    rel = np.max(np.abs(got - want) / np.maximum(want, 1e-300))
    print(f"check graph N={m} E={ei2.shape[1]}: networkx {t_nx:.1f} s, device {t_dev * 1e3:.1f} ms, "
          f"max relative difference {rel:.2e}, zeros agree {np.array_equal(got == 0, want == 0)}")


if __name__ == "__main__":
    main()
