"""BASELINE config C5 at full size on ONE GPU's share: ogbn-products-shaped graph (2.45 M nodes, 123.7 M
directed edges), 512 anchors (= 4096 / 8 ranks).  Size-independent checks + a few columns against the C
oracle, and timings of the three stages."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from graphpope_b200 import device as dev, synth

name = sys.argv[1] if len(sys.argv) > 1 else "products-shape"
k = int(sys.argv[2]) if len(sys.argv) > 2 else 512
sh = synth.SHAPES[name]; n = sh.num_nodes
t = time.perf_counter(); ei = synth.make_graph(sh); print(f"graph {name}: N={n} E={ei.shape[1]} generated in {time.perf_counter()-t:.1f} s", flush=True)
anchors = synth.stochastic_anchors(n, 4096, 42)[:k]
ei_d = torch.as_tensor(ei).cuda(); a_d = torch.as_tensor(anchors).cuda()
eng = dev.GeodesicEngine(n, ei.shape[1], k)
ev = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
for rep in range(3):
    ev[0].record(); eng.csr.build(ei_d); ev[1].record(); eng.bfs.run(a_d); ev[2].record()
    hops = eng.bfs.hops_u16(); ev[3].record(); torch.cuda.synchronize()
    st = eng.bfs.stats()
    print(f"rep {rep}: csr build {ev[0].elapsed_time(ev[1]):.2f} ms, ms-bfs {ev[1].elapsed_time(ev[2]):.2f} ms (kernel {eng.bfs.kernel_ms():.2f}), "
          f"hops decode {ev[2].elapsed_time(ev[3]):.2f} ms; {k * ei.shape[1] / ev[1].elapsed_time(ev[2]) / 1e6:.1f} GTEPS (BFS only); {st}", flush=True)
info = eng.csr.info(); print("csr info", info, flush=True)
assert info["num_edges"] == ei.shape[1] and info["is_symmetric"] == 1
h = hops.view(torch.int16)  # [N, k] on device; torch cannot index uint16, so reinterpret and mask below
def as_i32(t):
    return t.to(torch.int32) & 0xFFFF
# (1) every anchor is at distance 0 from itself
assert bool((h[a_d, torch.arange(k, device="cuda")] == 0).all())
# (2) along every edge (u, v) of a symmetric graph hop counts differ by at most 1 (unreachable = 0xFFFF on both ends)
src, dst = ei_d[0], ei_d[1]
bad = 0
for c0 in range(0, k, 64):
    hu = as_i32(h[:, c0:c0 + 64])
    for e0 in range(0, ei.shape[1], 16_000_000):
        s_, d_ = src[e0:e0 + 16_000_000], dst[e0:e0 + 16_000_000]
        bad += int(((hu[s_] - hu[d_]).abs() > 1).sum().item())
print("edges violating |hops(u) - hops(v)| <= 1:", bad); assert bad == 0
# (3) every reached non-anchor (node, lane) has a neighbour one hop closer: checked through a scatter-min over edges
for c0 in range(0, k, 64):
    hu = as_i32(h[:, c0:c0 + 64])
    best = torch.full_like(hu, 1 << 20)
    for e0 in range(0, ei.shape[1], 16_000_000):
        s_, d_ = src[e0:e0 + 16_000_000], dst[e0:e0 + 16_000_000]
        best.scatter_reduce_(0, s_.unsqueeze(1).expand(-1, hu.size(1)), hu[d_], reduce="amin")
    reach = (hu > 0) & (hu < 0xFFFF)
    assert bool((best[reach] == hu[reach] - 1).all()), "a reached node without a parent one hop closer"
print("parent property holds for all reached (node, anchor) pairs")
# (4) a few columns against the CPU oracle (C single-source BFS)
from oracle import cbfs
cols = [0, k // 2, k - 1]
want = cbfs.bfs_hops(cbfs.InCsr(ei, n), anchors[cols])
got = h[:, cols].cpu().numpy().view(np.uint16)
assert np.array_equal(got, want), "hops differ from the oracle"
print("columns", cols, "bit-equal to the oracle; max hop", int(want[want != 0xFFFF].max()))
print("PRODUCTS CHECK PASSED")
