"""Where the time of the host-buffer entry point (gp_geodesic_embed_host) goes on this box."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from graphpope_b200 import device as dev, synth, _lib
from graphpope_b200._lib import check
from graphpope_b200.device import _ptr

sh = synth.SHAPES["flickr-shape"]; n, f, k = sh.num_nodes, sh.num_features, 256
ei = synth.make_graph(sh); anchors = synth.stochastic_anchors(n, k, 42)
ei_h = torch.as_tensor(ei).pin_memory(); x_h = torch.randn(n, f).pin_memory()
out_h = torch.empty(n, f + k).pin_memory(); blk_h = torch.empty(n, k).pin_memory()
lib = _lib.load()

def timeit(fn, reps=10, warm=3):
    for _ in range(warm): fn()
    torch.cuda.synchronize(); t = time.perf_counter()
    for _ in range(reps): fn()
    torch.cuda.synchronize(); return (time.perf_counter() - t) / reps * 1e3

print("full call, x concatenated        %.2f ms" % timeit(lambda: dev.geodesic_embed_host(ei_h, n, anchors, x_h, out=out_h)))
print("full call, no x (block only)     %.2f ms" % timeit(lambda: dev.geodesic_embed_host(ei_h, n, anchors, None, out=blk_h)))
import ctypes
st = _lib.MsbfsStats(); a_h = torch.as_tensor(np.asarray(anchors, dtype=np.int64))
print("full call, no x, strided out     %.2f ms" % timeit(lambda: check(lib.gp_geodesic_embed_host(_ptr(ei_h), ei_h.size(1), n, 0, _ptr(a_h), k, None, 0, _ptr(out_h), f + k, f, None, ctypes.byref(st)))))
print("host concat of x alone           %.2f ms" % timeit(lambda: check(lib.gp_host_concat(_ptr(x_h), f, None, 0, n, _ptr(out_h), f + k))))
ei_d = torch.empty_like(ei_h, device="cuda"); blk_d = torch.empty(n, k, device="cuda")
print("H2D edge_index (14.4 MB)         %.2f ms" % timeit(lambda: ei_d.copy_(ei_h, non_blocking=True)))
print("D2H block contiguous (91 MB)     %.2f ms" % timeit(lambda: blk_h.copy_(blk_d, non_blocking=True)))
print("D2H block into [N,F+K] (2-D)     %.2f ms" % timeit(lambda: out_h[:, f:].copy_(blk_d, non_blocking=True)))
a_d = torch.as_tensor(anchors).cuda(); eng = dev.GeodesicEngine(n, ei.shape[1], k)
print("device pipeline (no x)           %.2f ms" % timeit(lambda: eng.run(ei_d, a_d, None, blk_d)))
t = torch.empty(n, f + k); src = torch.randn(n, f)
print("pageable torch copy x -> out     %.2f ms" % timeit(lambda: t[:, :f].copy_(src)))
print("cpu threads", os.cpu_count(), torch.get_num_threads())
