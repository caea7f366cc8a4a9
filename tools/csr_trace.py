"""In-graph stage times of the csr build (GP_CSR_TRACE=1): events recorded after every launch of the replayed step."""
import os, sys, ctypes
os.environ["GP_CSR_TRACE"] = "1"
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from graphpope_b200 import device as dev, synth, _lib
lib = _lib.load()
sh = synth.SHAPES[sys.argv[1] if len(sys.argv) > 1 else "flickr-shape"]; n, f, k = sh.num_nodes, sh.num_features, 256
ei = synth.make_graph(sh); anchors = synth.stochastic_anchors(n, k, 42)
ei_d = torch.as_tensor(ei).cuda(); a_d = torch.as_tensor(anchors).cuda()
x = torch.randn(n, f, device="cuda"); out = torch.empty(n, f + k, device="cuda")
flush = torch.empty(64 * 1024 * 1024, device="cuda")
eng = dev.GeodesicEngine(n, ei.shape[1], k); eng.bfs.set_stage_events(True)
names = ["memsets", "count", "scan1", "scatter", "rowsort", "scan9", "desc"]
rows = []
for i in range(14):
    flush.fill_(float(i)); s = flush[:1024].sum()
    eng.run(ei_d, a_d, x, out); torch.cuda.synchronize()
    ms = (ctypes.c_float * 12)(); num = ctypes.c_int32()
    _lib.check(lib.gp_csr_trace_ms(eng.csr._h, ms, 12, ctypes.byref(num)))
    if i >= 4: rows.append([ms[j] * 1e3 for j in range(num.value)])
med = np.median(np.array(rows), axis=0)
print("csr build stages inside the replayed step (us, median of 10, L2 flushed):", dict(zip(names, np.round(med, 1))), "sum", round(float(med.sum()), 1))
print("pipeline stages (us):", [round(v * 1e3, 1) for v in eng.bfs.pipeline_stage_ms()])
