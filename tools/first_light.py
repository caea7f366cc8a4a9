"""Staged GPU diagnostics: CSR -> MS-BFS -> epilogue vs the oracle, with coarse timings."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from graphpope_b200 import device as dev, synth
from oracle import cbfs, geodesic as g

def stage(name, fn):
    t = time.time(); r = fn(); torch.cuda.synchronize(); print("[%-28s] %.3f s" % (name, time.time() - t), flush=True); return r

print(dev.device_info(), flush=True)
for shape_name, K in (("tiny", 8), ("pubmed-shape", 256), ("flickr-shape", 256), ("flickr-shape", 1024)):
    if shape_name == "tiny":
        n = 300; ei = synth.random_digraph(n, 900, seed=7); f = 5
    else:
        sh = synth.SHAPES[shape_name]; n = sh.num_nodes; ei = synth.make_graph(sh); f = sh.num_features
    anchors = synth.stochastic_anchors(n, K, 42)
    print("==", shape_name, "N", n, "E", ei.shape[1], "K", K, flush=True)
    ei_d = torch.as_tensor(ei).cuda(); a_d = torch.as_tensor(anchors).cuda()
    x_d = torch.randn(n, f, device="cuda")
    csr = stage("csr create", lambda: dev.DeviceCsr(n, ei.shape[1]))
    stage("csr build", lambda: csr.build(ei_d))
    info = stage("csr info", lambda: csr.info()); print(info, flush=True)
    rp, col = csr.export("out")
    rp_w, col_w = g.out_csr(ei, n)
    print("csr out ok:", np.array_equal(rp.cpu().numpy(), rp_w), np.array_equal(col.cpu().numpy(), col_w), flush=True)
    bfs = stage("bfs create", lambda: dev.MsBfs(csr, K))
    stage("bfs run", lambda: bfs.run(a_d))
    st = stage("bfs stats", lambda: bfs.stats()); print(st, flush=True)
    hops = stage("hops u16", lambda: bfs.hops_u16()).cpu().numpy()
    want = cbfs.bfs_hops(cbfs.InCsr(ei, n), anchors)
    print("hops ok:", np.array_equal(hops, want), "mismatches", int((hops != want).sum()), flush=True)
    feats = stage("features", lambda: bfs.features(x_d)).cpu().numpy()
    print("features ok:", np.array_equal(feats[:, f:].view(np.uint32), cbfs.normalise(want).view(np.uint32)),
          "x ok:", np.array_equal(feats[:, :f], x_d.cpu().numpy()), flush=True)
    # timing: whole pipeline, device-resident inputs
    out = torch.empty(n, f + K, device="cuda")
    for _ in range(3):
        csr.build(ei_d); bfs.run(a_d); bfs.features(x_d, out)
    torch.cuda.synchronize()
    evs = [torch.cuda.Event(enable_timing=True) for _ in range(5)]
    reps = 20
    tt = np.zeros(4)
    for _ in range(reps):
        evs[0].record(); csr.build(ei_d); evs[1].record(); bfs.run(a_d); evs[2].record(); bfs.features(None, out[:, f:]) if False else bfs.features(x_d, out); evs[3].record()
        torch.cuda.synchronize()
        tt[:3] += [evs[i].elapsed_time(evs[i + 1]) for i in range(3)]
    tt /= reps
    e_u = info["num_edges"]
    print("avg ms: csr %.4f  bfs %.4f  epilogue+concat %.4f  total %.4f  -> %.1f GTEPS" % (tt[0], tt[1], tt[2], tt[:3].sum(), K * e_u / (tt[:3].sum() * 1e-3) / 1e9), flush=True)
    host_out, _, st2 = dev.geodesic_embed_host(ei, n, anchors, x_d.cpu().numpy())
    print("host entry ok:", np.array_equal(host_out.numpy()[:, f:].view(np.uint32), cbfs.normalise(want).view(np.uint32)), flush=True)
    t = time.time(); dev.geodesic_embed_host(ei, n, anchors, x_d.cpu().numpy()); print("host entry wall %.2f ms" % ((time.time() - t) * 1e3), flush=True)
print("FIRST LIGHT DONE", flush=True)
