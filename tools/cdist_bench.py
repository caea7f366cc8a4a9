"""Timing of the node2vec block (BASELINE config C3): tcgen05 GEMM + min-max on Flickr-shape sizes."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from graphpope_b200 import device as dev, synth
n, k, d = 89250, 256, 128
emb = torch.as_tensor(synth.node2vec_table(n, d, 3)).cuda()
anc = emb[torch.as_tensor(synth.stochastic_anchors(n, k, 42)).cuda()].contiguous()
out = torch.empty(n, k, device="cuda")
for mode in ("euclidean", "distance", "similarity"):
    for mm in (False, True):
        for _ in range(3): dev.cdist_minmax(emb, anc, mode, mm, out=out)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(20): dev.cdist_minmax(emb, anc, mode, mm, out=out)
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 20
        byt = 4 * (n * d + k * d + n * k)
        print(f"{mode:10s} minmax={mm!s:5s} {ms*1e3:8.1f} us  gemm-bytes {byt/1e6:.1f} MB -> {byt/ms/1e6:.0f} GB/s (pairwise only), {2*n*k*d*3/ms/1e9:.1f} TFLOP/s bf16x3", flush=True)
