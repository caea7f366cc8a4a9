"""Driver for an ncu capture of one gp_betweenness launch on a BASELINE-shaped graph:
    GP_BC_GROUP=2 ncu --set full --import-source on -k regex:bc_kernel -c 1 -o out python tools/bc_profile.py"""
import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from graphpope_b200 import device as dev, synth
shape = synth.SHAPES[sys.argv[1] if len(sys.argv) > 1 else "flickr-shape"]
ei = synth.make_graph(shape)
csr = dev.DeviceCsr(shape.num_nodes, ei.shape[1]).build(torch.as_tensor(ei).cuda())
csr.info()
s = csr.betweenness()
torch.cuda.synchronize()
print("done", float(s.sum()))
