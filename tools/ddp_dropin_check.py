"""torchrun check of the drop-in under DDP: every rank calls utils.Graphpope with GRAPHPOPE_SHARED=1; all ranks must
return the SAME node-shared matrix, bit-equal to the oracle (rank 0 checks), for a stochastic and a device sampler."""
import os, sys
os.environ["GRAPHPOPE_SHARED"] = "1"; os.environ["GRAPHPOPE_QUIET"] = "1"
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch, torch.distributed as dist
from graphpope_b200 import synth, utils
rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
ok = True
for name, k, method in (("pubmed-shape", 256, "stochastic"), ("flickr-shape", 100, "degree_centrality"), ("pubmed-shape", 3, "stochastic")):
    sh = synth.SHAPES[name]; n = sh.num_nodes
    ei = synth.make_graph(sh)
    x = torch.randn(n, 20, generator=torch.Generator().manual_seed(5))
    class D: pass
    d = D(); d.num_nodes, d.edge_index, d.x = n, torch.as_tensor(ei), x
    np.random.seed(42); utils.clear_cache()
    out = utils.Graphpope(d, name, "geodesic", method, k, None, num_workers=6)
    good = (not out.is_cuda) and tuple(out.shape) == (n, 20 + k) and out.dtype == torch.float32
    # one matrix per node: a write by rank 0 is visible to every rank
    dist.barrier()
    probe = float(out[0, 0].item())
    if rank == 0: out[0, 0] = 12345.0
    dist.barrier()
    shared_ok = float(out[0, 0].item()) == 12345.0
    dist.barrier()
    if rank == 0: out[0, 0] = probe
    dist.barrier()
    if rank == 0:
        from oracle import cbfs, geodesic
        want = geodesic.concat_features(x.numpy(), cbfs.geodesic_features(ei, n, np.asarray(d.anchor_nodes)))
        same = bool(np.array_equal(out.numpy().view(np.uint32), want.view(np.uint32)))
        print(f"[{name} K={k} {method} G={world}] shape/dtype {good}; one shared matrix {shared_ok}; == oracle {same}", flush=True)
        good &= same
    ok &= good and shared_ok
t = torch.tensor([int(ok)], device="cuda"); dist.all_reduce(t, op=dist.ReduceOp.MIN)
if rank == 0: print("DDP DROP-IN CHECK", "PASSED" if int(t.item()) else "FAILED", flush=True)
dist.destroy_process_group()
sys.exit(0 if int(t.item()) else 1)
