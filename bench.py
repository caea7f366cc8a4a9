#!/usr/bin/env python
"""bench.py — GraphPOPE geodesic embedding generation on B200 (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload flickr-shape]

A *step* is one full pass of the hot path over one synthetic graph already resident in HBM:
device CSR build (dedup) -> multi-source BFS (persistent kernel) -> fused normalise + concat
epilogue writing the float32 [N, F+K] matrix.  Metric: anchor-BFS GTEPS = K*|E'|/t with |E'| the
de-duplicated directed edge count.  At N>1 the anchors are sharded (256 per GPU, weak scaling) and
the bit-sliced results all-gathered so every rank holds the full [N, F + 256*N] matrix.

Prints ONE JSON line (rank 0).  `--impl reference` times the reference algorithm (utils.py:64-114:
N*K networkx shortest_path calls in a process pool) restated in oracle/ on a bounded row sample.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import numpy as np

METRIC = "anchor-BFS GTEPS (K*|E|/s), geodesic GraphPOPE embedding, Flickr-shape K=256 per GPU"
UNIT = "GTEPS"
K_PER_GPU = 256


# --------------------------------------------------------------------------- CPU reference legs
def _t0_worker(args):
    """One pool job of utils.py:99-100: shortest_path_length over a node slice."""
    ei, n, anchors, rows = args
    from oracle import geodesic
    G = geodesic.to_digraph(ei, n)
    t = time.perf_counter()
    geodesic.t0_rows(G, anchors, rows)
    return time.perf_counter() - t, len(rows)


def time_reference_sample(ei, n, anchors, num_workers, budget_s, probe_rows=3):
    """Time the reference algorithm (oracle T0 = utils.py:64-81 restated) on a bounded row sample.

    Rows are independent, so the full-graph time is extrapolated linearly (stated in `sample`).
    Returns (extrapolated_seconds_full, seconds_measured, rows_timed).
    """
    import multiprocessing as mp
    from oracle import geodesic

    anchors = [int(a) for a in anchors]
    G = geodesic.to_digraph(ei, n)
    rng = np.random.default_rng(0)
    probe = rng.integers(0, n, probe_rows).tolist()
    t = time.perf_counter()
    geodesic.t0_rows(G, anchors, probe)
    per_row = max((time.perf_counter() - t) / probe_rows, 1e-6)
    del G
    rows_total = int(max(num_workers, min(n, budget_s * num_workers / per_row)))
    stride = max(1, n // rows_total)
    rows = np.arange(0, n, stride)[:rows_total]
    chunks = [c.tolist() for c in np.array_split(rows, num_workers) if len(c)]
    ctx = mp.get_context("fork")
    t = time.perf_counter()
    with ctx.Pool(processes=len(chunks)) as pool:
        res = pool.map(_t0_worker, [(ei, n, anchors, c) for c in chunks])
    wall = time.perf_counter() - t
    # the pool's wall time includes building G in every worker, as the reference pickles G per job
    busy = max(r[0] for r in res)
    timed_rows = sum(r[1] for r in res)
    full = busy * (n / max(1, max(r[1] for r in res) * len(chunks)))
    return full, wall, timed_rows


def run_reference_arm(args, shape, ei, anchors, e_unique):
    cores = min(os.cpu_count() or 1, 32)
    total_budget = 150.0
    per_step = max(2.0, total_budget / max(1, args.steps + args.warmup))
    vals = []
    rows = 0
    for i in range(args.warmup + args.steps):
        full, wall, rows = time_reference_sample(ei, shape.num_nodes, anchors, cores, per_step * 0.6)
        if i >= args.warmup:
            vals.append(full)
    full = float(np.mean(vals))
    value = K_PER_GPU * e_unique / full / 1e9
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": full * 1e3, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "u64", "data": "synthetic",
        "config": {"workload": workload_name(shape, K_PER_GPU), "num_workers": cores},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port",
                         "sample": f"oracle T0 (utils.py:64-81 restated: N*K nx.shortest_path, mp pool of {cores}); "
                                   f"{rows} of {shape.num_nodes} rows timed per step, extrapolated linearly"},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), flush=True)


def workload_name(shape, k):
    return (f"{shape.name} synthetic graph ({shape.num_nodes} nodes, {shape.num_directed_edges} directed edges), "
            f"geodesic, stochastic sampling (seed 42), {k} anchors per GPU, F={shape.num_features}")


# --------------------------------------------------------------------------- clocks
class ClockSampler:
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index=0):
        self.samples, self.proc, self.index = [], None, index

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.samples.append(line.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for s in self.samples:
            p = [x.strip() for x in s.split(",")]
            if len(p) < 7:
                continue
            try:
                sm.append(float(p[0])); mx.append(float(p[1]))
            except ValueError:
                continue
            for name, v in zip(names, p[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


# --------------------------------------------------------------------------- main
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="flickr-shape")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))

    from graphpope_b200 import synth
    shape = synth.SHAPES[args.workload]

    if args.impl == "reference":
        if rank != 0:
            return
        ei = synth.make_graph(shape)
        anchors = synth.stochastic_anchors(shape.num_nodes, K_PER_GPU, 42)
        from oracle import geodesic
        e_unique = geodesic.dedup_edges(ei, shape.num_nodes)[0].size
        run_reference_arm(args, shape, ei, anchors, e_unique)
        return

    n, f = shape.num_nodes, shape.num_features
    k_total = K_PER_GPU * world
    ei = synth.make_graph(shape)
    anchors = synth.stochastic_anchors(n, k_total, 42)  # identical on every rank (same seed)

    # ---- CPU baseline first (before CUDA is initialised: the pool forks), rank 0 at N=1 only
    cpu_baseline = None
    if world == 1 and rank == 0 and not args.no_cpu_baseline:
        from oracle import cbfs, geodesic
        e_unique_cpu = geodesic.dedup_edges(ei, n)[0].size
        workers = 6  # reference default, main.py:39
        full, wall, rows = time_reference_sample(ei, n, anchors[:K_PER_GPU], workers, budget_s=15.0)
        t = time.perf_counter()
        cbfs.bfs_hops(cbfs.InCsr(ei, n), anchors[:K_PER_GPU])
        fair = time.perf_counter() - t
        cpu_baseline = {
            "value": K_PER_GPU * e_unique_cpu / full / 1e9, "unit": UNIT, "cores": workers, "kind": "port",
            "sample": f"oracle T0 (utils.py:64-81 restated: N*K nx.shortest_path calls, mp pool of {workers} = reference "
                      f"default num_workers); {rows} of {n} rows timed in {wall:.1f} s, extrapolated linearly to "
                      f"{full:.0f} s for the full graph",
            "fair_cpu_gteps": K_PER_GPU * e_unique_cpu / fair / 1e9,
            "fair_cpu_note": f"oracle C tier (K single-source BFS, 1 thread, full workload) {fair:.2f} s",
            "host_cpus": os.cpu_count(),
        }

    import torch
    import torch.distributed as dist
    from graphpope_b200 import device as dev
    from graphpope_b200 import distributed as gpd

    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    ei_d = torch.as_tensor(ei).cuda()
    a_d = torch.as_tensor(anchors).cuda()
    x_d = torch.randn(n, f, device="cuda")
    out_d = torch.empty(n, f + k_total, device="cuda")
    engine = dev.GeodesicEngine(n, ei.shape[1], K_PER_GPU)
    flush = torch.empty(256 * 1024 * 1024 // 4, dtype=torch.float32, device="cuda")
    flush_r = torch.zeros(256 * 1024 * 1024 // 4, dtype=torch.float32, device="cuda")

    def flush_l2(i):
        """Untimed, between steps: write 256 MiB (evicts every line of the previous step from the 126 MB L2),
        then read a second 256 MiB buffer so the L2 is left full of CLEAN foreign lines — after the write alone
        it is full of dirty ones, whose write-back would be charged to the next timed step."""
        flush.fill_(float(i))
        return flush_r.sum()

    peer = gpd.PeerAssembly(engine) if world > 1 else None
    deep_flags = []

    def step():
        if world == 1:
            engine.run(ei_d, a_d, x_d, out_d)
        else:
            # NVLink peer-to-peer assembly; the flag says (on the device) whether the packed
            # hop format was valid (hops <= 15) — checked once after the timed loop
            deep_flags.append(peer.run(ei_d, a_d, x_d, out_d)[1])

    # clocks are sampled from before the warm-up to the end of the e2e loop (the device-timed region
    # alone lasts a few milliseconds, shorter than one nvidia-smi sampling period)
    clocks = ClockSampler(local_rank)
    if rank == 0:
        clocks.start()
    for _ in range(args.warmup):
        step()
    barrier()
    e_unique = engine.csr.info()["num_edges"]
    starts = [torch.cuda.Event(enable_timing=True) for _ in range(args.steps)]
    stops = [torch.cuda.Event(enable_timing=True) for _ in range(args.steps)]
    bfs_ms = []
    launches0 = dev.launch_count()
    barrier()
    for i in range(args.steps):
        flush_l2(i)  # untimed: evict the previous step's lines from the 126 MB L2
        # untimed alignment: without it the ranks drift apart during the flush and the wait for the slowest
        # one inside the step's all-reduce lands in the timed region of the others (N=4: 0.439 -> 0.465 ms)
        if world > 1:
            dist.barrier()
        starts[i].record()
        step()
        stops[i].record()
        bfs_ms.append(engine.bfs.kernel_ms())
    barrier()
    launches = dev.launch_count() - launches0
    if peer is not None and peer.trace_events and rank == 0:
        peer.trace_events = peer.trace_events[-args.steps:]
        print("peer step phases (median ms): csr+bfs+pack %.4f, flag all-reduce %.4f, peer decode %.4f"
              % tuple(peer.trace_summary()), file=sys.stderr, flush=True)
    if world > 1 and int(deep_flags[-1].item()) != 0:
        raise RuntimeError("a shard had hops > 15: the packed exchange is invalid for this graph; use "
                           "distributed.sharded_geodesic_features (all-gather path)")
    step_ms = [s.elapsed_time(e) for s, e in zip(starts, stops)]
    total_ms = float(sum(step_ms))
    stats = engine.bfs.stats()

    if world > 1:
        t = torch.tensor([total_ms], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        total_ms = float(t.item())
    ms_per_step = total_ms / args.steps
    value = k_total * e_unique / (ms_per_step * 1e-3) / 1e9

    # ---- per-stage device times (N = 1 only; same inputs, each stage timed alone after the L2 flush)
    stages = None
    if world == 1:
        def timed(fn, reps=10):
            ms = []
            for i in range(reps):
                flush_l2(i)
                a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                a.record(); fn(); b.record(); torch.cuda.synchronize()
                ms.append(a.elapsed_time(b))
            return float(np.median(ms))
        engine.csr.build(ei_d); engine.bfs.run(a_d); engine.bfs.features(x_d, out_d); torch.cuda.synchronize()
        csr_ms = timed(lambda: engine.csr.build(ei_d))
        bfs_ms_alone = timed(lambda: engine.bfs.run(a_d))
        dec_ms = timed(lambda: engine.bfs.features(x_d, out_d))
        b_epi = 6 * n * K_PER_GPU + 8 * n * f  # B_epi + B_cat
        stages = {"csr_build_ms": csr_ms, "msbfs_run_ms": bfs_ms_alone, "decode_concat_ms": dec_ms,
                  "decode_concat_gbs": b_epi / (dec_ms * 1e-3) / 1e9,
                  "note": "each stage launched eagerly on its own (the step replays them from one CUDA graph); "
                          "decode_concat bytes = 6*N*K + 8*N*F"}

    # ---- e2e: the public host-buffer API, H2D + D2H inside the timed region
    ei_h = torch.as_tensor(ei).pin_memory()
    x_h = torch.randn(n, f).pin_memory()
    out_h = torch.empty(n, f + k_total).pin_memory()
    lo, hi = gpd.shard_bounds(k_total, world, rank)

    e2e_staging = {}

    def e2e_step():
        if world == 1:
            dev.geodesic_embed_host(ei_h, n, anchors, x_h, out=out_h)
        else:
            gpd.sharded_geodesic_embed_host(engine, ei_h, anchors, x_h, out_h, e2e_staging, peer=peer)

    for _ in range(3):
        e2e_step()
    barrier()
    e2e_steps = max(3, min(args.steps, 10))
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        e2e_step()
    barrier()
    e2e_s = (time.perf_counter() - t0) / e2e_steps
    if world > 1:
        t = torch.tensor([e2e_s], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e2e_s = float(t.item())
    e2e_value = k_total * e_unique / e2e_s / 1e9
    clock_info = clocks.stop() if rank == 0 else None

    if rank == 0:
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except Exception:
            pass
        peak = float(peaks.get("hbm_gbs", 6650.0))
        w_words = stats["lane_words"]
        # B_bfs = W*(4(N+1) + 4E + 8E + 16N) + 2*N*K_g algorithmic bytes per launch
        b_bfs = w_words * (4 * (n + 1) + 12 * e_unique + 16 * n) + 2 * n * K_PER_GPU
        bfs_avg_ms = float(np.mean(bfs_ms))
        achieved = b_bfs / (bfs_avg_ms * 1e-3) / 1e9
        traffic = None  # dram bytes read+written per launch from the committed ncu --set full capture
        try:
            traffic = json.load(open(os.path.join(ROOT, "profiles", "msbfs_traffic.json")))["dram_bytes_per_launch"]
        except Exception:
            pass
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "u64", "data": "synthetic",
            "config": {"workload": workload_name(shape, K_PER_GPU), "total_anchors": k_total,
                       "dedup_edges": e_unique, "parallelism": f"anchor-shard x{world}",
                       "l2": "between steps (untimed): 256 MiB buffer written, then a second 256 MiB buffer read, so the "
                             "L2 holds neither the previous step's lines nor dirty lines to write back",
                       "step": "csr build + ms-bfs + fused normalise/concat epilogue" +
                               (" + pack + NVLink peer-to-peer assembly in the epilogue" if world > 1 else "")},
            "roofline": {"bound": "hbm", "kernel": "msbfs_kernel", "achieved": achieved, "peak": peak,
                         "unit": "GB/s", "frac": achieved / peak, "traffic": traffic,
                         "peak_source": "MEASURED_PEAKS.json hbm_gbs (measured)" if peaks else "fallback 6650",
                         "algorithmic_bytes": b_bfs, "kernel_ms": bfs_avg_ms,
                         "kernel_share_of_step": bfs_avg_ms / ms_per_step},
            "bfs": {"levels": stats["levels_run"], "max_level": stats["max_level"],
                    "edges_examined_per_WE": stats["edges_examined"] / max(1, w_words * e_unique),
                    "grid_blocks": stats["grid_blocks"]},
            "e2e": {"value": e2e_value, "unit": UNIT, "ms": e2e_s * 1e3,
                    "h2d_bytes_per_step": int(ei_h.numel() * 8 + (hi - lo) * 8),
                    "d2h_bytes_per_step": int(n * k_total * 4),
                    "api": ("graphpope_b200.device.geodesic_embed_host (gp_geodesic_embed_host)" if world == 1 else
                            "graphpope_b200.distributed.sharded_geodesic_embed_host") +
                           ", pinned host buffers; x is concatenated on the host"},
            "gpu_launches": int(launches),
            "clocks": clock_info,
        }
        if stages:
            stages["decode_concat_frac_of_hbm_peak"] = stages["decode_concat_gbs"] / peak
            line["stages"] = stages
        if cpu_baseline:
            line["cpu_baseline"] = cpu_baseline
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
