"""bench.py without a GPU: the algorithmic-byte formulas the roofline fractions divide by and
the JSON line of the reference arm (the one leg of bench.py that runs on host cores only)."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def test_algorithmic_bytes_of_the_headline_workload():
    import bench

    n, e, k, f = 89250, 899756, 256, 500
    assert bench.bytes_bfs(n, e, k) == 4 * (4 * (n + 1) + 4 * e + 8 * e + 16 * n) + 2 * n * k == 96_024_304
    assert bench.bytes_epilogue(n, k, f) == 137_088_000 + 357_000_000
    assert bench.bytes_bfs(n, e, 1) == bench.bytes_bfs(n, e, 64) - 2 * n * 63  # one lane word up to 64 anchors
    assert bench.bytes_csr(n, e, e) == 32 * e + 8 * e + 28 * n
    cfg = bench.headline_config(__import__("graphpope_b200.synth", fromlist=["SHAPES"]).SHAPES["flickr-shape"], 4, e)
    assert cfg["total_anchors"] == 1024 and "256 anchors per GPU" in cfg["workload"] and "model" not in cfg


def test_reference_arm_prints_the_contract_line():
    env = dict(os.environ, GP_BENCH_REF_BUDGET_S="3", CUDA_VISIBLE_DEVICES="")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1",
                        "--warmup", "0"], capture_output=True, text=True, timeout=600, env=env, cwd=ROOT)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [ln for ln in r.stdout.splitlines() if ln.startswith("{")]
    assert len(lines) == 1
    line = json.loads(lines[0])
    assert line["impl"] == "reference" and line["metric"] and line["unit"] == line["cpu_baseline"]["unit"]
    assert line["higher_is_better"] is True and line["vs_baseline"] is None and line["n_gpus"] == 1
    assert line["value"] > 0 and line["value"] == line["cpu_baseline"]["value"] == line["e2e"]["value"]
    assert line["e2e"]["h2d_bytes_per_step"] == 0 and line["e2e"]["d2h_bytes_per_step"] == 0
    assert line["cpu_baseline"]["kind"] == "port" and line["cpu_baseline"]["cores"] >= 1
    assert "workload" in line["config"] and line["steps"] == 1 and line["warmup"] == 0
    # a non-zero rank of a torchrun launch prints nothing and exits 0
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2"],
                       capture_output=True, text=True, timeout=600, cwd=ROOT,
                       env=dict(env, RANK="1", WORLD_SIZE="2", LOCAL_RANK="1"))
    assert r.returncode == 0 and not [ln for ln in r.stdout.splitlines() if ln.startswith("{")]


def test_a_failing_secondary_line_is_reported_not_fatal(capsys):
    import bench

    def boom(x):
        raise RuntimeError("CUDA out of memory (simulated) %d" % x)

    res, ok = bench.guarded(boom, 7)
    assert ok is False and res["parity_checked"] is False and "simulated) 7" in res["error"]
    assert bench.guarded(lambda a, b: ({"v": a + b}, True), 1, 2) == ({"v": 3}, True)
