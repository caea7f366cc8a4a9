"""Print the interesting parts of a bench.py JSON line (skips library chatter around it)."""
import json, sys
for path in sys.argv[1:]:
    d = None
    for line in open(path):
        if line.startswith("{"):
            d = json.loads(line)
    if d is None:
        print(path, "no JSON line"); continue
    print(f"== {path}: N={d['n_gpus']} {d['ms_per_step']*1e3:.1f} us/step {d['value']:.0f} GTEPS parity={d.get('parity_checked')} e2e {d['e2e']['ms']:.2f} ms ({d['e2e']['value']:.0f} GTEPS)")
    st = d["roofline"].get("stages", {})
    print("   stages:", {k: (round(v["ms"]*1e3, 1), round(v["frac"], 3)) for k, v in st.items() if isinstance(v, dict)}, "bfs", d.get("bfs"))
    print("   parity:", {k: v for k, v in (d.get("parity") or {}).items() if k not in ("oracle", "checksum")})
    s = d.get("secondary", {})
    if "C3_node2vec_block" in s:
        print("   C3:", {m: round(v["us"], 1) for m, v in s["C3_node2vec_block"]["modes"].items()}, s["C3_node2vec_block"]["parity_checked"])
    if "C4_flickr_k1024_centrality_anchors" in s and "samplers" in s["C4_flickr_k1024_centrality_anchors"]:
        print("   C4:", {m: (round(v["step_ms"], 3), v["parity"]["columns_bit_equal"], v["anchor_list_equals_oracle"]) for m, v in s["C4_flickr_k1024_centrality_anchors"]["samplers"].items()})
    if s.get("C5_products_k4096"):
        c5 = s["C5_products_k4096"]; print("   C5:", round(c5["step_ms"], 2), "ms", round(c5["gteps"]), "GTEPS", {k: round(v, 2) for k, v in c5["stage_ms_rank0"].items()}, c5["parity_checked"])
    if "cold_one_shot_call" in s:
        print("   cold:", s["cold_one_shot_call"])
