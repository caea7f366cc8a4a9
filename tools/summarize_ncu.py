"""Turn ncu outputs brought back in gpurun_out/ into the small text summaries kept under profiles/."""
import collections, csv, subprocess, sys

def launches(path):
    rows = list(csv.reader(open(path))); hdr = None; data = []
    for r in rows:
        if 'Kernel Name' in r: hdr = r; continue
        if hdr and len(r) == len(hdr): data.append(dict(zip(hdr, r)))
    agg = collections.OrderedDict()
    for d in data:
        k = d['Kernel Name'][:70]; v = float(d['Metric Value'].replace(',', ''))
        a = agg.setdefault(k, [0, 0.0]); a[0] += 1; a[1] += v
    tot = sum(a[1] for a in agg.values())
    out = ["kernel | launches | total us | avg us | share of all profiled GPU time", "---|---|---|---|---"]
    for k, a in sorted(agg.items(), key=lambda x: -x[1][1]):
        out.append("%s | %d | %.1f | %.2f | %.3f" % (k, a[0], a[1] / 1e3, a[1] / a[0] / 1e3, a[1] / tot))
    return "\n".join(out)

def raw(rep, keys):
    txt = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(txt.splitlines())); hdr, units = rows[0], rows[1]
    out = []
    for vals in rows[2:]:
        name = vals[hdr.index("Kernel Name")]
        out.append("## " + name[:90])
        for i, h in enumerate(hdr):
            if any(h == k or h.startswith(k + ".") and h.count(".") <= k.count(".") + 1 for k in keys) and vals[i] not in ("", "n/a"):
                out.append("%s [%s] = %s" % (h, units[i], vals[i]))
    return "\n".join(out)

KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__t_sector_hit_rate.pct", "l1tex__t_sector_hit_rate.pct", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
        "l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "launch__registers_per_thread", "launch__grid_size", "launch__block_size", "smsp__inst_executed.sum",
        "sm__cycles_elapsed.max", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_membar_per_issue_active.ratio"]

if __name__ == "__main__":
    if sys.argv[1] == "launches": print(launches(sys.argv[2]))
    else: print(raw(sys.argv[2], KEYS))
