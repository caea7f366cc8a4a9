"""Top stalled SASS instructions from `ncu --page source --csv` output (per kernel section)."""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
topn = int(sys.argv[2]) if len(sys.argv) > 2 else 25
hdr = None; data = []
for r in rows:
    if r and r[0] == "Address": hdr = r; continue
    if hdr and len(r) == len(hdr) and r[0].startswith("0x"): data.append(r)
ix = {h: i for i, h in enumerate(hdr)}
tot = sum(int(r[ix['# Samples']]) for r in data)
print('total samples', tot, 'instructions', len(data))
keys = [h for h in hdr if h.startswith('stall_') and 'Not Issued' not in h]
for pos, r in sorted(enumerate(data), key=lambda t: -int(t[1][ix['# Samples']]))[:topn]:
    st = {k[6:]: int(r[ix[k]]) for k in keys if int(r[ix[k]])}
    print("%5d %6s %5.1f%% exec=%8s  %-58s %s" % (pos, r[ix['# Samples']], 100 * int(r[ix['# Samples']]) / max(tot, 1),
          r[ix['Instructions Executed']], r[ix['Source']].strip()[:58], st))
