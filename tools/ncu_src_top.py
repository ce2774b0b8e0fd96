"""Top stalled SASS instructions of one kernel from `ncu -i rep --page source --csv --kernel-id :::N`."""
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
hdr = rows[1]
idx = {h: i for i, h in enumerate(hdr)}
data = [r for r in rows[2:] if len(r) == len(hdr) and r[idx['# Samples']].isdigit()]
tot = sum(int(r[idx['# Samples']]) for r in data)
print("kernel:", rows[0][1][:100], "| total samples", tot, "| instructions", len(data))
stall_cols = [h for h in hdr if h.startswith('stall_') and 'Not Issued' not in h]
agg = {h: sum(int(r[idx[h]]) for r in data) for h in stall_cols}
print("stall totals:", sorted(((v, k) for k, v in agg.items() if v), reverse=True)[:8])
n = int(sys.argv[2]) if len(sys.argv) > 2 else 40
for r in sorted(data, key=lambda r: -int(r[idx['# Samples']]))[:n]:
    st = sorted(((int(r[idx[h]]), h[6:]) for h in stall_cols if int(r[idx[h]]) > 0), reverse=True)[:3]
    print(r[idx['# Samples']].rjust(6), r[idx['Source']].strip()[:60].ljust(60), st, 'exec', r[idx['Instructions Executed']])
