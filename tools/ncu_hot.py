"""Hot spots of an ncu --set full --import-source on capture (development helper):
   ncu -i x.ncu-rep --page source --csv --print-source sass > src.csv ; python tools/ncu_hot.py src.csv [top N]
Prints the instructions with the most warp-stall samples and what they were stalled on, and the totals per stall reason."""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
hdr = rows[1]
col = {h: i for i, h in enumerate(hdr)}
stall_cols = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
data = []
for r in rows[2:]:
    if len(r) < len(hdr):
        continue
    n = int(r[col["# Samples"]] or 0)
    data.append((n, r))
total = sum(n for n, _ in data)
print("total samples", total, " instructions executed", sum(int(r[col["Instructions Executed"]] or 0) for _, r in data))
tot = {s: sum(int(r[col[s]] or 0) for _, r in data) for s in stall_cols}
print("by reason:", {k: v for k, v in sorted(tot.items(), key=lambda x: -x[1]) if v})
for n, r in sorted(data, key=lambda x: -x[0])[:top]:
    reasons = {s[6:]: int(r[col[s]] or 0) for s in stall_cols if int(r[col[s]] or 0)}
    print(f"{n:6d} {100.0 * n / total:5.1f}%  {r[col['Address']][-5:]}  {r[col['Source']][:60]:60s} exec {r[col['Instructions Executed']]:>8s}  {dict(sorted(reasons.items(), key=lambda x: -x[1])[:4])}")
