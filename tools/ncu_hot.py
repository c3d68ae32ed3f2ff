"""Top stall-sample SASS instructions of an ncu source-page CSV (ncu -i X.ncu-rep --page source --csv)."""
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
n = int(sys.argv[2]) if len(sys.argv) > 2 else 30
hi = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
hdr = rows[hi]
ci = {h: i for i, h in enumerate(hdr)}
stall_cols = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
data = []
tot = 0
for idx, r in enumerate(rows[hi + 1:]):
    try:
        v = int(r[ci["# Samples"]])
    except (ValueError, IndexError):
        continue
    tot += v
    data.append((v, idx, r))
print("total samples", tot, "instructions", len(data))
for v, idx, r in sorted(data, key=lambda t: -t[0])[:n]:
    st = sorted(((int(r[ci[c]]), c[6:]) for c in stall_cols if r[ci[c]] not in ("", "0")), reverse=True)[:3]
    print(f"{v:6d} {100 * v / tot:5.1f}%  #{idx:5d} {r[ci['Source']].strip()[:70]:70s} {st}")
