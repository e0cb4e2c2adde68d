"""Summarise an ncu source-page CSV: top SASS lines by stall samples with their dominant stall reason."""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
hdr = rows[1]
ia, isrc, isamp = hdr.index("Address"), hdr.index("Source"), hdr.index("# Samples")
stall = [(i, h) for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
data = []
for k, r in enumerate(rows[2:]):
    try:
        s = float(r[isamp])
    except Exception:
        continue
    data.append((s, k, r))
tot = sum(d[0] for d in data)
print("total samples", tot)
n = int(sys.argv[2]) if len(sys.argv) > 2 else 30
for s, k, r in sorted(data, key=lambda x: -x[0])[:n]:
    top = sorted(((float(r[i] or 0), h) for i, h in stall), reverse=True)[:2]
    print(f"{s:8.0f} {100*s/tot:5.1f}%  #{k:5d} {r[isrc].strip()[:70]:70s} {top[0][1]}={top[0][0]:.0f} {top[1][1]}={top[1][0]:.0f}")
