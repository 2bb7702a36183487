"""Top instructions by warp-stall samples from an .ncu-rep (SASS view of `ncu --page source`), with the dominant
stall reason of each: python profiles/ncu_hot.py <report> [top N]"""
import csv, io, subprocess, sys
out = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hdr = next(r for r in rows if r and r[0] == "Address")
body = rows[rows.index(hdr) + 1:]
col = {h: i for i, h in enumerate(hdr)}
stalls = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
total = sum(int(r[col["# Samples"]] or 0) for r in body)
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
print("total samples", total)
agg = {}
for r in body:
    for h in stalls:
        agg[h] = agg.get(h, 0) + int(r[col[h]] or 0)
print("by reason:", ", ".join(f"{k[6:]}={v * 100 // max(total, 1)}%" for k, v in sorted(agg.items(), key=lambda kv: -kv[1]) if v * 100 >= total))
for k, r in sorted(enumerate(body), key=lambda kr: -int(kr[1][col["# Samples"]] or 0))[:top]:
    n = int(r[col["# Samples"]] or 0)
    why = sorted(((int(r[col[h]] or 0), h[6:]) for h in stalls), reverse=True)[:2]
    print(f"{k:5d} {n * 100.0 / total:5.1f}%  {r[col['Source']].strip()[:70]:70s} {why[0][1]}:{why[0][0]} {why[1][1]}:{why[1][0]}")
