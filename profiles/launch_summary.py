"""Summarise an ncu launch list (`--metrics gpu__time_duration.sum --csv`): launches, total time and
share per kernel.  usage: launch_summary.py launches.csv "<command that was profiled>" """
import csv, sys, collections, re

path, what = sys.argv[1], (sys.argv[2] if len(sys.argv) > 2 else "")
rows = [r for r in csv.reader(open(path, errors="replace")) if len(r) > 10]
hdr = rows[0]
ik, iv, iu = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
tot = collections.Counter(); cnt = collections.Counter()
for r in rows[1:]:
    if r[hdr.index("Metric Name")] != "gpu__time_duration.sum":
        continue
    name = re.sub(r"\(.*", "", r[ik]).replace("void ", "").strip()
    v = float(r[iv].replace(",", "")) * {"ns": 1e-3, "us": 1.0, "ms": 1e3}.get(r[iu], 1e-3)
    tot[name] += v; cnt[name] += 1
total = sum(tot.values())
print(f"launches of `{what}` under ncu (gpu__time_duration.sum, cold-cache, serialised)")
for name, v in tot.most_common():
    print(f"{cnt[name]:5d} launches {v:12.1f} us {100 * v / total:5.1f}%  {name}")
