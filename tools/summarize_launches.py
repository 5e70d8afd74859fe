"""Aggregate an `ncu --metrics gpu__time_duration.sum --csv` launch list into per-kernel shares of
one pass.  usage: summarize_launches.py launches.csv > summary.md"""
import collections
import csv
import re
import sys

rows = list(csv.reader(open(sys.argv[1])))
hdr, seq = None, []
for r in rows:
    if "Kernel Name" in r:
        hdr = r
        continue
    if hdr is None or len(r) != len(hdr):
        continue
    d = dict(zip(hdr, r))
    name = re.sub(r"\((const|int|ffb6d|unsigned|float).*", "", d["Kernel Name"])
    name = name.replace("void ffb6d::", "").replace("ffb6d::", "").replace("void ", "")
    seq.append((name, float(d["Metric Value"].replace(",", "")) / 1e3))
# one pass = from the first grid_prepare after a gather/copy to the next one
starts = [i for i, (n, _) in enumerate(seq) if "grid_prepare" in n and i > 0 and "grid_" not in seq[i - 1][0]
          and "knn_" not in seq[i - 1][0]]
a, b = (starts[0], starts[1]) if len(starts) >= 2 else (0, len(seq))
one = seq[a:b]
tot = sum(t for _, t in one)
agg = collections.OrderedDict()
for n, t in one:
    x = agg.setdefault(n, [0, 0.0])
    x[0] += 1
    x[1] += t
print("| kernel | launches / pass | device time / pass (us) | share |")
print("|---|---:|---:|---:|")
for n, (c, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print("| `%s` | %d | %.1f | %.1f %% |" % (n[:90], c, t, 100 * t / tot))
print("| **total** | %d | %.1f | 100 %% |" % (len(one), tot))
