"""Developer tool: text report of a tools/pass_timeline.py JSON: per-replay span, busy/idle time, concurrency,
and the kernels in start order with their streams.  usage: timeline_report.py timeline.json [replay]"""
import json
import re
import sys

rows = json.load(open(sys.argv[1]))
which = int(sys.argv[2]) if len(sys.argv) > 2 else 1
n = len(rows) // 3
rep = rows[which * n:(which + 1) * n]
t0 = min(r["ts"] for r in rep)
t1 = max(r["ts"] + r["dur"] for r in rep)
print("replay %d: %d gpu events, span %.1f us, sum of durations %.1f us" % (which, len(rep), t1 - t0, sum(r["dur"] for r in rep)))
ev = []
for r in rep:
    ev.append((r["ts"], 1))
    ev.append((r["ts"] + r["dur"], -1))
ev.sort()
lvl, last, hist = 0, t0, {}
for t, d in ev:
    hist[lvl] = hist.get(lvl, 0.0) + (t - last)
    last = t
    lvl += d
print("time at concurrency level:", {k: round(v, 1) for k, v in sorted(hist.items())})
print("streams:", sorted({r["stream"] for r in rep}))


def short(nm):
    nm = re.sub(r"void |ffb6d::|\(.*", "", nm)
    return nm[:44]


for r in rep:
    print("%8.1f %8.1f  s%-3s %-46s grid %s" % (r["ts"] - t0, r["dur"], r["stream"], short(r["name"]), r["grid"]))
