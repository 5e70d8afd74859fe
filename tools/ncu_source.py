"""Developer tool: SASS-level hot spots of an .ncu-rep captured with --import-source on: consecutive
instructions with (almost) equal execution counts are merged into segments.
usage: ncu_source.py report.ncu-rep [min_percent]"""
import csv
import io
import subprocess
import sys

path = sys.argv[1]
thr = float(sys.argv[2]) if len(sys.argv) > 2 else 0.5
out = subprocess.run(["ncu", "-i", path, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hdr = rows[1]
iS, iI, iT, iP = hdr.index("Source"), hdr.index("Instructions Executed"), hdr.index("Thread Instructions Executed"), hdr.index("# Samples")
data = [(n, r[iS].strip(), int(r[iI]), int(r[iT]), int(r[iP])) for n, r in enumerate(rows[2:]) if r[iI].isdigit()]
tot = sum(d[2] for d in data)
ts = max(1, sum(d[4] for d in data))
print("%s\n%d warp instructions, %d samples" % (rows[0][1][:100], tot, ts))
seg, cur = [], None
for d in data:
    if cur and abs(d[2] - cur["c"]) <= 0.02 * max(cur["c"], 1):
        cur["n"] += 1; cur["inst"] += d[2]; cur["thr"] += d[3]; cur["smp"] += d[4]; cur["end"] = d[0]
    else:
        cur = {"start": d[0], "end": d[0], "c": d[2], "n": 1, "inst": d[2], "thr": d[3], "smp": d[4], "first": d[1]}
        seg.append(cur)
for s in seg:
    if 100 * s["inst"] / tot > thr or 100 * s["smp"] / ts > thr:
        print("%4d-%4d n=%3d execs=%10d inst%%=%5.1f smp%%=%5.1f thr/inst=%4.1f  %s" % (
            s["start"], s["end"], s["n"], s["c"], 100 * s["inst"] / tot, 100 * s["smp"] / ts,
            s["thr"] / max(s["inst"], 1), s["first"][:48]))
