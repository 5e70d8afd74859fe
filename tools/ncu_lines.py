"""Developer tool: per-CUDA-source-line instruction counts and stall samples of an .ncu-rep captured with
--import-source on (kernels built with -lineinfo).  usage: ncu_lines.py report.ncu-rep [top_n]"""
import csv
import io
import subprocess
import sys

path = sys.argv[1]
top = int(sys.argv[2]) if len(sys.argv) > 2 else 50
out = subprocess.run(["ncu", "-i", path, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
data, fname, hdr = [], "", None
for r in rows:
    if len(r) == 2 and r[0] == "File Path":
        fname = r[1].split("/")[-1]
    elif len(r) > 8 and r[0] == "Line No":
        hdr = r
        iI, iP, iT = hdr.index("Instructions Executed"), hdr.index("# Samples"), hdr.index("Thread Instructions Executed")
    elif hdr and len(r) == len(hdr) and r[0].isdigit():
        data.append((fname, int(r[0]), r[1].strip(), int(r[iI]), int(r[iP]), int(r[iT])))
tot, ts = sum(d[3] for d in data), max(1, sum(d[4] for d in data))
print("%d warp instructions, %d samples" % (tot, ts))
for d in sorted(data, key=lambda d: -d[3])[:top]:
    print("%-14s %5d inst%%=%5.1f smp%%=%5.1f lanes=%4.1f  %s" % (d[0][:14], d[1], 100.0 * d[3] / tot, 100.0 * d[4] / ts, d[5] / max(d[3], 1), d[2][:100]))
