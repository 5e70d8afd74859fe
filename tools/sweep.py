"""BASELINE configs[4]: stress sweep N0 in {4096, 12288, 40960, 131072} x K in {8, 16, 32} on one
GPU through bench.py (graph replay, inputs resident); prints one markdown row per point and
writes the JSON lines to gpurun_out/sweep.jsonl.
usage: python tools/sweep.py [--steps 10]"""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
steps = sys.argv[sys.argv.index("--steps") + 1] if "--steps" in sys.argv else "10"
# frames per step chosen so that the per-step working set stays at a few GB for every N0
BATCH = {4096: 32, 12288: 32, 40960: 16, 131072: 8}
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
out = open(os.path.join(ROOT, "gpurun_out", "sweep.jsonl"), "w")
print("| N0 | K | frames/step | ms/step | points/s | alg GB/step | pass GB/s | frac of HBM peak |")
print("|---:|---:|---:|---:|---:|---:|---:|---:|")
for n0 in (4096, 12288, 40960, 131072):
    for k in (8, 16, 32):
        r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--steps", steps, "--warmup", "3",
                            "--n-points", str(n0), "--k", str(k), "--batch", str(BATCH[n0]),
                            "--no-cpu-baseline", "--no-mlp"], capture_output=True, text=True, timeout=900)
        if r.returncode != 0:
            print("| %d | %d | failed: %s |" % (n0, k, r.stderr.strip().splitlines()[-1:]))
            continue
        j = json.loads(r.stdout.strip().splitlines()[-1])
        out.write(json.dumps(j) + "\n")
        out.flush()
        pr = j.get("pass_roofline", {})
        print("| %d | %d | %d | %.3f | %.3e | %.2f | %.0f | %.3f |" % (
            n0, k, BATCH[n0], j["ms_per_step"], j["value"], pr.get("alg_bytes_per_frame", 0) * BATCH[n0] / 1e9,
            pr.get("achieved", 0), pr.get("frac", 0)), flush=True)
