"""Developer tool: GPU timeline (kernel start / end / stream, from CUPTI through torch.profiler) of CUDA-graph replays
of the pass: which kernels overlap, where the device idles.  usage: pass_timeline.py out.json [B] [variant env]"""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
from torch.profiler import ProfilerActivity, profile  # noqa: E402
from ffb6d_b200.pipeline import FusionPass  # noqa: E402
from ffb6d_b200.synthetic import make_batch  # noqa: E402

out = sys.argv[1]
B = int(sys.argv[2]) if len(sys.argv) > 2 else 32
dev = torch.device("cuda:0")
batch = make_batch(range(B))
cld = torch.from_numpy(batch["cld"]).to(dev)
xyz = torch.from_numpy(batch["dpt_xyz"]).to(dev)
cho = torch.from_numpy(batch["choose"]).to(dev)
p = FusionPass(B, device=dev)
for _ in range(3):
    p(cld, xyz, cho)
torch.cuda.synchronize()
rep = p.capture(lambda: p(cld, xyz, cho))
for _ in range(3):
    rep()
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    for _ in range(3):
        rep()
    torch.cuda.synchronize()
tmp = out + ".trace.json"
prof.export_chrome_trace(tmp)
tr = json.load(open(tmp))
ev = [e for e in tr["traceEvents"] if e.get("cat") in ("kernel", "gpu_memcpy", "gpu_memset")]
rows = sorted(({"name": e["name"][:90], "ts": e["ts"], "dur": e["dur"], "stream": e.get("args", {}).get("stream"),
                "grid": e.get("args", {}).get("grid"), "block": e.get("args", {}).get("block"),
                "regs": e.get("args", {}).get("registers per thread"), "smem": e.get("args", {}).get("shared memory")}
               for e in ev), key=lambda r: r["ts"])
json.dump(rows, open(out, "w"))
os.remove(tmp)
print("wrote", out, len(rows), "gpu events")
