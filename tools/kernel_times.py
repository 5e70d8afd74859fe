"""Developer tool: per-kernel GPU time of the pass from torch.profiler (CUPTI), normal clocks,
warm caches -- complements the cold/serialised ncu launch list.  usage: kernel_times.py [B] [steps]"""
import collections
import os
import re
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
from torch.profiler import ProfilerActivity, profile  # noqa: E402
from ffb6d_b200.pipeline import FusionPass  # noqa: E402
from ffb6d_b200.synthetic import make_batch  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 32
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
layout = sys.argv[3] if len(sys.argv) > 3 else "nchw"
batch = make_batch(range(B))
dev = torch.device("cuda:0")
cld = torch.from_numpy(batch["cld"]).to(dev)
xyz = torch.from_numpy(batch["dpt_xyz"]).to(dev)
cho = torch.from_numpy(batch["choose"]).to(dev)
p = FusionPass(B, device=dev, layout=layout, n_streams=int(os.environ.get("NSTREAMS", "1")))
for _ in range(3):
    p(cld, xyz, cho)
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    for _ in range(steps):
        p(cld, xyz, cho)
    torch.cuda.synchronize()
agg = collections.OrderedDict()
seq = []
for ev in prof.events():
    if ev.device_type is not None and "cuda" in str(ev.device_type).lower():
        name = re.sub(r"\(.*", "", ev.name).replace("void ", "").replace("ffb6d::", "")
        d = agg.setdefault(name, [0, 0.0])
        d[0] += 1
        d[1] += ev.device_time
        seq.append((ev.time_range.start, name, ev.device_time))
tot = sum(v[1] for v in agg.values())
print("B=%d steps=%d layout=%s  total kernel time %.3f ms/step" % (B, steps, layout, tot / steps / 1e3))
for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print("%-64s n/step=%5.1f %9.1f us/step %5.1f%%" % (k[:64], v[0] / steps, v[1] / steps, 100 * v[1] / tot))
if os.environ.get("SEQ"):
    seq.sort()
    n = len(seq) // steps
    for t, name, d in seq[:n]:
        print("%-60s %8.1f" % (name[:60], d))
