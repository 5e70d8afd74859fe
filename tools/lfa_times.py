"""Developer tool: per-kernel GPU time (CUPTI) of the four encoder Dilated_res_blocks at FFB6D widths, inference.
usage: lfa_times.py [B]"""
import collections
import os
import re
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
from torch.profiler import ProfilerActivity, profile  # noqa: E402
from ffb6d_b200 import modules as M  # noqa: E402
from ffb6d_b200.schedule import build_ffb6d_indices  # noqa: E402
from ffb6d_b200.synthetic import make_batch  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 32
dev = torch.device("cuda:0")
batch = make_batch(range(B))
inp = build_ffb6d_indices(torch.from_numpy(batch["cld"]).to(dev), torch.from_numpy(batch["dpt_xyz"]).to(dev))
blocks, feats, d_in = [], [], 8
g = torch.Generator(device=dev).manual_seed(0)
for i, d in enumerate((32, 64, 128, 256)):
    blocks.append(M.Dilated_res_block(d_in, d).to(dev).eval())
    feats.append(torch.randn((B, d_in, 12288 // 4 ** i, 1), generator=g, device=dev))
    d_in = 2 * d


def run():
    with torch.no_grad():
        return [blk(f, inp["cld_xyz%d" % i], inp["cld_nei_idx%d" % i]) for i, (blk, f) in enumerate(zip(blocks, feats))]


for _ in range(3):
    run()
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    run()
    torch.cuda.synchronize()
agg = collections.OrderedDict()
for ev in prof.events():
    if ev.device_type is not None and "cuda" in str(ev.device_type).lower():
        name = re.sub(r"\(.*", "", ev.name).replace("void ", "").replace("ffb6d::", "")
        d = agg.setdefault(name, [0, 0.0])
        d[0] += 1
        d[1] += ev.device_time
tot = sum(v[1] for v in agg.values())
print("B=%d: %.3f ms of kernel time, %d launches" % (B, tot / 1e3, sum(v[0] for v in agg.values())))
for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print("%-70s n=%3d %9.1f us %5.1f%%" % (k[:70], v[0], v[1], 100 * v[1] / tot))
