"""Developer tool: sweep the grid-search knobs (cell scale, estimate quantile) and time the
index build of one 32-frame batch (sequential launches, CUDA events)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
from ffb6d_b200 import _lib  # noqa: E402
from ffb6d_b200.pipeline import FusionPass  # noqa: E402
from ffb6d_b200.synthetic import make_batch  # noqa: E402

B = 32
batch = make_batch(range(B))
dev = torch.device("cuda:0")
cld = torch.from_numpy(batch["cld"]).to(dev)
xyz = torch.from_numpy(batch["dpt_xyz"]).to(dev)
cho = torch.from_numpy(batch["choose"]).to(dev)
p = FusionPass(B, device=dev, n_streams=1)
if "--k1" in sys.argv:      # cell scale of the K = 1 grids only
    combos = [(s, 17, s1) for s1 in (0.7, 1.0, 1.4, 2.0, 2.8, 4.0) for s in (1.0,)]
else:
    combos = [(s, q, 1.0) for q in (8, 13, 17, 21, 25, 29) for s in (0.8, 1.0, 1.25, 1.6)]
for scale, quant, scale1 in combos:
    _lib.lib.ffb6d_knn_grid_tune(scale, quant)
    _lib.lib.ffb6d_knn_grid_tune_k1(scale1)
    for _ in range(2):
        p.build_indices(cld, xyz, cho)
    torch.cuda.synchronize()
    torch.cuda._sleep(int(20e6))
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(3):
        p.build_indices(cld, xyz, cho)
    e1.record()
    torch.cuda.synchronize()
    print("scale %.2f quantile %2d scale_k1 %.2f : %.3f ms per index build" % (scale, quant, scale1, e0.elapsed_time(e1) / 3), flush=True)
