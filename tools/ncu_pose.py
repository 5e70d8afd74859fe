"""Developer tool: one keypoint-voting launch (9 vote sets x N points) for `ncu -k regex:mean_shift_kernel -c 1`.
usage: ncu_pose.py [N] [max_iter]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
import ffb6d_b200 as F  # noqa: E402

N = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
max_iter = int(sys.argv[2]) if len(sys.argv) > 2 else 30
g = torch.Generator().manual_seed(0)
truth = torch.rand(9, 1, 3, generator=g) * 0.3 + torch.tensor([0.0, 0.0, 0.8])
v = truth + torch.randn(9, N, 3, generator=g) * 0.01
v[:, ::7] += torch.rand(9, (N + 6) // 7, 3, generator=g) * 0.4 - 0.2
c, lab, it = F.mean_shift_fit(v.cuda(), None, 0.04, max_iter)
torch.cuda.synchronize()
print("done", it.tolist())
