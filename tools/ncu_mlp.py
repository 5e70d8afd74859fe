"""Developer tool: a few launches of one fusion-MLP layer for `ncu -k regex:fusion_mlp_packed -s 2 -c 1`.
usage: ncu_mlp.py [C1 C2 Co P [B]]   (default: ds3 p2r_fuse, 1024+1024 -> 1024 at 60x80, B = 32)"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
import ffb6d_b200 as F  # noqa: E402

a = [int(v) for v in sys.argv[1:]]
C1, C2, Co, P = a[:4] if len(a) >= 4 else (1024, 1024, 1024, 4800)
B = a[4] if len(a) > 4 else 32
x1 = torch.randn(B, C1, P, 1, device="cuda")
x2 = torch.randn(B, C2, P, 1, device="cuda") if C2 else None
w = F.fusion_mlp_pack(torch.randn(Co, C1 + C2, device="cuda") / (C1 + C2) ** 0.5)
sc = torch.rand(Co, device="cuda") + 0.5
sh = torch.randn(Co, device="cuda")
for _ in range(4):
    y = F.fusion_mlp(x1, x2, w, sc, sh)
torch.cuda.synchronize()
print("done", float(y.abs().mean()))
