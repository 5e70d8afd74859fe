"""Developer tool: N sequential (single-stream, eager) passes at BASELINE configs[1] sizes, for
`ncu -k regex:... -s ... -c 1` captures of individual kernels.  usage: ncu_pass.py [passes] [B]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
from ffb6d_b200.pipeline import FusionPass  # noqa: E402
from ffb6d_b200.synthetic import make_batch  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 2
B = int(sys.argv[2]) if len(sys.argv) > 2 else 32
batch = make_batch(range(B))
dev = torch.device("cuda:0")
cld = torch.from_numpy(batch["cld"]).to(dev)
xyz = torch.from_numpy(batch["dpt_xyz"]).to(dev)
cho = torch.from_numpy(batch["choose"]).to(dev)
p = FusionPass(B, device=dev, n_streams=1)
for _ in range(n):
    p(cld, xyz, cho)
torch.cuda.synchronize()
print("done", n, "passes")
