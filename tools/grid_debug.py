"""Developer tool: dump the uniform-grid parameters the KNN picked for every call of the
schedule on one synthetic frame batch, and profile the Python side of an eager pass."""
import cProfile
import os
import pstats
import struct
import sys

os.environ["FFB6D_DEBUG_GRID"] = "1"
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
from ffb6d_b200 import ops  # noqa: E402
from ffb6d_b200.pipeline import FusionPass  # noqa: E402
from ffb6d_b200.schedule import knn_schedule  # noqa: E402
from ffb6d_b200.synthetic import make_batch  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 4
batch = make_batch(range(B))
dev = torch.device("cuda:0")
cld = torch.from_numpy(batch["cld"]).to(dev)
xyz = torch.from_numpy(batch["dpt_xyz"]).to(dev)
cho = torch.from_numpy(batch["choose"]).to(dev)
p = FusionPass(B, device=dev)
p.build_indices(cld, xyz, cho)
torch.cuda.synchronize()
print("%-18s %6s %6s %3s | %9s %4s %4s %4s %8s %7s %6s %5s" % (
    "call", "S", "Q", "K", "h", "nx", "ny", "nz", "ncells", "pts/occ", "ovf", "dup"))
for (Bc, S, Q, K, ws), in [(w,) for w in ops._debug_ws]:
    if ws is None:
        print("%-18s %6d %6d %3d | tiled scan" % ("", S, Q, K))
        continue
    raw = ws[:64].cpu().numpy().tobytes()
    lo = struct.unpack("3f", raw[0:12])
    h, inv_h, slack = struct.unpack("3f", raw[12:24])
    n = struct.unpack("3i", raw[24:36])
    ncells, ovf, dup, rep = struct.unpack("4i", raw[36:52])
    print("%-18s %6d %6d %3d | %9.5f %4d %4d %4d %8d %7s %6d %5d" % (
        "", S, Q, K, h, n[0], n[1], n[2], ncells, "", ovf, dup))

ops._debug_ws.clear()
ops._DEBUG_GRID = False
for _ in range(3):
    p(cld, xyz, cho)
torch.cuda.synchronize()
pr = cProfile.Profile()
pr.enable()
for _ in range(5):
    p(cld, xyz, cho)
torch.cuda.synchronize()
pr.disable()
pstats.Stats(pr).sort_stats("cumulative").print_stats(22)
