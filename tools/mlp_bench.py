"""Developer tool: time the fused fusion-MLP kernel on the four big p2r_fuse shapes (SURVEY.md
App. A.3) against torch's conv+BN+ReLU (cuDNN, fp32 with and without TF32)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
import torch.nn.functional as Fn  # noqa: E402
import ffb6d_b200 as F  # noqa: E402

B = 32
shapes = [(64, 64, 64, 19200), (128, 128, 128, 4800), (512, 512, 512, 4800), (1024, 1024, 1024, 4800),
          (256, 256, 256, 19200), (64, 64, 64, 76800)]


def timeit(fn, n=5):
    for _ in range(2):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n


for C1, C2, Co, P in shapes:
    x1 = torch.randn(B, C1, P, 1, device="cuda")
    x2 = torch.randn(B, C2, P, 1, device="cuda")
    w = torch.randn(Co, C1 + C2, 1, 1, device="cuda") / (C1 + C2) ** 0.5
    sc = torch.rand(Co, device="cuda") + 0.5
    sh = torch.randn(Co, device="cuda")
    flops = 2.0 * B * Co * (C1 + C2) * P

    wp = F.fusion_mlp_pack(w)

    def ours():
        return F.fusion_mlp(x1, x2, wp, sc, sh)

    def ours_raw():
        return F.fusion_mlp(x1, x2, w, sc, sh)

    def torch_ref():
        y = Fn.conv2d(torch.cat((x1, x2), 1), w)
        return torch.relu(y * sc.view(1, -1, 1, 1) + sh.view(1, -1, 1, 1))

    t_ours = timeit(ours)
    t_raw = timeit(ours_raw)
    torch.backends.cudnn.allow_tf32 = True
    t_tf32 = timeit(torch_ref)
    torch.backends.cudnn.allow_tf32 = False
    t_fp32 = timeit(torch_ref)
    print("Ci=%4d Co=%4d P=%6d : ours %.3f ms (%.1f TFLOP/s fp32-equiv; raw weights %.3f ms) | torch tf32 %.3f ms | "
          "torch fp32 %.3f ms" % (C1 + C2, Co, P, t_ours, flops / t_ours / 1e9, t_raw, t_tf32, t_fp32), flush=True)
