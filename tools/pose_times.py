"""Developer tool: time keypoint voting (9 vote sets of one object) against the reference's algorithm written
with stock torch ops on the same GPU (N x N matrices, one host sync per iteration, one call per keypoint)."""
import math
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
import ffb6d_b200 as F  # noqa: E402


def torch_mean_shift(A, bw=0.04, max_iter=300):
    N = A.shape[0]
    C = A.clone()
    it = 0
    while True:
        it += 1
        dis = torch.norm(C.reshape(1, N, 3) - C.reshape(N, 1, 3), dim=2)
        w = (torch.exp(-0.5 * (dis / bw) ** 2) / (bw * math.sqrt(2 * math.pi))).reshape(N, N, 1)
        new_C = torch.sum(w * C, dim=1) / torch.sum(w, dim=1)
        Cdis = torch.norm(new_C - C, dim=1)
        C = new_C
        if torch.max(Cdis) < bw * 1e-3 or it > max_iter:
            break
    dis = torch.norm(C.view(N, 1, 3) - C.view(1, N, 3), dim=2)
    num_in = torch.sum(dis < bw, dim=1)
    _, mi = torch.max(num_in, 0)
    return C[mi], dis[mi] < bw, it


def votes(G, N, seed=0):
    g = torch.Generator().manual_seed(seed)
    truth = torch.rand(G, 1, 3, generator=g) * 0.3 + torch.tensor([0.0, 0.0, 0.8])
    v = truth + torch.randn(G, N, 3, generator=g) * 0.01
    v[:, ::7] += torch.rand(G, (N + 6) // 7, 3, generator=g) * 0.4 - 0.2
    return v.cuda()


def timed(fn, n=3):
    fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n


for N in (1024, 4096, 12288):
    v = votes(9, N)
    t_ours = timed(lambda: F.mean_shift_fit(v, None, 0.04, 300))
    c, lab, it = F.mean_shift_fit(v, None, 0.04, 300)
    t_torch = None
    if N <= 4096:
        t_torch = timed(lambda: [torch_mean_shift(v[g]) for g in range(9)], n=1)
        ct = torch.stack([torch_mean_shift(v[g])[0] for g in range(9)])
        err = (ct - c).abs().max().item()
    pairs = float(N) * N * float(it.sum().item())        # sets stop at different rounds
    print("N=%5d: ours %.3f ms (%d rounds, %.1f Gpair/s)%s" % (
        N, t_ours, it.max().item(), pairs / t_ours / 1e6,
        "" if t_torch is None else " | torch ops %.1f ms (%.0fx) max |dcentre| %.2e" % (t_torch, t_torch / t_ours, err)), flush=True)
