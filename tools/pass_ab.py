"""Developer tool: A/B timing of scheduling variants of the pass inside ONE process (captured CUDA graphs,
interleaved rounds so that clock / thermal drift hits every variant alike) + a bitwise check that every variant
produces the results of the sequential single-stream pass.  usage: pass_ab.py [B] [rounds] [replays]"""
import hashlib
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
from ffb6d_b200.pipeline import FusionPass  # noqa: E402
from ffb6d_b200.synthetic import make_batch  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 32
rounds = int(sys.argv[2]) if len(sys.argv) > 2 else 5
replays = int(sys.argv[3]) if len(sys.argv) > 3 else 20
only = sys.argv[4].split(",") if len(sys.argv) > 4 else None
dev = torch.device("cuda:0")
batch = make_batch(range(B))
cld = torch.from_numpy(batch["cld"]).to(dev)
xyz = torch.from_numpy(batch["dpt_xyz"]).to(dev)
cho = torch.from_numpy(batch["choose"]).to(dev)


def digest(inputs, outs):
    h = hashlib.sha256()
    for k in sorted(inputs):
        if "idx" in k:
            h.update(inputs[k].contiguous().cpu().numpy().tobytes())
    for o in outs:
        h.update(o.contiguous().cpu().numpy().tobytes())
    return h.hexdigest()[:16]


class Split:
    """Two half-batch passes software-pipelined inside one graph (each half has its own streams)."""

    def __init__(self, n, **kw):
        self.n = n
        self.parts = [FusionPass(B // n, device=dev, seed=0, **kw) for _ in range(n)]
        full = FusionPass(B, device=dev, seed=0, n_streams=1)
        for j, p in enumerate(self.parts):      # same features as the full-batch pass, sliced
            p.features = [f[j * (B // n):(j + 1) * (B // n)] for f in full.features]
        self.forks = [torch.cuda.Stream(device=dev) for _ in range(n)]
        self.device = dev

    def capture(self, fn):
        return self.parts[0].capture(fn)

    def __call__(self, cld, xyz, cho):
        main = torch.cuda.current_stream(dev)
        res = []
        m = B // self.n
        for j, (p, st) in enumerate(zip(self.parts, self.forks)):
            st.wait_stream(main)
            with torch.cuda.stream(st):
                res.append(p(cld[j * m:(j + 1) * m], xyz[j * m:(j + 1) * m], cho[j * m:(j + 1) * m]))
        for st in self.forks:
            main.wait_stream(st)
        inputs = {k: torch.cat([r[0][k] for r in res]) for k in res[0][0]}
        outs = [torch.cat([r[1][i] for r in res]) for i in range(len(res[0][1]))]
        return inputs, outs


def mk(**kw):
    return lambda: FusionPass(B, device=dev, **kw)


def with_env(make, **env):
    def f():
        os.environ.update({k: str(v) for k, v in env.items()})     # read by the scheduler while the variant is captured
        return make()
    return f


variants = {
    "base": with_env(mk(), FFB6D_SUBSET_NN=0, FFB6D_SELF_FIRST=1),
    "subset_nn": with_env(mk(), FFB6D_SUBSET_NN=1, FFB6D_SELF_FIRST=1),
    "subset_nn_noselffirst": with_env(mk(), FFB6D_SUBSET_NN=1, FFB6D_SELF_FIRST=0),
    "noselffirst": with_env(mk(), FFB6D_SUBSET_NN=0, FFB6D_SELF_FIRST=0),
}
if only:
    variants = {k: v for k, v in variants.items() if k in only}

seq = FusionPass(B, device=dev, n_streams=1)
ref = digest(*seq(cld, xyz, cho))
del seq
print("sequential digest", ref, flush=True)
runs = {}
for name, make in variants.items():
    p = make()
    for _ in range(3):
        p(cld, xyz, cho)
    torch.cuda.synchronize()
    if isinstance(p, Split):
        def fn(p=p):      # time the halves only (the concatenation is for the check)
            main = torch.cuda.current_stream(dev)
            m = B // p.n
            res = []
            for j, (q, st) in enumerate(zip(p.parts, p.forks)):
                st.wait_stream(main)
                with torch.cuda.stream(st):
                    res.append(q(cld[j * m:(j + 1) * m], xyz[j * m:(j + 1) * m], cho[j * m:(j + 1) * m]))
            for st in p.forks:
                main.wait_stream(st)
            return res
        d = digest(*p(cld, xyz, cho))
    else:
        def fn(p=p):
            return p(cld, xyz, cho)
        d = None
    rep = p.capture(fn)
    for _ in range(3):
        res = rep()
    torch.cuda.synchronize()
    if d is None:
        d = digest(*res)
    print("%-14s digest %s %s" % (name, d, "OK" if d == ref else "MISMATCH"), flush=True)
    runs[name] = (rep, [], p)      # p owns the feature tensors the graph reads: keep it alive
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
for r in range(rounds):
    for name, (rep, ts, _) in runs.items():
        torch.cuda.synchronize()
        e0.record()
        for _ in range(replays):
            rep()
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1) / replays)
for name, (rep, ts, _) in runs.items():
    ts = sorted(ts)
    print("%-14s median %.4f ms  min %.4f  max %.4f" % (name, ts[len(ts) // 2], ts[0], ts[-1]), flush=True)
