"""GPU: the restructured fusion stage (W . cat(rgb, interp(p)) = W1 . rgb + interp(W2 . p))
against (a) the reference's own composition of one stage -- its pt_utils.Conv2d layers and its
random_sample / nearest_interpolation, executed from its source (tests/golden/fusion_cases.npz) -- and
(b) a float64 evaluation of the same lines of FFB6D.forward (models/ffb6d.py:245-263) at a FULL-SIZE stage
(ds3: 1024-channel image features at 60x80, 48 points).  Floating point: 1e-5 of the output scale and
1e-5 element-wise relative above a floor of 10 % of the scale."""
import os

import numpy as np
import pytest
import torch

import ffb6d_b200 as F
from ffb6d_b200 import fusion
from conftest import GOLDEN

pytestmark = pytest.mark.gpu


def close(got, want, what, tol=1e-5):
    got, want = np.asarray(got, np.float64), np.asarray(want, np.float64)
    scale = max(np.abs(want).max(), 1.0)
    d = np.abs(got - want)
    assert d.max() <= tol * scale, "%s: max abs err %.3e at output scale %.3e" % (what, d.max(), scale)
    big = np.abs(want) > 0.1 * scale
    if big.any():
        rel = (d[big] / np.abs(want[big])).max()
        assert rel <= 2 * tol, "%s: element-wise relative err %.3e" % (what, rel)


def layer_from_sd(sd, prefix):
    w = sd[prefix + ".conv.weight"]
    var, mean = sd[prefix + ".normlayer.bn.running_var"], sd[prefix + ".normlayer.bn.running_mean"]
    scale = sd[prefix + ".normlayer.bn.weight"] / torch.sqrt(var + 1e-5)
    return fusion.FusedConv(w, scale, sd[prefix + ".normlayer.bn.bias"] - mean * scale)


@pytest.mark.parametrize("name", ["stage_64", "stage_ragged"])
def test_stage_matches_reference_composition(cuda, name):
    z = np.load(os.path.join(GOLDEN, "fusion_cases.npz"))
    c = {k[len(name) + 1:]: z[k] for k in z.files if k.startswith(name + "/")}
    sd = {k[3:]: torch.from_numpy(v).cuda() for k, v in c.items() if k.startswith("sd.")}
    stage = fusion.FusionStage(*(layer_from_sd(sd, n) for n in ("r2p_pre", "r2p_fuse", "p2r_pre", "p2r_fuse")))
    rgb0, p0 = torch.from_numpy(c["rgb_emb0"]).cuda(), torch.from_numpy(c["p_emb0"]).cuda()
    p2r_idx, r2p_idx = torch.from_numpy(c["p2r_idx"]).cuda(), torch.from_numpy(c["r2p_idx"]).cuda()
    for restructured in (True, False):
        rgb, p = stage(rgb0, p0, p2r_idx, r2p_idx, restructured=restructured)
        close(rgb.cpu().numpy(), c["rgb_emb"], "%s rgb_emb (restructured=%s)" % (name, restructured))
        close(p.cpu().numpy(), c["p_emb"], "%s p_emb" % name)
    # int32 indices (what the index build produces) give the same bits as int64
    rgb32, _ = stage(rgb0, p0, p2r_idx.int(), r2p_idx.int())
    assert torch.equal(rgb32, stage(rgb0, p0, p2r_idx, r2p_idx)[0])


# (C_r, C_p, h, w, N'): ds3 -- the 20 GFLOP/frame layer -- and up1 (K <= 128 kernel variant, 240x320 map)
@pytest.mark.parametrize("shape", [(1024, 512, 60, 80, 48), (64, 128, 240, 320, 768)], ids=["ds3", "up1"])
def test_full_size_stage_vs_float64(cuda, shape):
    Cr, Cp, h, w, N1 = shape
    B, K = 2, 16
    g = torch.Generator().manual_seed(Cr)

    def rnd(*s):
        return torch.randn(s, generator=g)

    def mk(cin, cout):
        wgt = (rnd(cout, cin) / cin ** 0.5).cuda()
        scale = (torch.rand(cout, generator=g) + 0.5).cuda()
        shift = (rnd(cout) * 0.1).cuda()
        return fusion.FusedConv(wgt, scale, shift)

    layers = {"r2p_pre": mk(Cr, Cp), "r2p_fuse": mk(2 * Cp, Cp), "p2r_pre": mk(Cp, Cr), "p2r_fuse": mk(2 * Cr, Cr)}
    stage = fusion.FusionStage(layers["r2p_pre"], layers["r2p_fuse"], layers["p2r_pre"], layers["p2r_fuse"])
    rgb0 = rnd(B, Cr, h, w).cuda()
    p0 = rnd(B, Cp, N1, 1).cuda()
    p2r_idx = torch.randint(0, N1, (B, h * w, 1), generator=g).cuda()
    r2p_idx = torch.randint(0, h * w, (B, N1, K), generator=g).cuda()
    rgb, p = stage(rgb0, p0, p2r_idx, r2p_idx)

    def layer64(L, x):     # conv(bias=False) -> BatchNorm(eval, folded) -> ReLU in float64
        y = torch.einsum("oc,bcn->bon", L.weight.double(), x.flatten(2).double())
        return torch.relu(y * L.scale.double()[None, :, None] + L.shift.double()[None, :, None])

    # models/ffb6d.py:245-263 in float64
    p2r = layer64(layers["p2r_pre"], p0)                                               # [B,Cr,N1]
    up = torch.gather(p2r, 2, p2r_idx.view(B, 1, -1).expand(-1, Cr, -1))               # nearest_interpolation
    want_rgb = layer64(layers["p2r_fuse"], torch.cat((rgb0.flatten(2).double(), up), 1)).view(B, Cr, h, w)
    f = rgb0.flatten(2).double()
    nb = torch.gather(f, 2, r2p_idx.view(B, 1, -1).expand(-1, Cr, -1)).view(B, Cr, N1, K).max(dim=3)[0]   # random_sample
    want_p = layer64(layers["r2p_fuse"], torch.cat((p0.flatten(2).double(), layer64(layers["r2p_pre"], nb)), 1))
    close(rgb.cpu().numpy(), want_rgb.cpu().numpy(), "rgb_emb")
    close(p.flatten(2).cpu().numpy(), want_p.cpu().numpy(), "p_emb")
    # the epilogue pieces on their own: channels-last store == transpose of the NCHW store
    y = layers["p2r_pre"](p0)
    zt = F.fusion_mlp(y, None, stage.split.w2, stage.split.ones, stage.split.zeros, relu=False, out_channels_last=True)
    zn = F.fusion_mlp(y, None, stage.split.w2, stage.split.ones, stage.split.zeros, relu=False)
    assert torch.equal(zt, zn.flatten(2).transpose(1, 2).contiguous())
