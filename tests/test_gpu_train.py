"""GPU: TRAINING mode of the 1x1 layers and RandLA blocks (batch-statistics BatchNorm, backward) against
the reference's own modules run in training mode with autograd (tests/golden/train_cases.npz, made by
tests/golden/make_golden.py from models/pytorch_utils.py and models/RandLA/RandLANet.py).  The modules are
loaded through ``load_state_dict(strict=True)`` with the reference's state dicts: parameter names are part
of the contract.  Floating point: 1e-5 of each tensor's scale."""
import os

import numpy as np
import pytest
import torch

import ffb6d_b200 as F
from ffb6d_b200 import modules as M
from conftest import GOLDEN

pytestmark = pytest.mark.gpu


def load(name):
    z = np.load(os.path.join(GOLDEN, "train_cases.npz"))
    return {k[len(name) + 1:]: z[k] for k in z.files if k.startswith(name + "/")}


def close(got, want, what, tol=1e-5):
    got = got.detach().cpu().numpy().astype(np.float64) if isinstance(got, torch.Tensor) else np.asarray(got, np.float64)
    want = np.asarray(want, np.float64)
    assert got.shape == want.shape, (what, got.shape, want.shape)
    scale = max(np.abs(want).max(), 1e-3)
    err = np.abs(got - want).max()
    assert err <= tol * scale, "%s: max abs err %.3e at scale %.3e (%.2e relative)" % (what, err, scale, err / scale)


def sd_of(case, prefix="sd."):
    return {k[len(prefix):]: torch.from_numpy(v) for k, v in case.items() if k.startswith(prefix)}


@pytest.mark.parametrize("name,C1,C2,Co", [("conv_cat", 24, 40, 48), ("conv_pre", 64, 0, 32), ("conv_wide", 256, 256, 128)])
def test_fusion_conv_train_matches_reference(cuda, name, C1, C2, Co):
    c = load(name)
    layer = M.Conv2d(C1 + C2, Co, kernel_size=(1, 1), bn=True)
    sd = sd_of(c)
    sd["normlayer.bn.num_batches_tracked"] = torch.tensor(0)
    layer.load_state_dict(sd, strict=True)
    layer.cuda().train()
    x1 = torch.from_numpy(c["x1"]).cuda().requires_grad_(True)
    x2 = torch.from_numpy(c["x2"]).cuda().requires_grad_(True) if C2 else None
    out = layer(x1, x2)
    close(out, c["out"], name + " out")
    out.backward(torch.from_numpy(c["gout"]).cuda())
    close(x1.grad, c["gx1"], name + " grad x1")
    if C2:
        close(x2.grad, c["gx2"], name + " grad x2")
    close(layer.conv.weight.grad, c["gw"], name + " grad W")
    close(layer.normlayer.bn.weight.grad, c["ggamma"], name + " grad gamma")
    close(layer.normlayer.bn.bias.grad, c["gbeta"], name + " grad beta")
    close(layer.normlayer.bn.running_mean, c["after.normlayer.bn.running_mean"], name + " running_mean")
    close(layer.normlayer.bn.running_var, c["after.normlayer.bn.running_var"], name + " running_var")
    assert int(layer.normlayer.bn.num_batches_tracked) == 1
    # the same layer on the materialised concat (one input tensor) gives the same bits
    layer.zero_grad()
    xa = torch.cat((x1.detach(), x2.detach()), 1) if C2 else x1.detach()
    assert torch.equal(layer(xa), out)
    # eval mode: the fused inference kernel == torch's eval-mode layer
    layer.eval()
    with torch.no_grad():
        got = layer(x1.detach(), x2.detach() if C2 else None)
        bn = layer.normlayer.bn
        ref = torch.relu(torch.nn.functional.batch_norm(
            torch.nn.functional.conv2d(xa.double(), layer.conv.weight.double()), bn.running_mean.double(),
            bn.running_var.double(), bn.weight.double(), bn.bias.double(), False, 0.0, bn.eps))
    close(got, ref.cpu().numpy(), name + " eval")


@pytest.mark.parametrize("name,d_in,d_out", [("blk_train_8_16", 8, 16), ("blk_train_64_64", 64, 64)])
def test_dilated_res_block_train_matches_reference(cuda, name, d_in, d_out):
    c = load(name)
    blk = M.Dilated_res_block(d_in, d_out)
    sd = sd_of(c)
    for k in list(blk.state_dict()):
        if k.endswith("num_batches_tracked"):
            sd[k] = torch.tensor(0)
    blk.load_state_dict(sd, strict=True)          # the reference's parameter names
    blk.cuda().train()
    feature = torch.from_numpy(c["feature"]).cuda().requires_grad_(True)
    xyz, idx = torch.from_numpy(c["xyz"]).cuda(), torch.from_numpy(c["idx"]).cuda()
    out = blk(feature, xyz, idx)
    close(out, c["out"], name + " out", tol=2e-5)
    out.backward(torch.from_numpy(c["gout"]).cuda())
    close(feature.grad, c["gfeature"], name + " grad feature", tol=5e-5)
    for k, prm in blk.named_parameters():
        close(prm.grad, c["grad." + k], name + " grad " + k, tol=5e-5)
    for k, v in blk.state_dict().items():
        if "running_" in k:
            close(v, c["after." + k], name + " " + k)
    # eval mode (frozen statistics, fused residual GEMM) agrees with the modules' own eval composition
    blk.eval()
    with torch.no_grad():
        got = blk(feature.detach(), xyz, idx)
        f_pc = blk.mlp2(blk.lfa(xyz, blk.mlp1(feature.detach()), idx))
        want = torch.nn.functional.leaky_relu(f_pc + blk.shortcut(feature.detach()), 0.2)
    close(got, want.cpu().numpy(), name + " eval", tol=2e-5)


def test_att_pool_backward_vs_autograd(cuda):
    g = torch.Generator().manual_seed(11)
    for (B, C1, C2, N, K) in ((2, 16, 16, 50, 16), (1, 5, 0, 33, 7)):
        f1 = torch.randn(B, C1, N, K, generator=g).cuda().requires_grad_(True)
        f2 = torch.randn(B, C2, N, K, generator=g).cuda().requires_grad_(True) if C2 else None
        att = (torch.randn(B, C1 + C2, N, K, generator=g) * 2).cuda().requires_grad_(True)
        go = torch.randn(B, C1 + C2, N, 1, generator=g).cuda()
        out = M._AttPool.apply(f1, f2, att)
        out.backward(go)
        f1d, attd = f1.detach().double().requires_grad_(True), att.detach().double().requires_grad_(True)
        f2d = f2.detach().double().requires_grad_(True) if C2 else None
        fs = torch.cat((f1d, f2d), 1) if C2 else f1d
        ref = torch.sum(fs * torch.softmax(attd, dim=3), dim=3, keepdim=True)
        ref.backward(go.double())
        close(f1.grad, f1d.grad.cpu().numpy(), "att_pool grad f1")
        close(att.grad, attd.grad.cpu().numpy(), "att_pool grad att")
        if C2:
            close(f2.grad, f2d.grad.cpu().numpy(), "att_pool grad f2")


@pytest.mark.parametrize("B,C1,C2,Co,P", [(2, 64, 64, 64, 3072), (1, 36, 8, 70, 50), (3, 5, 0, 3, 7), (2, 1024, 1024, 1024, 600),
                                          (8, 64, 64, 64, 19200), (2, 10, 0, 16, 98304), (2, 32, 32, 64, 1001), (1, 16, 16, 32, 3072)])
def test_wgrad_vs_float64(cuda, B, C1, C2, Co, P):
    from ffb6d_b200._lib import lib, check
    from ffb6d_b200.ops import _stream
    g = torch.Generator().manual_seed(P)
    dz = torch.randn(B, Co, P, generator=g).cuda()
    x1 = torch.randn(B, C1, P, generator=g).cuda()
    x2 = torch.randn(B, C2, P, generator=g).cuda() if C2 else None
    gw = torch.full((Co, C1 + C2), float("nan"), device="cuda")
    check(lib.ffb6d_fusion_mlp_wgrad(dz.data_ptr(), x1.data_ptr(), C1, x2.data_ptr() if C2 else None, C2, B, Co, P,
                                     gw.data_ptr(), _stream(dz.device)))
    x = torch.cat((x1, x2), 1) if C2 else x1
    want = torch.einsum("bop,bcp->oc", dz.double(), x.double())
    close(gw, want.cpu().numpy(), "wgrad")
    ref32 = torch.einsum("bop,bcp->oc", dz, x)
    e_ours = (gw.double() - want).abs().max().item()
    e_32 = (ref32.double() - want).abs().max().item()
    assert e_ours <= max(8 * e_32, 1e-6 * want.abs().max().item()), (e_ours, e_32)
