"""GPU: gather + max-pool, nearest gather, neighbour gather, relative position encoding,
forward (bitwise: pure selections) and backward (1e-5: atomic summation order), through the
C ABI, against the reference's outputs (tests/golden) and the oracle."""
import numpy as np
import pytest
import torch

import ffb6d_b200 as F
from oracle import cpu_oracle as O

pytestmark = pytest.mark.gpu


def t(a, **kw):
    return torch.from_numpy(np.ascontiguousarray(a)).cuda(**kw)


@pytest.mark.parametrize("idx_dtype", [torch.int64, torch.int32])
@pytest.mark.parametrize("channels_last", [False, True])
@pytest.mark.parametrize("name", ["rs_small", "rs_k8", "rs_wide"])
def test_random_sample_golden(cuda, gather_golden, name, channels_last, idx_dtype):
    c = gather_golden[name]
    feat = t(c["feat"]).requires_grad_(True)
    x = feat.contiguous(memory_format=torch.channels_last) if channels_last else feat
    idx = t(c["idx"]).to(idx_dtype)
    out = F.random_sample(x, idx)
    assert out.shape == c["out"].shape
    assert np.array_equal(out.detach().cpu().numpy(), c["out"])            # bitwise
    out.backward(t(c["gout"]))
    np.testing.assert_allclose(feat.grad.cpu().numpy(), c["gfeat"], rtol=1e-5, atol=1e-6)


@pytest.mark.parametrize("channels_last", [False, True])
@pytest.mark.parametrize("name", ["ni_small", "ni_wide"])
def test_nearest_interpolation_golden(cuda, gather_golden, name, channels_last):
    c = gather_golden[name]
    feat = t(c["feat"]).requires_grad_(True)
    x = feat.contiguous(memory_format=torch.channels_last) if channels_last else feat
    out = F.nearest_interpolation(x, t(c["idx"]))
    assert np.array_equal(out.detach().cpu().numpy(), c["out"])
    out.backward(t(c["gout"]))
    np.testing.assert_allclose(feat.grad.cpu().numpy(), c["gfeat"], rtol=1e-5, atol=1e-6)


@pytest.mark.parametrize("name", ["gn_xyz", "gn_feat", "gn_odd"])
def test_gather_neighbour_golden(cuda, gather_golden, name):
    c = gather_golden[name]
    pc = t(c["pc"]).requires_grad_(True)
    out = F.gather_neighbour(pc, t(c["idx"]))
    assert np.array_equal(out.detach().cpu().numpy(), c["out"])
    out.backward(t(c["gout"]))
    np.testing.assert_allclose(pc.grad.cpu().numpy(), c["gpc"], rtol=1e-5, atol=1e-6)


def test_relative_pos_encoding(cuda, gather_golden):
    c = gather_golden["rpe"]
    got = F.relative_pos_encoding(t(c["xyz"]), t(c["idx"])).cpu().numpy()
    np.testing.assert_allclose(got, c["out"], rtol=1e-5, atol=1e-6)          # vs reference (torch CPU)
    assert np.array_equal(got, O.relative_pos_encoding(c["xyz"], c["idx"]))   # vs oracle: bitwise


# every (C, S, Q, K) of FFB6D.forward, one frame, fp32 N(0,1) features
def _fwd_cases():
    from ffb6d_b200.schedule import gather_schedule
    return [(op, key, C, S, Q, K) for op, key, C, S, Q, K in gather_schedule()]


@pytest.mark.parametrize("case", _fwd_cases(), ids=lambda c: "%s-%s" % (c[0], c[1]))
def test_schedule_shapes_vs_oracle(cuda, case):
    op, key, C, S, Q, K = case
    import zlib
    g = torch.Generator().manual_seed(zlib.crc32(key.encode()) % 1000)
    B = 2
    feat = torch.randn(B, C, S, 1, generator=g)
    idx = torch.randint(0, S, (B, Q, K), generator=g, dtype=torch.int64)
    if op == "random_sample":
        want = O.random_sample(feat.numpy(), idx.numpy())
        got = F.random_sample(feat.cuda(), idx.cuda())
        got_cl = F.random_sample(feat.cuda().contiguous(memory_format=torch.channels_last), idx.cuda().int())
    elif op == "nearest_interpolation":
        want = O.nearest_interpolation(feat.numpy(), idx.numpy())
        got = F.nearest_interpolation(feat.cuda(), idx.cuda())
        got_cl = F.nearest_interpolation(feat.cuda().contiguous(memory_format=torch.channels_last), idx.cuda().int())
    else:
        B = 1
        emb = feat[:1].reshape(1, C, 480, 640)
        ch = idx[:1].reshape(1, 1, Q)
        want = O.choose_gather(emb.numpy(), ch.numpy())
        got = F.choose_gather(emb.cuda(), ch.cuda())
        got_cl = F.choose_gather(emb.cuda().contiguous(memory_format=torch.channels_last), ch.cuda().int())
    assert np.array_equal(got.cpu().numpy(), want)
    assert np.array_equal(got_cl.cpu().numpy(), want)


@pytest.mark.parametrize("shape", [(5, 64, 19200, 3072, 16), (3, 64, 12288, 3072, 16), (2, 128, 4800, 2500, 16),
                                   (3, 96, 3072, 1500, 8), (1, 160, 768, 192, 16), (7, 33, 1000, 1025, 16)],
                         ids=lambda s: "B%d-C%d-S%d-Q%d-K%d" % s)
def test_random_sample_large_shapes(cuda, shape):
    """Max-pool gathers at batch sizes where B*C exceeds the SM count (every dispatch branch of launch_ncs at
    more than one frame), odd channel counts, a partial last query group, NaN / -inf rows, last-element indices."""
    B, C, S, Q, K = shape
    g = torch.Generator().manual_seed(B * 1000 + C)
    feat = torch.randn(B, C, S, 1, generator=g)
    feat[B - 1, C // 2, ::97] = float("nan")
    feat[0, 1, :] = float("-inf")
    idx = torch.randint(0, S, (B, Q, K), generator=g, dtype=torch.int64)
    idx[:, 0, :] = S - 1                                           # the last element of a row
    want = O.random_sample(feat.numpy(), idx.numpy())
    for ii in (idx.cuda(), idx.cuda().int()):
        got = F.random_sample(feat.cuda(), ii).cpu().numpy()
        assert np.array_equal(np.isnan(got), np.isnan(want))
        assert np.array_equal(got[~np.isnan(got)], want[~np.isnan(want)])


def test_nan_and_inf_propagate_like_torch_max(cuda):
    feat = torch.randn(1, 4, 20, 1)
    feat[0, 0, 3] = float("nan")
    feat[0, 1, 5] = float("inf")
    feat[0, 2, :] = float("-inf")
    idx = torch.randint(0, 20, (1, 50, 16))
    idx[0, 0, 7] = 3
    want = torch.gather(feat.squeeze(3), 2, idx.reshape(1, -1).unsqueeze(1).repeat(1, 4, 1)) \
        .reshape(1, 4, 50, 16).max(dim=3, keepdim=True)[0]
    got = F.random_sample(feat.cuda(), idx.cuda()).cpu()
    assert torch.equal(torch.isnan(got), torch.isnan(want))
    assert torch.equal(got[~torch.isnan(got)], want[~torch.isnan(want)])


def test_generic_k_and_three_dim_feature(cuda):
    g = torch.Generator().manual_seed(5)
    feat = torch.randn(2, 9, 333, generator=g)                      # 3-D accepted (ffb6d.py:166-167)
    for K in (2, 3, 5, 12, 33, 64):
        idx = torch.randint(0, 333, (2, 70, K), generator=g)
        want = O.random_sample(feat.numpy(), idx.numpy())
        assert np.array_equal(F.random_sample(feat.cuda(), idx.cuda()).cpu().numpy(), want)
        assert np.array_equal(F.random_sample(feat.cuda(), idx.cuda().int()).cpu().numpy(), want)


def test_backward_matches_torch_autograd_on_gpu(cuda):
    """Gradient parity against the reference expression evaluated by torch autograd on the GPU."""
    g = torch.Generator().manual_seed(9)
    feat = torch.randn(2, 32, 500, 1, generator=g).cuda()
    idx = torch.randint(0, 500, (2, 200, 16), generator=g).cuda()
    go = torch.randn(2, 32, 200, 1, generator=g).cuda()
    a = feat.clone().requires_grad_(True)
    ref = torch.gather(a.squeeze(3), 2, idx.reshape(2, -1).unsqueeze(1).repeat(1, 32, 1)) \
        .reshape(2, 32, 200, 16).max(dim=3, keepdim=True)[0]
    ref.backward(go)
    b = feat.clone().requires_grad_(True)
    out = F.random_sample(b, idx)
    out.backward(go)
    assert torch.equal(out, ref)
    torch.testing.assert_close(b.grad, a.grad, rtol=1e-5, atol=1e-6)
