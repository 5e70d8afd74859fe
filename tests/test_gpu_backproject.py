"""GPU: depth -> point sets (ffb6d_backproject) against the reference's dpt_2_pcld outputs
(tests/golden/backproject_digest.json) and the index build from depth against the reference's
22 index arrays."""
import hashlib
import json
import os

import numpy as np
import pytest
import torch

import ffb6d_b200 as F
from conftest import GOLDEN
from ffb6d_b200.synthetic import INTRINSICS, image_pyramid_np, make_frame

pytestmark = pytest.mark.gpu


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


@pytest.mark.parametrize("case", ["seed0_linemod", "seed7_ycb_K1"])
def test_backproject_bit_exact(cuda, case):
    d = json.load(open(os.path.join(GOLDEN, "backproject_digest.json")))[case]
    fr = make_frame(d["seed"], n_points=768, intrinsics=d["intrinsics"])
    assert sha(fr["depth"]) == d["sha256_depth"]
    assert sha(fr["dpt_xyz"]) == d["sha256_xyz_f32"]          # generator == reference (pinned on CPU too)
    depth = torch.from_numpy(fr["depth"])[None].cuda()
    choose = torch.from_numpy(fr["choose"])[None].cuda()
    cld, pyr = F.backproject(depth, INTRINSICS[d["intrinsics"]], choose)
    assert np.array_equal(cld[0].cpu().numpy().view(np.uint32), fr["cld"].view(np.uint32))   # bitwise incl. -0.0
    ref = image_pyramid_np(fr["dpt_xyz"])
    for sr in (2, 4, 8):
        assert np.array_equal(pyr[sr][0].cpu().numpy().view(np.uint32), ref[sr].view(np.uint32)), sr
    # per-frame intrinsics + batch of 2
    K2 = np.stack([INTRINSICS[d["intrinsics"]]] * 2)
    cld2, _ = F.backproject(depth.repeat(2, 1, 1), K2, choose.repeat(2, 1, 1))
    assert torch.equal(cld2[0], cld[0]) and torch.equal(cld2[1], cld[0])


def test_indices_from_depth_match_reference_digest(cuda):
    d = json.load(open(os.path.join(GOLDEN, "schedule_digest.json")))["frames"]["seed0_n12288"]
    fr = make_frame(0, n_points=12288)
    depth = torch.from_numpy(fr["depth"])[None].cuda()
    choose = torch.from_numpy(fr["choose"])[None].cuda()
    inputs = F.build_ffb6d_indices_from_depth(depth, INTRINSICS["linemod"], choose)
    for key, meta in d["keys"].items():
        assert sha(inputs[key][0].cpu().numpy()) == meta["sha256"], key


def test_sample_valid_pixels_properties(cuda):
    """Device-side valid-pixel compaction + seeded sampling (ycb_dataset.py:218-235): every pick is a valid pixel,
    picks are distinct when enough valid pixels exist, deterministic per seed and different across seeds /
    frames, close to uniform; with fewer valid pixels than points every valid pixel is used floor(N/n) or
    ceil(N/n) times ('wrap')."""
    fr = [make_frame(s, n_points=768) for s in (0, 1)]
    depth = torch.from_numpy(np.stack([f["depth"] for f in fr])).cuda()
    N = 12288
    ch, cnt = F.sample_valid_pixels(depth, N, seed=5, return_count=True)
    assert ch.shape == (2, 1, N) and ch.dtype == torch.int32
    c = ch.cpu().numpy()[:, 0]
    for b in range(2):
        valid = fr[b]["depth"].reshape(-1) > 1e-8
        assert int(cnt[b]) == int(valid.sum())
        assert valid[c[b]].all()
        assert len(np.unique(c[b])) == N
        # roughly uniform over the image: chi-square over a 12 x 16 grid of tiles, expected ~ N * valid share
        tiles = (c[b] // 640 // 40) * 16 + (c[b] % 640) // 40
        obs = np.bincount(tiles, minlength=192).astype(np.float64)
        exp = N * np.array([valid.reshape(480, 640)[r * 40:(r + 1) * 40, q * 40:(q + 1) * 40].sum()
                            for r in range(12) for q in range(16)]) / valid.sum()
        assert ((obs - exp) ** 2 / exp).sum() < 192 + 6 * (2 * 192) ** 0.5
        assert abs(np.corrcoef(np.arange(N), c[b])[0, 1]) < 0.05        # order carries no raster trend
    assert not np.array_equal(c[0], c[1])
    assert torch.equal(F.sample_valid_pixels(depth, N, seed=5), ch)                 # deterministic
    assert not torch.equal(F.sample_valid_pixels(depth, N, seed=6), ch)
    # fewer valid pixels than points: 'wrap'
    few = torch.zeros(1, 48, 64).cuda()
    few[0, 10:20, 10:40] = 1.0                                                        # 300 valid pixels
    ch2, cnt2 = F.sample_valid_pixels(few, 1000, seed=1, return_count=True)
    assert int(cnt2[0]) == 300
    times = np.bincount(ch2.cpu().numpy().reshape(-1), minlength=48 * 64)
    assert set(np.unique(times[times > 0])) <= {3, 4} and (times > 0).sum() == 300
    assert (few.reshape(-1)[ch2.reshape(-1).long()] > 0).all()
    # the picks feed the index build: depth -> choose -> cloud -> 22 index tensors, all on the device
    cld, pyr = F.backproject(depth, INTRINSICS["linemod"], ch)
    assert torch.isfinite(cld).all() and (cld[..., 2] > 0).all()
