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
