"""CPU: host-side logic -- schedules, byte model, synthetic generator, argument checks."""
import numpy as np
import pytest
import torch

from ffb6d_b200 import schedule as S
from ffb6d_b200 import synthetic, ops
from ffb6d_b200._lib import LAYOUT_NCS, LAYOUT_NSC


def test_knn_schedule_matches_reference_table():
    calls = S.knn_schedule()
    assert len(calls) == 22
    sizes = [(S.set_size(s), S.set_size(q), k) for _, s, q, k in calls]
    assert sizes[0] == (12288, 12288, 16) and sizes[1] == (3072, 12288, 1)
    assert sizes[2] == (19200, 3072, 16) and sizes[3] == (3072, 19200, 1)
    assert sizes[-2] == (76800, 3072, 16) and sizes[-1] == (3072, 76800, 1)
    assert sum(q for _, q, _ in sizes) == 247152
    assert sum(q * k for _, q, k in sizes) == 613632
    assert sum(s * q for s, q, _ in sizes) == 926161920      # brute-force pairs / frame (9.26e8)
    keys = [c[0] for c in calls]
    assert len(set(keys)) == 22


def test_byte_model_matches_baseline_md():
    assert S.frame_alg_bytes() == (8239296, 160186368)
    kb, gb = S.frame_alg_bytes(4096)
    assert kb + gb == 128276544
    kb, gb = S.frame_alg_bytes(40960, k=32)
    assert kb + gb == 307077760
    kb, gb = S.frame_alg_bytes(131072, k=8)
    assert kb + gb == 495030272


def test_gather_schedule():
    g = S.gather_schedule()
    assert len(g) == 23
    assert sum(1 for x in g if x[0] == "random_sample") == 11
    assert sum(1 for x in g if x[0] == "nearest_interpolation") == 11
    assert g[-1] == ("choose", "choose", 64, 307200, 12288, 1)
    assert ("random_sample", "r2p_ds_nei_idx3", 1024, 4800, 48, 16) in g
    assert ("nearest_interpolation", "cld_interp_idx3", 512, 48, 192, 1) in g


def test_synthetic_frame_deterministic_and_tie_free():
    a = synthetic.make_frame(4, n_points=768)
    b = synthetic.make_frame(4, n_points=768)
    for k in a:
        assert np.array_equal(a[k], b[k])
    assert a["dpt_xyz"].shape == (480, 640, 3) and a["dpt_xyz"].dtype == np.float32
    assert a["cld"].shape == (768, 3) and a["choose"].shape == (1, 768)
    assert len(np.unique(a["choose"])) == 768                   # no 'wrap' padding
    holes = (a["dpt_xyz"][..., 2] == 0)
    assert 0.05 < holes.mean() < 0.15
    assert (a["dpt_xyz"][holes] == 0).all()
    assert np.array_equal(a["cld"], a["dpt_xyz"].reshape(-1, 3)[a["choose"][0]])
    pyr = synthetic.image_pyramid_np(a["dpt_xyz"])
    assert [len(pyr[s]) for s in (1, 2, 4, 8)] == [307200, 76800, 19200, 4800]
    assert np.array_equal(pyr[4].reshape(120, 160, 3), a["dpt_xyz"][::4, ::4])


def test_image_pyramid_torch_matches_numpy():
    a = synthetic.make_frame(1, n_points=768)
    t = S.image_pyramid(torch.from_numpy(a["dpt_xyz"])[None])
    n = synthetic.image_pyramid_np(a["dpt_xyz"])
    for sr in (1, 2, 4, 8):
        assert np.array_equal(t[sr][0].numpy(), n[sr])


def test_layout_detection():
    f = torch.zeros(2, 8, 10, 1)
    assert ops._layout_of(f.squeeze(3))[1] == LAYOUT_NCS
    cl = f.contiguous(memory_format=torch.channels_last)
    t, lay = ops._layout_of(cl.squeeze(3))
    assert lay == LAYOUT_NSC and t.data_ptr() == cl.data_ptr()
    odd = torch.zeros(2, 8, 20)[:, :, ::2]
    t, lay = ops._layout_of(odd)
    assert lay == LAYOUT_NCS and t.is_contiguous()


def test_shape_errors():
    with pytest.raises(ValueError):
        ops.knn_search(np.zeros((1, 4, 3), np.float32), np.zeros((2, 4, 3), np.float32), 2)
    with pytest.raises(RuntimeError):
        ops.grid_sub_sampling(np.zeros((4, 2), np.float32))


def test_synthetic_backprojection_is_the_references():
    """make_frame's organised cloud == the reference's dpt_2_pcld on the same depth map (digest made
    by tests/golden/make_golden.py from the reference's own source)."""
    import hashlib
    import json
    import os
    from conftest import GOLDEN
    d = json.load(open(os.path.join(GOLDEN, "backproject_digest.json")))
    for case in d.values():
        fr = synthetic.make_frame(case["seed"], n_points=768, intrinsics=case["intrinsics"])
        assert hashlib.sha256(np.ascontiguousarray(fr["depth"]).tobytes()).hexdigest() == case["sha256_depth"]
        assert hashlib.sha256(np.ascontiguousarray(fr["dpt_xyz"]).tobytes()).hexdigest() == case["sha256_xyz_f32"]


def test_tile_kernel_reciprocal_division_is_exact():
    """grid_search_k1_tile_kernel (csrc/knn_grid.cu) replaces `tile / tiles_x` and `lane / by` by
    `(int)((i + 0.5f) * (1.0f / n))` whenever `tiles + tiles_x < 4 000 000` (the launcher passes `exact_div` otherwise).
    numpy float32 performs the same IEEE round-to-nearest operations: the identity holds on the whole admitted range."""
    for n in list(range(1, 48)) + [80, 160, 320, 333, 640, 1000, 4096, 65535, 1000003, 3999999]:
        t = np.arange(0, 4000000 - n, dtype=np.int64)
        got = ((t.astype(np.float32) + np.float32(0.5)) * (np.float32(1.0) / np.float32(n))).astype(np.int64)
        assert np.array_equal(got, t // n), n
    lanes = np.arange(32)
    for by in range(1, 33):
        got = ((lanes.astype(np.float32) + np.float32(0.5)) * (np.float32(1.0) / np.float32(by))).astype(np.int64)
        assert np.array_equal(got, lanes // by), by
