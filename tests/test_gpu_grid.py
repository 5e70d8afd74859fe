"""GPU: voxel-grid subsampling through ffb6d_grid_subsample_host against the reference's
outputs (tests/golden/grid_cases.npz) and the oracle.  Barycentres and mean features are
bitwise equal; labels may differ only where several labels tie for the maximal count."""
import numpy as np
import pytest

import ffb6d_b200 as F
from oracle import cpu_oracle as O

pytestmark = pytest.mark.gpu


def lexsorted(p, *others):
    o = np.lexsort((p[:, 2], p[:, 1], p[:, 0]))
    return [p[o]] + [x[o] for x in others]


def label_is_a_mode(points, labels, grid, sub_points_sorted_by_key, sub_labels, keys):
    """every chosen label must have the maximal count inside its voxel"""
    pts = np.asarray(points, np.float32)
    mn = pts.min(0)
    inv = np.float32(1) / np.float32(grid)
    org = np.floor(mn * inv) * np.float32(grid)
    mx = pts.max(0)
    NX = np.uint64(np.floor((mx[0] - org[0]) / np.float32(grid))) + np.uint64(1)
    NY = np.uint64(np.floor((mx[1] - org[1]) / np.float32(grid))) + np.uint64(1)
    ijk = np.floor((pts - org) / np.float32(grid)).astype(np.uint64)
    key = ijk[:, 0] + NX * ijk[:, 1] + NX * NY * ijk[:, 2]
    assert np.array_equal(np.unique(key), keys)
    for v, k in enumerate(keys[:300]):
        labs = labels[key == k]
        vals, cnt = np.unique(labs, return_counts=True)
        assert cnt[list(vals).index(sub_labels[v])] == cnt.max()


@pytest.mark.parametrize("name", ["g010", "g004"])
def test_grid_golden(cuda, grid_golden, name):
    c = grid_golden[name]
    pts, feats, labels = grid_golden["points"], grid_golden["features"], grid_golden["labels"]
    sp, sf, sl = F.grid_sub_sampling(pts, features=feats, labels=labels, grid_size=float(c["dl"]))
    assert sp.dtype == np.float32 and sf.dtype == np.float32 and sl.dtype == np.int32
    assert sl.shape == (len(sp), 1)                         # wrapper.cpp:240-243: classes come back [M, ld]
    p, f, l = lexsorted(sp, sf, sl[:, 0])
    assert np.array_equal(p, c["sub_points"])               # bit-exact barycentres
    assert np.array_equal(f, c["sub_features"])             # bit-exact mean features
    # against the oracle: identical incl. the tie rule (smallest label), rows by ascending key
    op, of, ol, keys = O.grid_sub_sampling(pts, feats, labels, float(c["dl"]))
    assert np.array_equal(sp, op) and np.array_equal(sf, of) and np.array_equal(sl, ol)
    label_is_a_mode(pts, labels, float(c["dl"]), sp, sl[:, 0], keys)


def test_grid_variants(cuda, grid_golden):
    pts, feats, labels = grid_golden["points"], grid_golden["features"], grid_golden["labels"]
    only = F.grid_sub_sampling(pts, grid_size=0.1)
    assert isinstance(only, np.ndarray)
    assert np.array_equal(lexsorted(only)[0], grid_golden["g010"]["points_only"])
    sp, sf = F.DataProcessing.grid_sub_sampling(pts, features=feats, grid_size=0.1)
    assert np.array_equal(sp, only)
    sp2, sl2 = F.grid_sub_sampling(pts, labels=np.stack([labels, labels[::-1]], 1), grid_size=0.1)
    assert sl2.shape == (len(sp2), 2)
    op, ol, _ = O.grid_sub_sampling(pts, None, np.stack([labels, labels[::-1]], 1), 0.1)
    assert np.array_equal(sl2, ol)


@pytest.mark.parametrize("seed,n,dl", [(0, 1, 0.1), (1, 17, 0.05), (2, 200000, 0.02), (3, 5000, 10.0)])
def test_grid_vs_oracle_random(cuda, seed, n, dl):
    rs = np.random.RandomState(seed)
    pts = (rs.randn(n, 3) * 0.7).astype(np.float32)
    feats = rs.rand(n, 3).astype(np.float32)
    labels = rs.randint(-3, 9, (n,)).astype(np.int32)
    sp, sf, sl = F.grid_sub_sampling(pts, features=feats, labels=labels, grid_size=dl)
    op, of, ol, _ = O.grid_sub_sampling(pts, feats, labels, dl)
    assert np.array_equal(sp, op) and np.array_equal(sf, of) and np.array_equal(sl, ol)
    # idempotence-like property: one point per occupied voxel; counts add up
    assert len(sp) <= n
