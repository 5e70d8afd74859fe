"""CPU: the oracle (oracle/cpu_oracle.py) against the committed golden vectors, which are
outputs of the reference's own code (tests/golden/make_golden.py)."""
import hashlib
import json
import os

import numpy as np
import pytest

from conftest import GOLDEN, frame_point_sets
from oracle import cpu_oracle as O

KNN_NAMES = ["self_768_k16", "interp_192_768_k1", "r2p_4800_192_k16", "p2r_192_4800_k1",
             "self_48_k16", "uniform_1000_500_k1", "uniform_1000_500_k8", "uniform_1000_500_k32",
             "k_gt_s_10_k16", "batch3_768_k16"]


@pytest.mark.parametrize("name", KNN_NAMES)
def test_knn_oracle_bit_exact(knn_golden, name):
    c = knn_golden[name]
    sup, qry = c["support"], c["query"]
    if sup.ndim == 2:
        sup, qry, want = sup[None], qry[None], c["idx"][None]
    else:
        want = c["idx"]
    got = O.knn_search(sup, qry, int(c["k"]))
    assert got.dtype == np.int32
    assert np.array_equal(got, want)


def test_knn_oracle_k_gt_s_trailing_zero(knn_golden):
    c = knn_golden["k_gt_s_10_k16"]
    assert (c["idx"][:, 10:] == 0).all()          # the reference's behaviour (NN/knn_.cxx:120-121)


def test_knn_oracle_ties_contract(knn_golden):
    """Duplicated support points: the reference's order among equal distances is a KD-tree
    artefact; the oracle must agree on the sorted distances of every row."""
    c = knn_golden["ties_256_k8"]
    sup, qry, want = c["support"][None], c["query"][None], c["idx"][None]
    got = O.knn_search(sup, qry, int(c["k"]))
    ok, ndiff, nexc, msg = O.knn_matches(sup, qry, got, want)
    assert ok, msg
    assert ndiff > 0, "fixture should contain tie rows that differ"
    # and knn_matches must reject a genuinely wrong answer
    bad = got.copy()
    bad[0, 0, -1] = (bad[0, 0, -1] + 97) % 256
    assert not O.knn_matches(sup, qry, bad, want)[0]


def test_schedule_digest_small_frame():
    """All 22 KNN calls of one synthetic frame (n_points=3072) against the reference digests."""
    from ffb6d_b200.schedule import knn_schedule
    from ffb6d_b200.synthetic import make_frame
    d = json.load(open(os.path.join(GOLDEN, "schedule_digest.json")))["frames"]["seed2_n3072"]
    fr = make_frame(d["seed"], n_points=d["n_points"])
    ps = frame_point_sets(fr, d["n_points"])
    for key, s, q, k in knn_schedule(d["n_points"]):
        idx = O.knn_search(ps[s][None], ps[q][None], k)[0]
        assert list(idx.shape) == d["keys"][key]["shape"], key
        assert hashlib.sha256(np.ascontiguousarray(idx).tobytes()).hexdigest() == \
            d["keys"][key]["sha256"], key


@pytest.mark.parametrize("name", ["rs_small", "rs_k8", "rs_wide"])
def test_random_sample_oracle(gather_golden, name):
    c = gather_golden[name]
    assert np.array_equal(O.random_sample(c["feat"], c["idx"]), c["out"])
    gf = O.gather_max_backward(c["feat"][..., 0], c["idx"], c["gout"][..., 0])
    np.testing.assert_allclose(gf, c["gfeat"][..., 0], rtol=1e-5, atol=1e-6)


@pytest.mark.parametrize("name", ["ni_small", "ni_wide"])
def test_nearest_interpolation_oracle(gather_golden, name):
    c = gather_golden[name]
    assert np.array_equal(O.nearest_interpolation(c["feat"], c["idx"]), c["out"])


@pytest.mark.parametrize("name", ["gn_xyz", "gn_feat", "gn_odd"])
def test_gather_neighbour_oracle(gather_golden, name):
    c = gather_golden[name]
    assert np.array_equal(O.gather_neighbour(c["pc"], c["idx"]), c["out"])


def test_relative_pos_encoding_oracle(gather_golden):
    c = gather_golden["rpe"]
    got = O.relative_pos_encoding(c["xyz"], c["idx"])
    assert np.array_equal(got[..., 1:], c["out"][..., 1:])        # selections / differences: exact
    # the norm: torch CPU's vectorised sqrt is not correctly rounded (<= 1 ulp off)
    np.testing.assert_allclose(got[..., 0], c["out"][..., 0], rtol=2e-7, atol=0)


def _lexsorted(p, *others):
    o = np.lexsort((p[:, 2], p[:, 1], p[:, 0]))
    return [p[o]] + [x[o] for x in others]


@pytest.mark.parametrize("name", ["g010", "g004"])
def test_grid_oracle(grid_golden, name):
    c = grid_golden[name]
    pts, feats, labels = grid_golden["points"], grid_golden["features"], grid_golden["labels"]
    sp, sf, sl, keys = O.grid_sub_sampling(pts, feats, labels, float(c["dl"]))
    assert (np.diff(keys.astype(np.int64)) > 0).all()
    p, f, l = _lexsorted(sp, sf, sl)
    assert np.array_equal(p, c["sub_points"])                      # bit-exact barycentres
    assert np.array_equal(f, c["sub_features"])                    # bit-exact mean features
    # labels: any label with the maximal count is a valid answer (hash-map order in the reference)
    assert l.shape == c["sub_labels"].shape
    assert (l == c["sub_labels"]).mean() > 0.5


def test_grid_oracle_points_only(grid_golden):
    (sp, keys) = O.grid_sub_sampling(grid_golden["points"], None, None, 0.1)
    (p,) = _lexsorted(sp)
    assert np.array_equal(p, grid_golden["g010"]["points_only"])
