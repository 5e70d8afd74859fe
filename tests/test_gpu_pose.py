"""GPU: keypoint voting (ffb6d_mean_shift_fit) and pose fitting (ffb6d_best_fit_transform) against the
reference's MeanShiftTorch.fit / best_fit_transform outputs (tests/golden/pose_cases.npz) and the oracle.

Mean shift is floating point with a data-dependent stop; the sums run in another order than torch's, so the
contract is a tolerance (oracle/pose_oracle.py explains the figure): centres within 2e-4 m of the reference's
(0.5 % of the bandwidth; the centre is one member of the winning collapsed cluster and the members tie), labels
equal except for points within 3e-4 m of the bandwidth boundary."""
import os

import numpy as np
import pytest
import torch

import ffb6d_b200 as F
from conftest import GOLDEN
from ffb6d_b200 import pose as P
from oracle import pose_oracle as PO

pytestmark = pytest.mark.gpu
CASES = ["small", "mid", "wide", "capped", "single"]


@pytest.fixture(scope="module")
def pose_golden():
    return np.load(os.path.join(GOLDEN, "pose_cases.npz"))


def _labels_ok(got, want, modes, centre, bw):
    bad = got != want
    if not bad.any():
        return True
    d = np.linalg.norm(modes[bad] - centre, axis=1)
    return bool((np.abs(d - bw) < 3e-4).all())


@pytest.mark.parametrize("case", CASES)
def test_mean_shift_matches_reference(cuda, pose_golden, case):
    votes = torch.from_numpy(pose_golden[case + "_votes"]).cuda()
    bw, max_iter = pose_golden[case + "_params"]
    centres, labels, iters, modes = F.mean_shift_fit(votes, None, bw, int(max_iter), return_modes=True)
    c = centres.cpu().numpy()
    assert np.abs(c - pose_golden[case + "_centres"]).max() < 2e-4
    for g in range(votes.shape[0]):
        assert _labels_ok(labels[g].cpu().numpy(), pose_golden[case + "_labels"][g].astype(bool), modes[g].cpu().numpy(),
                          c[g], bw), (case, g)
    it = iters.cpu().numpy()
    if case == "capped":
        assert (it == int(max_iter) + 1).all()
    else:
        assert (it >= 1).all() and (it <= int(max_iter) + 1).all()
    # the single-set mirror of the reference class gives the same answer as the batched call
    ms = P.MeanShiftTorch(bandwidth=bw, max_iter=int(max_iter))
    c0, l0 = ms.fit(votes[0])
    assert torch.equal(c0, centres[0]) and torch.equal(l0, labels[0])


def test_mean_shift_masks_and_empty_sets(cuda, pose_golden):
    votes = torch.from_numpy(pose_golden["mid_votes"]).cuda()
    G, N, _ = votes.shape
    g = torch.Generator().manual_seed(3)
    shared = (torch.rand(N, generator=g) < 0.6).cuda()
    per_set = (torch.rand(G, N, generator=g) < 0.5).cuda()
    per_set[1] = False                                             # a set nobody votes in
    c_sh, l_sh, it_sh = F.mean_shift_fit(votes, shared, 0.04, 300)
    c_ps, l_ps, it_ps = F.mean_shift_fit(votes, per_set, 0.04, 300)
    for k in range(G):
        want_c, want_l, _ = PO.mean_shift_fit(votes[k][shared].cpu().numpy(), 0.04, 300)      # votes[mask], like the reference
        assert np.abs(c_sh[k].cpu().numpy() - want_c).max() < 2e-4
        assert (l_sh[k][shared].cpu().numpy() != want_l).sum() <= 2 and not l_sh[k][~shared].any()
        if k == 1:
            assert it_ps[k].item() == 0 and not l_ps[k].any() and c_ps[k].abs().max().item() == 0
            continue
        want_c, want_l, _ = PO.mean_shift_fit(votes[k][per_set[k]].cpu().numpy(), 0.04, 300)
        assert np.abs(c_ps[k].cpu().numpy() - want_c).max() < 2e-4
        assert (l_ps[k][per_set[k]].cpu().numpy() != want_l).sum() <= 2 and not l_ps[k][~per_set[k]].any()
    # bit-reproducible run to run (fixed work decomposition, no float atomics)
    c2, l2, _ = F.mean_shift_fit(votes, shared, 0.04, 300)
    assert torch.equal(c2, c_sh) and torch.equal(l2, l_sh)


def test_mean_shift_full_size(cuda):
    """BASELINE size: 9 vote sets (8 keypoints + centre) of all 12288 points; property checks only (the
    N x N oracle does not fit): the centre is a fixed point of the update and lies in the dense cluster."""
    g = torch.Generator().manual_seed(0)
    G, N = 9, 12288
    truth = torch.rand(G, 1, 3, generator=g) * 0.3 + torch.tensor([0.0, 0.0, 0.8])
    votes = truth + torch.randn(G, N, 3, generator=g) * 0.01
    votes[:, ::7] += torch.rand(G, (N + 6) // 7, 3, generator=g) * 0.4 - 0.2       # outliers
    votes = votes.cuda()
    centres, labels, iters = F.mean_shift_fit(votes, None, 0.04, 300)
    assert (centres.cpu() - truth[:, 0]).norm(dim=1).max().item() < 2e-3
    assert (labels.float().mean(dim=1) > 0.8).all()
    assert (iters > 1).all() and (iters <= 301).all()
    # fixed point: one more Gaussian-weighted mean around the centre moves it by less than the stop threshold
    d2 = ((votes - centres[:, None]) ** 2).sum(2)
    w = torch.exp(-0.5 * d2 / 0.04 ** 2)
    # (weights over the ORIGINAL votes define the mode of the kernel density estimate only approximately for
    # blurring mean shift, so this is a loose sanity bound, not the convergence criterion)
    shift = ((w[..., None] * votes).sum(1) / w.sum(1, keepdim=True) - centres).norm(dim=1)
    assert shift.max().item() < 5e-3


def test_best_fit_transform_matches_reference(cuda, pose_golden):
    A = torch.from_numpy(pose_golden["fit_A"]).cuda()
    B = torch.from_numpy(pose_golden["fit_B"]).cuda()
    T = F.best_fit_transform(A, B).cpu().numpy()
    assert np.abs(T - pose_golden["fit_T"]).max() < 1e-9
    for k in range(T.shape[0]):
        R = T[k, :, :3]
        assert abs(np.linalg.det(R) - 1.0) < 1e-12 and np.abs(R @ R.T - np.eye(3)).max() < 1e-12
    one = P.best_fit_transform(pose_golden["fit_A"][2], pose_golden["fit_B"][2])       # numpy in / numpy out, like the reference
    assert one.shape == (3, 4) and np.abs(one - pose_golden["fit_T"][2]).max() < 1e-9


def test_cal_frame_poses_recovers_pose(cuda):
    """End to end on synthetic predictions: points of two objects vote with noisy offsets; the fitted poses
    must reproduce the transforms the votes were generated from (LineMOD and YCB entry points)."""
    g = np.random.RandomState(5)
    n_pts, n_kps = 4096, 8
    mesh = {1: g.uniform(-0.08, 0.08, (n_kps + 1, 3)), 2: g.uniform(-0.06, 0.06, (n_kps + 1, 3))}
    for k in mesh:
        mesh[k][n_kps] = 0.0                                        # the centre is the object origin
    poses = {}
    pcld = g.uniform(-0.3, 0.3, (n_pts, 3)).astype(np.float32) + np.array([0, 0, 1.0], np.float32)
    mask = np.zeros(n_pts, np.int64)
    mask[:1500], mask[1500:2600] = 1, 2
    kp_of = g.normal(0, 0.2, (n_kps, n_pts, 3)).astype(np.float32)       # background: noise
    ctr_of = g.normal(0, 0.2, (1, n_pts, 3)).astype(np.float32)
    for cls_id, sl in ((1, slice(0, 1500)), (2, slice(1500, 2600))):
        q, _ = np.linalg.qr(g.normal(size=(3, 3)))
        if np.linalg.det(q) < 0:
            q[:, 0] *= -1
        t = np.array([0.1 * cls_id - 0.15, 0.05, 1.0])
        poses[cls_id] = np.concatenate((q, t[:, None]), 1)
        kps_cam = mesh[cls_id] @ q.T + t
        n = sl.stop - sl.start
        for k in range(n_kps):
            kp_of[k, sl] = pcld[sl] - kps_cam[k] + g.normal(0, 0.004, (n, 3))
        ctr_of[0, sl] = pcld[sl] - kps_cam[n_kps] + g.normal(0, 0.004, (n, 3))
    # a few mislabelled points inside object 1's mask, voting nonsense
    kp_of[:, 100:160] = g.normal(0, 0.3, (n_kps, 60, 3))
    ctr_of[0, 100:160] = g.normal(0, 0.3, (60, 3))
    tp, tm, tc, tk = (torch.from_numpy(a).cuda() for a in (pcld, mask, ctr_of, kp_of))
    ids, pose_lst, kps_lst = P.cal_frame_poses(tp, tm, tc, tk, True, 3, True, mesh_kps=mesh, cls_radius=[0.15, 0.12])
    assert list(ids) == [1, 2] and len(pose_lst) == 2
    for cls_id, T, kps in zip(ids, pose_lst, kps_lst):
        assert np.abs(T - poses[int(cls_id)]).max() < 5e-3, cls_id
        assert kps.shape == (n_kps + 1, 3)
    lm = P.cal_frame_poses_lm(tp, tm, tc, tk, True, 2, True, 1, mesh_kps=mesh[1])
    assert len(lm) == 1 and np.abs(lm[0] - poses[1]).max() < 5e-3
    none = P.cal_frame_poses_lm(tp, torch.zeros_like(tm), tc, tk, True, 2, True, 1, mesh_kps=mesh[1])
    assert np.array_equal(none[0], np.identity(4)[:3, :])
