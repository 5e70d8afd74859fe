"""CPU: the pose-voting oracle (oracle/pose_oracle.py) against the outputs of the reference's own
MeanShiftTorch.fit / best_fit_transform (tests/golden/pose_cases.npz, made by make_golden.py)."""
import os

import numpy as np
import pytest

from conftest import GOLDEN
from oracle import pose_oracle as PO

CASES = ["small", "mid", "wide", "capped", "single"]


@pytest.fixture(scope="module")
def pose_golden():
    return np.load(os.path.join(GOLDEN, "pose_cases.npz"))


@pytest.mark.parametrize("case", CASES)
def test_mean_shift_oracle_matches_reference(pose_golden, case):
    votes = pose_golden[case + "_votes"]
    bw, max_iter = pose_golden[case + "_params"]
    for g in range(votes.shape[0]):
        c, lab, it = PO.mean_shift_fit(votes[g], bw, int(max_iter))
        assert np.abs(c - pose_golden[case + "_centres"][g]).max() < 2e-4      # metres: see oracle/pose_oracle.py on ties
        want = pose_golden[case + "_labels"][g].astype(bool)
        assert (lab != want).sum() <= max(1, votes.shape[1] // 200)            # points on the bandwidth boundary
        if case == "capped":
            assert it == int(max_iter) + 1                                      # `it > max_iter` stops AFTER max_iter + 1


def test_best_fit_oracle_matches_reference(pose_golden):
    A, B, T = pose_golden["fit_A"], pose_golden["fit_B"], pose_golden["fit_T"]
    for k in range(A.shape[0]):
        got = PO.best_fit_transform(A[k], B[k])
        assert np.abs(got - T[k]).max() < 1e-9
        R = got[:, :3]
        assert abs(np.linalg.det(R) - 1.0) < 1e-9 and np.abs(R @ R.T - np.eye(3)).max() < 1e-9
