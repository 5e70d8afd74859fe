"""oracle/pose_oracle.py -- TEST INFRASTRUCTURE ONLY. NOT A PRODUCT PATH.

CPU restatement (numpy, float32 like the reference's torch tensors) of the keypoint-voting step:

* ``mean_shift_fit``      -- MeanShiftTorch.fit, ffb6d/utils/meanshift_pytorch.py:33-57
* ``best_fit_transform``  -- ffb6d/utils/pvn3d_eval_utils_kpls.py:28-59

Pinned against the reference's own code executed on the CPU (oracle/ref_loader.pose_functions, fixtures in
tests/golden/pose_cases.npz made by tests/golden/make_golden.py).  Floating point: torch's reduction order is not
restated, so the pin is by tolerance, not bitwise.  The returned centre is ONE member of the winning collapsed
cluster -- the one with the most neighbours, and nearly all members tie on that count -- so two correct
implementations may return different members: when the iteration stops (max shift < bandwidth * 1e-3 = 4e-5 m, or the
iteration cap, which the outliers of realistic vote sets usually run into) the members of a cluster agree to about
1e-4 m.  The contract is therefore |centre - reference centre| < 2e-4 m (0.5 % of the 0.04 m bandwidth).
"""
import numpy as np


def mean_shift_fit(A, bandwidth=0.05, max_iter=300, return_modes=False):
    """``A [N,3]`` -> ``(centre [3] f32, labels [N] bool, iterations)`` (+ converged positions)."""
    A = np.asarray(A, np.float32)
    N = A.shape[0]
    bw = np.float32(bandwidth)
    stop = np.float32(bandwidth * 1e-3)                      # :31
    norm = np.float32(bandwidth * np.sqrt(2 * np.pi))
    C = A.copy()
    it = 0
    while True:                                              # :38-49
        it += 1
        diff = C[None, :, :] - C[:, None, :]
        dis = np.sqrt((diff * diff).sum(2, dtype=np.float32))
        w = np.exp(np.float32(-0.5) * (dis / bw) ** 2) / norm      # gaussian_kernel, :18-20
        new_C = (w[:, :, None] * C[None, :, :]).sum(1, dtype=np.float32) / w.sum(1, dtype=np.float32)[:, None]
        shift = np.sqrt(((new_C - C) ** 2).sum(1, dtype=np.float32))
        C = new_C.astype(np.float32)
        if shift.max() < stop or it > max_iter:
            break
    diff = C[:, None, :] - C[None, :, :]                     # :51-55
    dis = np.sqrt((diff * diff).sum(2, dtype=np.float32))
    num_in = (dis < bw).sum(1)
    max_idx = int(np.argmax(num_in))                         # first maximum, like torch.max on the CPU
    labels = dis[max_idx] < bw
    if return_modes:
        return C[max_idx].copy(), labels, it, C
    return C[max_idx].copy(), labels, it


def best_fit_transform(A, B):
    """``A, B [M,3]`` -> ``[3,4]`` float64 ``[R|t]`` mapping A onto B."""
    A = np.asarray(A, np.float64)
    B = np.asarray(B, np.float64)
    ca, cb = A.mean(0), B.mean(0)
    H = (A - ca).T @ (B - cb)                                # :47
    U, S, Vt = np.linalg.svd(H)
    R = Vt.T @ U.T
    if np.linalg.det(R) < 0:                                 # :52-54
        Vt[2, :] *= -1
        R = Vt.T @ U.T
    T = np.zeros((3, 4))
    T[:, :3] = R
    T[:, 3] = cb - R @ ca
    return T
