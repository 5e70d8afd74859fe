"""Keypoint voting and pose fitting on the GPU: the step after the network.

Mirrors of the reference's evaluation helpers, same names and argument meaning:

* :class:`MeanShiftTorch` -- ``utils/meanshift_pytorch.py:27-57`` (``fit``; ``fit_multi_clus`` is only used by
  that file's own ``__main__`` and is not provided);
* :func:`best_fit_transform` -- ``utils/pvn3d_eval_utils_kpls.py:28-59`` (numpy in, numpy ``[3,4]`` out);
* :func:`cal_frame_poses`, :func:`cal_frame_poses_lm` -- ``utils/pvn3d_eval_utils_kpls.py:65-160, 220-284``.

The reference reads mesh keypoints, mesh centres and object radii from dataset files through its module-level
``bs_utils`` / ``config``; here the caller passes them (``mesh_kps``, ``cls_radius``).  Everything numeric runs in
:func:`ffb6d_b200.ops.mean_shift_fit` (one persistent kernel for all keypoints of an object instead of one
N x N torch program per keypoint and a host sync per iteration) and :func:`ffb6d_b200.ops.best_fit_transform`.
"""
import numpy as np
import torch

from . import ops


class MeanShiftTorch:
    def __init__(self, bandwidth=0.05, max_iter=300):
        self.bandwidth = bandwidth
        self.stop_thresh = bandwidth * 1e-3
        self.max_iter = max_iter

    def fit(self, A, ret_mid_res=False):
        """``A [N,3]`` CUDA float32 -> ``(centre [3], labels [N] bool)``; with ``ret_mid_res`` the converged
        positions ``[N,3]`` and their pairwise distances ``[N,N]`` like the reference (that matrix is built
        with torch, for callers that want it; ``fit`` itself never forms it)."""
        out = ops.mean_shift_fit(A[None].float(), None, self.bandwidth, self.max_iter, return_modes=ret_mid_res)
        if not ret_mid_res:
            return out[0][0], out[1][0]
        C = out[3][0]
        return C, torch.cdist(C, C)

    def fit_batch(self, votes, valid=None):
        """``votes [G,N,3]``, ``valid [N]`` or ``[G,N]`` -> ``(centres [G,3], labels [G,N] bool)``: ``fit`` of every
        ``votes[g][valid]`` in one launch (labels are in the unmasked point order)."""
        centres, labels, _ = ops.mean_shift_fit(votes, valid, self.bandwidth, self.max_iter)
        return centres, labels


def best_fit_transform(A, B):
    """numpy ``[M,3]`` x2 -> numpy float64 ``[3,4]`` (pvn3d_eval_utils_kpls.py:28-59), computed on the GPU."""
    dev = torch.device("cuda", torch.cuda.current_device())
    a = torch.as_tensor(np.asarray(A, dtype=np.float32), device=dev)
    b = torch.as_tensor(np.asarray(B, dtype=np.float32), device=dev)
    return ops.best_fit_transform(a, b).cpu().numpy()


def _vote_object(pred_ctr, pred_kp, cls_msk, radius, use_ctr, use_ctr_clus_flter):
    """Centre + keypoints of one object: the two clustering steps of pvn3d_eval_utils_kpls.py:122-137."""
    ms = MeanShiftTorch(bandwidth=radius)
    ctr, ctr_labels = ms.fit_batch(pred_ctr[None], cls_msk)
    kp_msk = ctr_labels[0] if use_ctr_clus_flter else cls_msk       # `cls_voted_kps[:, ctr_labels, :]`
    kps, _ = ms.fit_batch(pred_kp, kp_msk)
    return torch.cat((kps, ctr), 0) if use_ctr else kps


def cal_frame_poses(pcld, mask, ctr_of, pred_kp_of, use_ctr, n_cls, use_ctr_clus_flter, gt_kps=None, gt_ctrs=None,
                    debug=False, kp_type="farthest", *, mesh_kps, cls_radius=None):
    """YCB-style multi-object voting (pvn3d_eval_utils_kpls.py:65-160).

    ``pcld [N,3]``, ``mask [N]`` predicted class ids, ``ctr_of [1,N,3]``, ``pred_kp_of [n_kps,N,3]`` (CUDA).
    ``mesh_kps``: mapping ``cls_id -> [n_kps(+1),3]`` array of object-frame keypoints (the centre last when
    ``use_ctr``), what the reference loads with ``bs_utils.get_kps`` / ``get_ctr``; ``cls_radius``: sequence
    indexed by ``cls_id - 1`` (``config.ycb_r_lst``), needed for the centre-clustering mask filter.
    Returns ``(pred_cls_ids, pred_pose_lst, pred_kps_lst)`` like the reference."""
    n_kps, n_pts, _ = pred_kp_of.shape
    pred_ctr = pcld - ctr_of[0]
    pred_kp = pcld.view(1, n_pts, 3) - pred_kp_of
    radius = 0.04
    pred_cls_ids = np.unique(mask[mask > 0].contiguous().cpu().numpy())
    if use_ctr_clus_flter and len(pred_cls_ids) > 0 and cls_radius is not None:
        # refine the mask: a point takes the class of the nearest voted centre when it is close enough (:85-109)
        ms = MeanShiftTorch(bandwidth=radius)
        msks = torch.stack([mask == int(c) for c in pred_cls_ids])
        ctrs, _ = ms.fit_batch(pred_ctr[None].expand(len(pred_cls_ids), n_pts, 3).contiguous(), msks)
        min_dis, min_idx = torch.cdist(pred_ctr, ctrs).min(dim=1)
        ids_t = torch.as_tensor(pred_cls_ids.astype(np.int64), device=mask.device)
        closest = ids_t[min_idx].to(mask.dtype)
        r = torch.as_tensor(np.asarray(cls_radius, np.float32), device=mask.device)[(ids_t - 1).clamp(min=0)][min_idx]
        update = (mask > 0) & (min_dis < r * 0.8)
        mask = torch.where(update, closest, mask)
    pred_pose_lst, pred_kps_lst = [], []
    for cls_id in pred_cls_ids:
        if cls_id == 0:
            break
        cls_msk = mask == int(cls_id)
        if int(cls_msk.sum()) < 1:
            pred_pose_lst.append(np.identity(4)[:3, :])
            pred_kps_lst.append(np.zeros((n_kps + 1, 3)))
            continue
        kpc = _vote_object(pred_ctr, pred_kp, cls_msk, radius, use_ctr, use_ctr_clus_flter)
        mk = torch.as_tensor(np.asarray(mesh_kps[int(cls_id)], np.float32), device=kpc.device)
        pred_pose_lst.append(ops.best_fit_transform(mk, kpc).cpu().numpy())
        pred_kps_lst.append(kpc.cpu().numpy())
    return pred_cls_ids, pred_pose_lst, pred_kps_lst


def cal_frame_poses_lm(pcld, mask, ctr_of, pred_kp_of, use_ctr, n_cls, use_ctr_clus_flter, obj_id, debug=False, *,
                       mesh_kps):
    """LineMOD-style single-object voting (pvn3d_eval_utils_kpls.py:220-284): class 1 is the object.
    ``mesh_kps [n_kps(+1),3]``: object-frame keypoints of ``obj_id`` (centre last when ``use_ctr``)."""
    n_kps, n_pts, _ = pred_kp_of.shape
    pred_ctr = pcld - ctr_of[0]
    pred_kp = pcld.view(1, n_pts, 3) - pred_kp_of
    cls_msk = mask == 1
    if int(cls_msk.sum()) < 1:
        return [np.identity(4)[:3, :]]
    kpc = _vote_object(pred_ctr, pred_kp, cls_msk, 0.04, use_ctr, use_ctr_clus_flter)
    mk = torch.as_tensor(np.asarray(mesh_kps, np.float32), device=kpc.device)
    return [ops.best_fit_transform(mk, kpc).cpu().numpy()]
