"""The per-frame KNN / gather schedule of FFB6D, and its one-call GPU index build.

The reference builds 22 neighbour-index arrays per frame on the CPU inside the
dataset (datasets/ycb/ycb_dataset.py:269-309 == datasets/linemod/linemod_dataset.py:313-353)
and consumes them in ``FFB6D.forward`` through 22 gathers plus the ``choose``
gather (models/ffb6d.py:231-312).  This module states both schedules as data and
runs the index build on the GPU for a whole batch: xyz goes in, the same dict
keys with the same shapes and dtypes (int32) come out, already on the device.
"""
import contextlib

import torch

from .ops import knn_search, KnnGrid, knn_uses_grid

# reference literals (ycb_dataset.py:269-271, 298)
RGB_DS_SR = (4, 8, 8, 8)
RGB_UP_SR = (4, 2, 2)
PCLD_SUB_S_R = (4, 4, 4, 4)
N_DS_LAYERS = 4
N_UP_LAYERS = 3
K_NEIGH = 16

# feature widths seen by the gathers (models/ffb6d.py:49-50, 89-95; common.py:26)
DS_RGB_OC = (64, 128, 512, 1024)
DS_RNDLA_OC = (64, 128, 256, 512)
UP_RGB_OC = (256, 64, 64)
UP_RNDLA_OC = (256, 128, 64, 64)


def knn_schedule(n_points=12288, h=480, w=640, k=K_NEIGH):
    """The 22 KNN calls of one frame as ``(key, support, query, K)`` where support /
    query name a point set: ``("cld", level)`` = first ``n_points / 4**level`` cloud
    points, ``("img", sr)`` = the stride-``sr`` image pyramid level (``h//sr * w//sr``
    points).  Order = the reference's call order."""
    calls = []
    for i in range(N_DS_LAYERS):
        sr = RGB_DS_SR[i]
        calls.append(("cld_nei_idx%d" % i, ("cld", i), ("cld", i), k))
        calls.append(("cld_interp_idx%d" % i, ("cld", i + 1), ("cld", i), 1))
        calls.append(("r2p_ds_nei_idx%d" % i, ("img", sr), ("cld", i + 1), k))
        calls.append(("p2r_ds_nei_idx%d" % i, ("cld", i + 1), ("img", sr), 1))
    for i in range(N_UP_LAYERS):
        sr = RGB_UP_SR[i]
        lvl = N_DS_LAYERS - i - 1
        calls.append(("r2p_up_nei_idx%d" % i, ("img", sr), ("cld", lvl), k))
        calls.append(("p2r_up_nei_idx%d" % i, ("cld", lvl), ("img", sr), 1))
    return calls


def set_size(name, n_points=12288, h=480, w=640):
    kind, a = name
    if kind == "cld":
        n = n_points
        for i in range(a):
            n //= PCLD_SUB_S_R[min(i, len(PCLD_SUB_S_R) - 1)]
        return n
    return (h // a) * (w // a)


def gather_schedule(n_points=12288, h=480, w=640):
    """The 23 gathers of ``FFB6D.forward`` as ``(op, index_key, C, S, Q, K)`` in call
    order (models/ffb6d.py:240-312).  ``op`` is ``"random_sample"``,
    ``"nearest_interpolation"`` or ``"choose"``."""
    N = [set_size(("cld", i), n_points) for i in range(5)]
    HW = {sr: set_size(("img", sr), n_points, h, w) for sr in (1, 2, 4, 8)}
    k = K_NEIGH
    g = []
    for i in range(N_DS_LAYERS):
        sr = RGB_DS_SR[i]
        g.append(("random_sample", "cld_sub_idx%d" % i, DS_RNDLA_OC[i], N[i], N[i + 1], k))
        g.append(("nearest_interpolation", "p2r_ds_nei_idx%d" % i, DS_RGB_OC[i], N[i + 1], HW[sr], 1))
        g.append(("random_sample", "r2p_ds_nei_idx%d" % i, DS_RGB_OC[i], HW[sr], N[i + 1], k))
    up_in = (DS_RNDLA_OC[3], UP_RNDLA_OC[0], UP_RNDLA_OC[1])  # width entering each interp
    for i in range(N_UP_LAYERS):
        sr = RGB_UP_SR[i]
        lvl = N_DS_LAYERS - i - 1
        g.append(("nearest_interpolation", "cld_interp_idx%d" % lvl, up_in[i], N[lvl + 1], N[lvl], 1))
        g.append(("nearest_interpolation", "p2r_up_nei_idx%d" % i, UP_RGB_OC[i], N[lvl], HW[sr], 1))
        g.append(("random_sample", "r2p_up_nei_idx%d" % i, UP_RGB_OC[i], HW[sr], N[lvl], k))
    g.append(("nearest_interpolation", "cld_interp_idx0", UP_RNDLA_OC[2], N[1], N[0], 1))
    g.append(("choose", "choose", UP_RGB_OC[2], HW[1], N[0], 1))
    return g


def fusion_mlp_schedule(n_points=12288, h=480, w=640):
    """The 28 fusion 1x1 MLPs of ``FFB6D.forward`` as ``(name, P, C1, C2, Co)``: positions, the two
    concatenated input widths (C2 = 0 for the ``*_pre`` layers) and the output width
    (models/ffb6d.py:55-80, 104-129)."""
    N = [set_size(("cld", i), n_points) for i in range(5)]
    HW = {sr: set_size(("img", sr), n_points, h, w) for sr in (1, 2, 4, 8)}
    layers = []
    for i in range(N_DS_LAYERS):
        cr, cp, n1, hw = DS_RGB_OC[i], DS_RNDLA_OC[i], N[i + 1], HW[RGB_DS_SR[i]]
        layers += [("ds%d_r2p_pre" % i, n1, cr, 0, cp), ("ds%d_r2p_fuse" % i, n1, cp, cp, cp),
                   ("ds%d_p2r_pre" % i, n1, cp, 0, cr), ("ds%d_p2r_fuse" % i, hw, cr, cr, cr)]
    for i in range(N_UP_LAYERS):
        cr, cp, n1, hw = UP_RGB_OC[i], UP_RNDLA_OC[i], N[N_DS_LAYERS - i - 1], HW[RGB_UP_SR[i]]
        layers += [("up%d_r2p_pre" % i, n1, cr, 0, cp), ("up%d_r2p_fuse" % i, n1, cp, cp, cp),
                   ("up%d_p2r_pre" % i, n1, cp, 0, cr), ("up%d_p2r_fuse" % i, hw, cr, cr, cr)]
    return layers


def knn_alg_bytes(S, Q, K):
    """Algorithmic HBM bytes of one KNN call: xyz in once, int32 idx out."""
    return 12 * S + 12 * Q + 4 * Q * K


def gather_alg_bytes(C, S, Q, K):
    """Algorithmic HBM bytes of one gather: touched rows once, int32 idx,
    output once."""
    return 4 * C * min(S, Q * K) + 4 * Q * K + 4 * C * Q


def frame_alg_bytes(n_points=12288, h=480, w=640, k=K_NEIGH):
    """(knn_bytes, gather_bytes) per frame; 8 239 296 + 160 186 368 at the defaults."""
    kb = sum(knn_alg_bytes(set_size(s, n_points, h, w), set_size(q, n_points, h, w), kk)
             for _, s, q, kk in knn_schedule(n_points, h, w, k))
    gb = 0
    for op, _, Cc, S, Q, K in gather_schedule(n_points, h, w):
        gb += gather_alg_bytes(Cc, S, Q, k if K == K_NEIGH else K)
    return kb, gb


def image_pyramid(dpt_xyz, levels=(1, 2, 4, 8)):
    """``dpt_xyz [B,H,W,3]`` -> ``{sr: [B, (H//sr)*(W//sr), 3]}`` for sr in ``levels``: the
    stride-``sr`` sub-grids of the organised cloud (ycb_dataset.py:253-267)."""
    B, H, W, _ = dpt_xyz.shape
    pyr = {}
    for sr in levels:
        nh, nw = H // sr, W // sr
        pyr[sr] = dpt_xyz[:, :nh * sr:sr, :nw * sr:sr, :].reshape(B, nh * nw, 3).contiguous()
    return pyr


def build_ffb6d_indices(cld, dpt_xyz=None, k=K_NEIGH, index_dtype=torch.int32, timer=None, streams=None,
                        pyramid=None, image_hw=None, events=None, priority=None):
    """All neighbour-index tensors of the FFB6D fusion stack for a batch, on the GPU.

    :param cld: ``[B, N0, 3]`` float32 CUDA, the sampled (already shuffled) cloud
    :param dpt_xyz: ``[B, H, W, 3]`` float32 CUDA, the organised cloud (zero rows at holes); or
      pass ``pyramid={2: [B,HW/4,3], 4: ..., 8: ...}`` + ``image_hw=(H, W)`` (what
      :func:`ffb6d_b200.ops.backproject` returns) and leave it None
    :param streams: optional list of side ``torch.cuda.Stream`` s to overlap the 22 searches on
    :param priority: optional ``{index key: float}``; with ``streams`` the searches (and the grid
      builds they need) are issued in descending priority instead of descending size, so that the
      searches whose consumers are expensive finish first
    :param events: optional dict; with ``streams`` it receives one ``torch.cuda.Event`` per index
      key, recorded on the side stream that produced it, and the function returns WITHOUT joining
      the side streams: the caller waits per key (``stream.wait_event``) and joins ``streams``
      itself (:class:`ffb6d_b200.pipeline.FusionPass` overlaps the gathers with the remaining searches)
    :param timer: optional object with ``start(name, alg_bytes)`` / ``stop()`` called around
      every KNN call (bench.py's per-op CUDA-event timer)
    :return: dict with the reference's keys (ycb_dataset.py:283-309), each with a leading
      batch dimension: ``cld_xyz{i}`` f32 ``[B,Ni,3]``; ``cld_nei_idx{i}`` ``[B,Ni,k]``;
      ``cld_sub_idx{i}`` ``[B,Ni/4,k]`` (the first Ni/4 rows of ``cld_nei_idx{i}``, :279);
      ``cld_interp_idx{i}`` ``[B,Ni,1]``; ``r2p_ds_nei_idx{i}`` ``[B,Ni/4,k]``;
      ``p2r_ds_nei_idx{i}`` ``[B,HW,1]``; ``r2p_up_nei_idx{i}``, ``p2r_up_nei_idx{i}``.
      Index tensors are ``index_dtype`` (int32 like the datasets; pass torch.int64 to skip
      the cast ``model_fn`` does, train_ycb.py:224-232).

    "Random sampling" is the reference's: the cloud was shuffled once, every level keeps the
    first quarter of the previous one (ycb_dataset.py:233-235, 278).
    """
    if cld.dim() != 3 or cld.shape[2] != 3:
        raise ValueError("expected cld [B,N,3]")
    cld = cld.contiguous().float()
    B, n0, _ = cld.shape
    used = sorted(set(RGB_DS_SR) | set(RGB_UP_SR))          # sr=1 is never searched
    if pyramid is not None:
        if image_hw is None:
            raise ValueError("image_hw=(H, W) is required with pyramid=")
        H, W = image_hw
        sets = {("img", sr): pyramid[sr] for sr in used}
    else:
        if dpt_xyz is None or dpt_xyz.dim() != 4 or cld.shape[0] != dpt_xyz.shape[0]:
            raise ValueError("expected dpt_xyz [B,H,W,3] (or pyramid=)")
        H, W = dpt_xyz.shape[1], dpt_xyz.shape[2]
        sets = {("img", sr): p for sr, p in image_pyramid(dpt_xyz.float(), used).items()}
    n = n0
    for i in range(N_DS_LAYERS + 1):
        sets[("cld", i)] = cld if i == 0 else cld[:, :n, :].contiguous()
        if i < N_DS_LAYERS:
            n //= PCLD_SUB_S_R[i]
    inputs = {}
    calls = knn_schedule(n0, H, W, k)
    # one grid per (point set, K class) shared by all the searches into it; supports whose
    # searches are all tiny keep the tiled scan
    groups = {}
    for key, s, q, kk in calls:
        groups.setdefault((s, kk), []).append((key, q))
    gridded = [g for g, members in groups.items()
               if any(knn_uses_grid(B, sets[g[0]].shape[1], sets[q].shape[1], g[1]) for _, q in members)]
    grids = {}
    main = torch.cuda.current_stream(cld.device)
    par = streams is not None and timer is None

    def fork():
        if par:
            for st in streams:
                st.wait_stream(main)

    def join():
        if par:
            for st in streams:
                main.wait_stream(st)

    def on(i):
        return torch.cuda.stream(streams[i % len(streams)]) if par else contextlib.nullcontext()

    fork()
    # a search waits only for ITS grid, a consumer only for ITS index tensor (events, not joins)
    built = {}
    if par and priority:
        order = sorted(calls, key=lambda c: -priority.get(c[0], 0.0))
        first_use = {}
        for pos, (key, s_, q_, kk_) in enumerate(order):
            first_use.setdefault((s_, kk_), pos)
        build_order = sorted(gridded, key=lambda g: first_use.get(g, len(order)))
    else:
        order = sorted(calls, key=lambda c: -(sets[c[2]].shape[1] * c[3])) if par else calls
        build_order = sorted(gridded, key=lambda g: -sets[g[0]].shape[1])
    for i, g in enumerate(build_order):
        with on(i):
            if timer is not None:
                timer.start("knn_build:%s%d:k%d" % (g[0][0], g[0][1], g[1]), 0)
            grids[g] = KnnGrid(sets[g[0]], g[1])
            if timer is not None:
                timer.stop()
            if par:
                built[g] = torch.cuda.Event()
                built[g].record()
    for i, (key, s, q, kk) in enumerate(order):
        sup, qry = sets[s], sets[q]
        with on(i):
            if timer is not None:
                timer.start("knn:" + key, knn_alg_bytes(sup.shape[1], qry.shape[1], kk) * B)
            if (s, kk) in grids:
                if par:
                    torch.cuda.current_stream(cld.device).wait_event(built[(s, kk)])
                inputs[key] = grids[(s, kk)].query(qry, kk, out_dtype=index_dtype)
            else:
                inputs[key] = knn_search(sup, qry, kk, out_dtype=index_dtype, algo=1)
            if timer is not None:
                timer.stop()
            done = [key]
            if key.startswith("cld_nei_idx"):   # cld_sub_idx_i = the first N_{i+1} rows (ycb_dataset.py:279)
                lvl = int(key[len("cld_nei_idx"):])
                n_sub = sets[("cld", lvl + 1)].shape[1]
                inputs["cld_sub_idx%d" % lvl] = inputs[key][:, :n_sub, :].contiguous()
                done.append("cld_sub_idx%d" % lvl)
            if par and events is not None:
                ev = torch.cuda.Event()
                ev.record()
                for name in done:
                    events[name] = ev
    if not (par and events is not None):
        join()
    for i in range(N_DS_LAYERS):
        inputs["cld_xyz%d" % i] = sets[("cld", i)]
    return inputs


def build_ffb6d_indices_native(cld, pyramid, image_hw, k=K_NEIGH, index_dtype=torch.int32):
    """:func:`build_ffb6d_indices` through the single C entry point ``ffb6d_build_indices`` (one call,
    one stream, caller-owned buffers): what a non-Python host of the library would run.  Same keys,
    shapes, dtypes and bits as :func:`build_ffb6d_indices`.

    :param cld: ``[B, N0, 3]`` float32 CUDA; :param pyramid: ``{2: [B,HW/4,3], 4: ..., 8: ...}``
    :param image_hw: ``(H, W)`` of the full-resolution image"""
    import ctypes
    from ._lib import lib, check
    from .ops import _stream
    if cld.dim() != 3 or cld.shape[2] != 3 or not cld.is_cuda:
        raise ValueError("expected cld [B,N,3] on a CUDA device")
    cld = cld.contiguous().float()
    B, n0, _ = cld.shape
    H, W = image_hw
    img = {sr: pyramid[sr].contiguous().float() for sr in (2, 4, 8)}
    for sr in (2, 4, 8):
        if tuple(img[sr].shape) != (B, (H // sr) * (W // sr), 3):
            raise ValueError("pyramid[%d] must be [B, %d, 3]" % (sr, (H // sr) * (W // sr)))
    calls = knn_schedule(n0, H, W, k)
    inputs, ptrs = {}, (ctypes.c_void_p * len(calls))()
    for j, (key, s, q, kk) in enumerate(calls):
        inputs[key] = torch.empty((B, set_size(q, n0, H, W), kk), dtype=index_dtype, device=cld.device)
        ptrs[j] = inputs[key].data_ptr()
    nbytes = int(lib.ffb6d_build_indices_workspace_bytes(B, n0, H, W, int(k)))
    ws = torch.empty(max(nbytes, 1), dtype=torch.uint8, device=cld.device)
    with torch.cuda.device(cld.device):
        check(lib.ffb6d_build_indices(cld.data_ptr(), img[2].data_ptr(), img[4].data_ptr(), img[8].data_ptr(), B, n0, H, W,
                                      int(k), ctypes.cast(ptrs, ctypes.c_void_p), int(index_dtype == torch.int64),
                                      ws.data_ptr(), nbytes, _stream(cld.device)))
    n = n0
    for i in range(N_DS_LAYERS):
        inputs["cld_xyz%d" % i] = cld if i == 0 else cld[:, :n, :].contiguous()
        n //= PCLD_SUB_S_R[i]
        inputs["cld_sub_idx%d" % i] = inputs["cld_nei_idx%d" % i][:, :n, :].contiguous()
    return inputs


def build_ffb6d_indices_from_depth(depth, K, choose, k=K_NEIGH, index_dtype=torch.int32, streams=None):
    """Depth map in, all index tensors out: back-projection, sampling and stride pyramids
    (datasets/ycb/ycb_dataset.py:165-176, 237, 253-267) followed by the 22 searches, everything on
    the GPU.  ``depth [B,H,W]`` float32 metres, ``K`` camera matrix, ``choose [B,1,N]`` pixel
    indices of the sampled points.  Returns the dict of :func:`build_ffb6d_indices`."""
    from .ops import backproject
    cld, pyr = backproject(depth, K, choose)
    return build_ffb6d_indices(cld, None, k=k, index_dtype=index_dtype, streams=streams, pyramid=pyr,
                               image_hw=(depth.shape[1], depth.shape[2]))
