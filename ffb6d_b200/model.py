"""``FFB6DFusionNet``: everything of ``FFB6D.forward`` (models/ffb6d.py:203-337) except the ResNet/PSPNet
image backbone, on this package's kernels and with the reference's module names.

The reference interleaves two branches: a CNN producing ``rgb_emb0`` at every stage (cuDNN, out of this
repository's scope) and RandLA-Net on the points, tied together after every stage by the
bidirectional fusion layers.  This module owns the RandLA branch (``rndla_pre_stages``, ``rndla_ds_stages``,
``rndla_up_stages``), all 28 fusion layers (``{ds,up}_fuse_{r2p,p2r}_{pre,fuse}_layers``) and the three
prediction heads (``rgbd_seg_layer``, ``kp_ofst_layer``, ``ctr_ofst_layer``) under the reference's attribute
names, so the matching part of a published FFB6D checkpoint loads with ``load_state_dict(..., strict=False)``.
The image branch enters through ``rgb_feats``: the eight tensors the CNN stages would produce
(``cnn_ds_stages[i](...)`` x4, ``cnn_up_stages[i](...)`` x3, and the final full-resolution map); with
``rgb_stage_fn`` a caller can instead run its own CNN stage on the fused map, exactly where the reference does.

Used by ``tools/train_bench.py`` for BASELINE configs 3-4 (forward + backward under DDP, synthetic data).
"""
import torch
import torch.nn as nn

from . import ops
from . import tables as T
from .modules import Conv1d, Conv2d, Dilated_res_block, RandLAConv1d, RandLAConv2d


def _head(in_c, out_c):
    """``pt_utils.Seq(in_c).conv1d(128, bn).conv1d(128, bn).conv1d(128, bn).conv1d(out_c, activation=None)``
    (models/ffb6d.py:133-155); children are named "0".."3" like ``Seq`` names them."""
    seq = nn.Sequential()
    c = in_c
    for i in range(3):
        seq.add_module(str(i), Conv1d(c, 128, bn=True, activation=nn.ReLU()))
        c = 128
    seq.add_module("3", Conv1d(c, out_c, activation=None))
    return seq


class FFB6DFusionNet(nn.Module):
    def __init__(self, n_classes=2, n_pts=12288, n_kps=8, in_c=9, d_out=(32, 64, 128, 256)):
        super().__init__()
        self.n_cls, self.n_pts, self.n_kps = n_classes, n_pts, n_kps
        # RandLA branch (models/RandLA/RandLANet.py:12-38)
        self.rndla_pre_stages = RandLAConv1d(in_c, 8, kernel_size=1, bn=True)
        self.rndla_ds_stages = nn.ModuleList()
        d_in = 8
        for d in d_out:
            self.rndla_ds_stages.append(Dilated_res_block(d_in, d))
            d_in = 2 * d
        self.rndla_up_stages = nn.ModuleList()
        d_o = d_in
        for j in range(len(d_out)):
            if j < 3:
                d_i = d_o + 2 * d_out[-j - 2]
                d_o = 2 * d_out[-j - 2]
            else:
                d_i = 4 * d_out[-4]
                d_o = 2 * d_out[-4]
            self.rndla_up_stages.append(RandLAConv2d(d_i, d_o, kernel_size=(1, 1), bn=True))
        # fusion layers (models/ffb6d.py:49-80, 89-129)
        self.ds_rgb_oc, self.ds_rndla_oc = list(T.DS_RGB_OC), [2 * d for d in d_out]
        self.up_rgb_oc = list(T.UP_RGB_OC)
        self.up_rndla_oc = [self.ds_rndla_oc[-j - 2] if j < 3 else self.ds_rndla_oc[0] for j in range(len(d_out))]
        for tag, rgb_oc, rnd_oc, n in (("ds", self.ds_rgb_oc, self.ds_rndla_oc, 4), ("up", self.up_rgb_oc, self.up_rndla_oc, 3)):
            r2p_pre, r2p_fuse, p2r_pre, p2r_fuse = nn.ModuleList(), nn.ModuleList(), nn.ModuleList(), nn.ModuleList()
            for i in range(n):
                r2p_pre.append(Conv2d(rgb_oc[i], rnd_oc[i], kernel_size=(1, 1), bn=True))
                r2p_fuse.append(Conv2d(rnd_oc[i] * 2, rnd_oc[i], kernel_size=(1, 1), bn=True))
                p2r_pre.append(Conv2d(rnd_oc[i], rgb_oc[i], kernel_size=(1, 1), bn=True))
                p2r_fuse.append(Conv2d(rgb_oc[i] * 2, rgb_oc[i], kernel_size=(1, 1), bn=True))
            setattr(self, tag + "_fuse_r2p_pre_layers", r2p_pre)
            setattr(self, tag + "_fuse_r2p_fuse_layers", r2p_fuse)
            setattr(self, tag + "_fuse_p2r_pre_layers", p2r_pre)
            setattr(self, tag + "_fuse_p2r_fuse_layers", p2r_fuse)
        c = self.up_rndla_oc[-1] + self.up_rgb_oc[-1]
        self.rgbd_seg_layer = _head(c, n_classes)
        self.ctr_ofst_layer = _head(c, 3)
        self.kp_ofst_layer = _head(c, n_kps * 3)

    def _fuse(self, tag, i, rgb_emb0, p_emb0, p2r_idx, r2p_idx):
        """One bidirectional fusion (models/ffb6d.py:245-263 / 281-298), reference order of operations."""
        bs, c, hr, wr = rgb_emb0.shape
        p2r = getattr(self, tag + "_fuse_p2r_pre_layers")[i](p_emb0)
        p2r = ops.nearest_interpolation(p2r, p2r_idx).view(bs, -1, hr, wr)
        rgb_emb = getattr(self, tag + "_fuse_p2r_fuse_layers")[i](rgb_emb0, p2r)          # concat fused in the layer
        r2p = ops.random_sample(rgb_emb0.reshape(bs, c, hr * wr, 1), r2p_idx)
        r2p = getattr(self, tag + "_fuse_r2p_pre_layers")[i](r2p)
        p_emb = getattr(self, tag + "_fuse_r2p_fuse_layers")[i](p_emb0, r2p)
        return rgb_emb, p_emb

    def forward(self, inputs, rgb_feats=None, rgb_stage_fn=None):
        """``inputs``: the reference's dict (``cld_rgb_nrm [B,9,N]``, ``cld_xyz{i}``, the index tensors of
        :func:`ffb6d_b200.schedule.build_ffb6d_indices`, ``choose [B,1,N]``).  ``rgb_feats``: the 8 CNN stage
        outputs, or ``rgb_stage_fn(stage_index, fused_rgb_emb_or_None) -> rgb_emb0`` to compute them on the fly.
        Returns the reference's ``end_points`` keys ``pred_rgbd_segs``, ``pred_kp_ofs``, ``pred_ctr_ofs`` plus
        ``fused_rgb`` (the fused image maps of the seven stages, inputs of the caller's next CNN stage)."""
        def rgb_stage(j, prev):
            return rgb_stage_fn(j, prev) if rgb_stage_fn is not None else rgb_feats[j]

        p_emb = self.rndla_pre_stages(inputs["cld_rgb_nrm"]).unsqueeze(3)            # [B,8,N,1]
        rgb_emb = None
        ds_emb, fused_rgb = [], []
        for i in range(4):
            rgb_emb0 = rgb_stage(i, rgb_emb)
            f_enc = self.rndla_ds_stages[i](p_emb, inputs["cld_xyz%d" % i], inputs["cld_nei_idx%d" % i])
            p_emb0 = ops.random_sample(f_enc, inputs["cld_sub_idx%d" % i])
            if i == 0:
                ds_emb.append(f_enc)
            rgb_emb, p_emb = self._fuse("ds", i, rgb_emb0, p_emb0, inputs["p2r_ds_nei_idx%d" % i],
                                        inputs["r2p_ds_nei_idx%d" % i])
            ds_emb.append(p_emb)
            fused_rgb.append(rgb_emb)
        n_up = len(self.rndla_up_stages)
        for i in range(n_up - 1):
            rgb_emb0 = rgb_stage(4 + i, rgb_emb)
            f_interp = ops.nearest_interpolation(p_emb, inputs["cld_interp_idx%d" % (n_up - i - 1)])
            p_emb0 = self.rndla_up_stages[i](ds_emb[-i - 2], f_interp)                   # cat fused in the layer
            rgb_emb, p_emb = self._fuse("up", i, rgb_emb0, p_emb0, inputs["p2r_up_nei_idx%d" % i],
                                        inputs["r2p_up_nei_idx%d" % i])
            fused_rgb.append(rgb_emb)
        rgb_emb = rgb_stage(7, rgb_emb)
        f_interp = ops.nearest_interpolation(p_emb, inputs["cld_interp_idx0"])
        p_emb = self.rndla_up_stages[n_up - 1](ds_emb[0], f_interp).squeeze(-1)        # [B,64,N]
        bs, di = rgb_emb.shape[0], rgb_emb.shape[1]
        rgb_emb_c = ops.choose_gather(rgb_emb, inputs["choose"])                      # [B,64,N]
        # heads: the first layer takes the concat as two inputs (no materialised cat)
        outs = []
        for head in (self.rgbd_seg_layer, self.kp_ofst_layer, self.ctr_ofst_layer):
            x = head[0](rgb_emb_c, p_emb)
            for layer in list(head)[1:]:
                x = layer(x)
            outs.append(x)
        rgbd_segs, pred_kp_ofs, pred_ctr_ofs = outs
        pred_kp_ofs = pred_kp_ofs.view(bs, self.n_kps, 3, -1).permute(0, 1, 3, 2).contiguous()
        pred_ctr_ofs = pred_ctr_ofs.view(bs, 1, 3, -1).permute(0, 1, 3, 2).contiguous()
        # `fused_rgb`: the seven fused image maps the reference hands to its next CNN stage (models/ffb6d.py:251, 287)
        return {"pred_rgbd_segs": rgbd_segs, "pred_kp_ofs": pred_kp_ofs, "pred_ctr_ofs": pred_ctr_ofs,
                "fused_rgb": fused_rgb}

    @staticmethod
    def rgb_feature_shapes(batch, h=480, w=640):
        """Shapes of the 8 ``rgb_feats`` (models/ffb6d.py:31-43, 84-87)."""
        shapes = [(batch, T.DS_RGB_OC[i], h // T.RGB_DS_SR[i], w // T.RGB_DS_SR[i]) for i in range(4)]
        shapes += [(batch, T.UP_RGB_OC[i], h // T.RGB_UP_SR[i], w // T.RGB_UP_SR[i]) for i in range(3)]
        shapes.append((batch, T.UP_RGB_OC[2], h, w))
        return shapes
