"""Drop-ins for the pooling / selection code of the reference's `models/utils/criterion.py` and the per-scene
means of `models/xmask3d.py`, on libxm3d.

`loss_contra` keeps the reference's signature (criterion.py:39-46) and return value `(loss dict, final_2d_mask)`;
bind it to the reference's class (`Criterion.loss_contra = xmask3d_b200.criterion.loss_contra`) or call it with
any object whose `.cfg.mask_shape` gives the up-sampling size.  The reference spends four `.item()` host syncs per
mask and scene inside its Python loops (:100-121); here the selection of all scenes is four kernel launches and the
pooled means of all selected masks one pooling call, with ONE host copy (the number of selected masks per scene,
which the ragged return value needs).
"""
from __future__ import annotations

from typing import List, Optional

import torch
import torch.nn.functional as F

from . import ops
from .fuser import masks_at_points


def contra_select(x_list, y_list, masks_up: torch.Tensor, binary_gts):
    """criterion.py:80-146 for every scene of the batch: (ContraSelection, member, seg_off device, seg_off host).
    masks_up [B,K,H,W] float32 logits up-sampled to cfg.mask_shape; binary_gts: list of [n_s] tensors."""
    member, _, seg_d, seg = masks_at_points(x_list, y_list, masks_up, "sigmoid_ge0.5")
    gt = torch.cat([g.reshape(-1) for g in binary_gts]).to(device=masks_up.device, dtype=torch.float32)
    k = int(masks_up.shape[1])
    return ops.contra_select(member, k, gt, seg_d, masks_up, cap=int(seg[-1])), member, seg_d, seg


def loss_contra(self, x_list, y_list, masks, clip_mask_embeddings, binary_gts, outputs, want_fused: bool = False):
    """`Criterion.loss_contra` (criterion.py:39-182).  As in the reference, `masks` and `clip_mask_embeddings` are
    taken from `outputs` (the arguments are shadowed at :47-49).  Gradients flow to `outputs["pure3d_pred_feature"]`
    through the pooled means.  `want_fused` additionally returns the pooled fused features the reference computes
    (:150-152) but never uses."""
    masks = outputs["pred_masks"]
    mask_embeds = outputs["mask_embed"]
    clip = outputs["mask_embed_clip"]
    features_3d = outputs["pure3d_pred_feature"]
    masks_up = F.interpolate(masks, size=tuple(self.cfg.mask_shape), mode="bilinear", align_corners=False)
    b, k = int(masks_up.shape[0]), int(masks_up.shape[1])
    csel, member, seg_d, seg = contra_select(x_list, y_list, masks_up, binary_gts)
    feat3 = features_3d if torch.is_tensor(features_3d) else torch.cat(list(features_3d), 0)
    from .autograd import pool_mean
    if torch.is_grad_enabled() and feat3.requires_grad:
        mean3, _ = pool_mean(feat3, seg_d, csel.sel_member, 5, cap_pairs=5 * int(seg[-1]) + 2)
    else:
        _, _, mean3 = ops.pool(feat3, seg_d, 5, member=csel.sel_member, cap_pairs=5 * int(seg[-1]) + 2)
    fused = None
    if want_fused:
        ff = outputs["fused_pred_feature"]
        ff = ff if torch.is_tensor(ff) else torch.cat(list(ff), 0)
        _, _, fused = ops.pool(ff, seg_d, 5, member=csel.sel_member, cap_pairs=5 * int(seg[-1]) + 2)
    # the one host copy: how many masks every scene pools, and which
    host = torch.cat([csel.n_sel.view(-1, 1), csel.sel], 1).cpu()
    n_sel, sel = host[:, 0], host[:, 1:]
    rows = [(s, j) for s in range(b) for j in range(int(n_sel[s]))]
    final_2d_mask = []
    if rows:
        si = torch.tensor([r[0] for r in rows], device=masks_up.device)
        ji = torch.tensor([r[1] for r in rows], device=masks_up.device)
        mi = torch.tensor([int(sel[s, j]) for s, j in rows], device=masks_up.device)
        embedding_3d = mean3[si, ji]
        embedding_gt = clip[si, mi].detach()
        loss_3d_contra = (1 - F.cosine_similarity(embedding_3d, embedding_gt)).mean()
        for s in range(b):
            if int(n_sel[s]):
                final_2d_mask.append((s, masks_up[s, sel[s, :int(n_sel[s])].to(masks_up.device).long()]))
        if fused is not None:
            fused = fused[si, ji]
    else:
        invalid_embed = torch.stack([mask_embeds[-1][0]])
        loss_3d_contra = (1 - F.cosine_similarity(invalid_embed, invalid_embed)).mean()
    loss = {"loss_3d_contra": loss_3d_contra}
    if want_fused:
        return loss, final_2d_mask, fused
    return loss, final_2d_mask


def scene_mean(features, seg_off: Optional[torch.Tensor] = None) -> torch.Tensor:
    """`torch.stack([f.mean(0) for f in features])` (models/xmask3d.py:239-258: the per-scene mean of the fused / 3D /
    2D features that feeds the caption losses) — the K = 1 all-ones case of the pooling kernel, one launch for the
    whole batch.  features: list of [n_s, C] tensors, or one concatenated tensor with seg_off int64 [B+1]."""
    if torch.is_tensor(features):
        feat, seg = features, seg_off
    else:
        feats: List[torch.Tensor] = list(features)
        feat = torch.cat(feats, 0)
        n = torch.tensor([0] + [int(f.shape[0]) for f in feats], dtype=torch.int64)
        seg = torch.cumsum(n, 0).to(feat.device)
    ones = torch.ones((feat.shape[0], 1), dtype=torch.int32, device=feat.device)      # every point in mask 0
    if torch.is_grad_enabled() and feat.requires_grad:
        from .autograd import pool_mean
        mean, _ = pool_mean(feat, seg, ones, 1, cap_pairs=int(feat.shape[0]))
    else:
        _, _, mean = ops.pool(feat, seg, 1, member=ones, cap_pairs=int(feat.shape[0]))
    return mean[:, 0]
