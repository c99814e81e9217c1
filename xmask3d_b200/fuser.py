"""Drop-in for `mask_mapper` of the reference's `models/utils/fuser.py` (:6-53; inline twin
models/xmask3d.py:405-473): mask-at-point gather, threshold, mask -> point scatter-mean, fusion.

The gather / threshold / scatter-mean run on libxm3d for all scenes of the batch in three
launches; the learned layers (`fuser`, `fc1`, `fc2`) stay the caller's torch modules.  The
scatter-mean output is bit-identical to the reference loop (same float32 op order).  When
mask_embed requires grad the scatter-mean goes through `autograd.scatter_mean`, whose backward is
the segmented pooling kernel (gradients w.r.t. mask_embed as in the reference's training loop).
"""
from __future__ import annotations

from typing import List

import torch

from . import ops


def masks_at_points(x_list, y_list, masks, mode: str = "ge0.5"):
    """mask[:, x_label, y_label] + threshold for every scene: returns (member [sum n, words] int32,
    counts int32 [B,k], seg_off int64 [B+1] CUDA).  masks: list / tensor of [k,h,w] per scene."""
    dev = masks[0].device
    n = [int(x.shape[0]) for x in x_list]
    seg = torch.zeros(len(n) + 1, dtype=torch.int64)
    seg[1:] = torch.cumsum(torch.tensor(n, dtype=torch.int64), 0)
    rowcol = torch.stack([torch.cat([x.reshape(-1) for x in x_list]),
                          torch.cat([y.reshape(-1) for y in y_list])], 1).to(device=dev, dtype=torch.int32)
    m = masks if torch.is_tensor(masks) else torch.stack(list(masks))
    seg_d = seg.to(dev)
    member, counts = ops.gather_masks(m, rowcol, seg_d, mode=mode, want_counts=True)
    return member, counts, seg_d, seg


def mask_mapper(x_list, y_list, masks, mask_embeds, pred_3ds, fuser, fc1, fc2, cfg):
    dev = pred_3ds[0].device
    b = len(x_list)
    member, counts, seg_d, seg = masks_at_points(x_list, y_list, masks, "ge0.5")
    # fuser.py:19-20 — when no mask holds any point, point 0 is put into mask 0
    empty = (counts.sum(1) == 0).cpu()
    for s in range(b):
        if bool(empty[s]) and seg[s + 1] > seg[s]:
            member[int(seg[s]), 0] |= 1
    emb = mask_embeds if torch.is_tensor(mask_embeds) else torch.stack(list(mask_embeds))
    total = int(seg[-1])
    if torch.is_grad_enabled() and emb.requires_grad:
        from .autograd import scatter_mean                  # training: gradients flow to mask_embed
        feat2d_all, counter_all = scatter_mean(emb, seg_d, member, total)
    else:
        feat2d_all, counter_all = ops.scatter(emb, seg_d, total, member=member)

    output, output_2d, output_3d, output_2d_pre = [], [], [], []
    for s in range(b):
        a, e = int(seg[s]), int(seg[s + 1])
        pred_3d = pred_3ds[s]
        mask_3d_feature = feat2d_all[a:e]
        covered = counter_all[a:e] >= 1
        single_2d = mask_3d_feature[covered]
        final = torch.zeros_like(pred_3d, device=dev)
        final[covered] = fuser(single_2d, pred_3d[covered])
        final[~covered] = pred_3d[~covered]
        output.append(final)
        output_2d.append(fc2(mask_3d_feature))
        output_3d.append(fc1(pred_3d))
        if cfg.caption_contra_2d_pre:
            output_2d_pre.append(single_2d)
    return output, output_2d, output_3d, output_2d_pre
