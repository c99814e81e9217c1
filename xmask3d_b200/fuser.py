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


def _stack_ragged_masks(masks, mode: str):
    """The reference's callers pass RAGGED per-scene lists (models/utils/criterion.py:262-340): K' bool masks
    [K',h,w] for a scene whose argmax partition kept K' masks, or K float32 zero masks when nothing was kept.
    Pad to the largest K with planes that never hit (0 for `m >= 0.5`, -inf for the sigmoid modes).  Mixed
    dtypes are brought to a common one: bool/uint8 stay uint8 only if every scene is bool/uint8; for the plain
    `>= 0.5` mode float masks are thresholded first (exactly the reference's comparison)."""
    if torch.is_tensor(masks):
        return masks, [int(masks.shape[1])] * int(masks.shape[0])
    masks = list(masks)
    ks = [int(m.shape[0]) for m in masks]
    kmax = max(max(ks), 1)
    all_int = all(m.dtype in (torch.bool, torch.uint8) for m in masks)
    same = all(m.shape == masks[0].shape and m.dtype == masks[0].dtype for m in masks)
    if same and ks[0] > 0:
        return torch.stack(masks), ks
    h, w = masks[0].shape[-2:]
    dev = masks[0].device
    if all_int or mode == "ge0.5":
        out = torch.zeros((len(masks), kmax, h, w), dtype=torch.uint8, device=dev)
        for i, m in enumerate(masks):
            if ks[i]:
                out[i, :ks[i]] = m.to(torch.uint8) if m.dtype in (torch.bool, torch.uint8) else (m >= 0.5).to(torch.uint8)
        return out, ks
    out = torch.full((len(masks), kmax, h, w), float("-inf"), dtype=torch.float32, device=dev)
    for i, m in enumerate(masks):
        if ks[i]:
            # bool planes in a float batch: True -> +inf (sigmoid = 1), False -> -inf (sigmoid = 0)
            out[i, :ks[i]] = m.to(torch.float32) if m.dtype.is_floating_point else \
                torch.where(m.bool(), float("inf"), float("-inf")).to(torch.float32)
    return out, ks


def masks_at_points(x_list, y_list, masks, mode: str = "ge0.5"):
    """mask[:, x_label, y_label] + threshold for every scene: returns (member [sum n, words] int32,
    counts int32 [B,k], seg_off int64 [B+1] CUDA, seg_off host).  masks: tensor [B,k,h,w] or a (possibly
    ragged, possibly mixed bool / float) list of [k_s,h,w] per scene; k = the largest k_s."""
    m, _ = _stack_ragged_masks(masks, mode)
    dev = m.device
    n = [int(x.shape[0]) for x in x_list]
    seg = torch.zeros(len(n) + 1, dtype=torch.int64)
    seg[1:] = torch.cumsum(torch.tensor(n, dtype=torch.int64), 0)
    rowcol = torch.stack([torch.cat([x.reshape(-1) for x in x_list]),
                          torch.cat([y.reshape(-1) for y in y_list])], 1).to(device=dev, dtype=torch.int32)
    seg_d = seg.to(dev)
    member, counts = ops.gather_masks(m, rowcol, seg_d, mode=mode, want_counts=True)
    return member, counts, seg_d, seg


def _stack_ragged_embeds(mask_embeds, kmax: int):
    if torch.is_tensor(mask_embeds):
        return mask_embeds
    embs = list(mask_embeds)
    if all(e.shape == embs[0].shape for e in embs) and embs[0].shape[0] == kmax:
        return torch.stack(embs)
    out = embs[0].new_zeros((len(embs), kmax, embs[0].shape[-1]))
    for i, e in enumerate(embs):
        if e.shape[0]:
            out[i, :e.shape[0]] = e           # differentiable w.r.t. every scene's embeddings
    return out


def mask_mapper(x_list, y_list, masks, mask_embeds, pred_3ds, fuser, fc1, fc2, cfg):
    dev = pred_3ds[0].device
    b = len(x_list)
    member, counts, seg_d, seg = masks_at_points(x_list, y_list, masks, "ge0.5")
    # fuser.py:19-20 — when no mask holds any point, point 0 is put into mask 0.  On the device, no host round
    # trip: all bits of such a scene are 0, so OR-ing bit 0 of its first point's word 0 is an integer add.
    total = int(seg[-1])
    if total:
        empty = ((counts.sum(1) == 0) & (seg_d[1:] > seg_d[:-1])).to(torch.int32)
        member[:, 0].index_add_(0, seg_d[:-1].clamp(max=total - 1), empty)
    emb = _stack_ragged_embeds(mask_embeds, int(counts.shape[1]))
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
