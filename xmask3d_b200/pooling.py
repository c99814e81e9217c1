"""Masked mean pooling of per-point 3D features under 2D masks — the inline code of the
reference at models/utils/criterion.py:83-94 (gather / threshold), :148-157
(`feature_3d[mask_3d[k]].mean(0)`), and models/xmask3d.py:356-367 (score pooling), as one
batched call on libxm3d.
"""
from __future__ import annotations

import torch

from . import ops
from .fuser import masks_at_points


def masked_mean_pool(feats, x_list, y_list, masks, mode: str = "sigmoid_ge0.5"):
    """feats: list of [n_s, C] float32 CUDA tensors (or one concatenated tensor); masks: [B,k,h,w].
    Returns (mean [B,k,C], cnt int32 [B,k]).  Rows of empty masks are 0 (torch.mean gives NaN)."""
    member, _, seg_d, _ = masks_at_points(x_list, y_list, masks, mode)
    f = feats if torch.is_tensor(feats) else torch.cat(list(feats), 0)
    k = masks.shape[1] if torch.is_tensor(masks) else masks[0].shape[0]
    _, cnt, mean = ops.pool(f, seg_d, k, member=member)
    return mean, cnt


def masked_score_pool(scores, x_list, y_list, masks, mode: str = "sigmoid_gt0.5"):
    """models/xmask3d.py:356-367: (sum(score * mask) / (sum(mask) + 1e-10) [B,k], keep = sum(mask) > 0)."""
    member, _, seg_d, _ = masks_at_points(x_list, y_list, masks, mode)
    s = scores if torch.is_tensor(scores) else torch.cat(list(scores), 0)
    k = masks.shape[1] if torch.is_tensor(masks) else masks[0].shape[0]
    tot, cnt, _ = ops.pool(s.reshape(-1, 1), seg_d, k, member=member, want_mean=False)
    return tot[..., 0] / (cnt.to(torch.float32) + 1e-10), cnt > 0


def label_mean_pool(feats, labels, seg_off, k):
    """Partition masks (one label per point, -1 = none): (mean [B,k,C], cnt [B,k])."""
    _, cnt, mean = ops.pool(feats, seg_off, k, label=labels)
    return mean, cnt
