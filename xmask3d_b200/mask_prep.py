"""Reference-shaped front end of the fused mask preparation.

The reference prepares the 2D masks with the same inline torch sequence in two places —
`XMASK3d.forward` (models/xmask3d.py:326-331, 391-435) and `Criterion.loss_contra`'s caller
(models/utils/criterion.py:239-244, 273-320):

    mask_pred_results = F.interpolate(mask_pred_results, size=mask_shape, mode="bilinear", align_corners=False)
    mask_pred = mask_pred_result.sigmoid();  keep = scores > thresh
    cur_prob_masks = cur_scores.view(-1, 1, 1) * cur_masks;  cur_mask_ids = cur_prob_masks.argmax(0)
    for k: mask = (cur_mask_ids == k) & (cur_masks[k] >= 0.5); keep it iff the areas are non-zero

`partition_masks` returns what that loop produces (the kept mask indices and the stacked bool masks)
without materialising any [K,H,W] float tensor; `partition_labels` returns the equivalent label
image, which the pooling / scatter kernels consume directly (`label=`).
"""
from __future__ import annotations

from typing import Optional, Sequence, Tuple

import torch

from . import ops


def partition_labels(mask_pred_lowres: torch.Tensor, scores: torch.Tensor, mask_shape: Sequence[int],
                     score_thresh: float = 0.0):
    """mask_pred_lowres [B,K,hs,ws] float32 logits, scores [B,K] (softmax max).  Returns
    (label int16 [B,H,W] with -1 = no mask, final_keep bool [B,K], areas int32 [B,K,3])."""
    keep = scores > score_thresh
    pm = ops.mask_prep(mask_pred_lowres, mask_shape, scores=scores, keep=keep, want_bits=False, want_partition=True)
    final_keep = pm.areas[..., 2] > 0          # mask_area > 0 and original_area > 0 and mask.sum() > 0
    return pm.label, final_keep, pm.areas


def partition_masks(mask_pred_lowres: torch.Tensor, scores: torch.Tensor, mask_shape: Sequence[int],
                    score_thresh: float = 0.0) -> Tuple[torch.Tensor, torch.Tensor]:
    """Single view: mask_pred_lowres [K,hs,ws], scores [K].  Returns (final_keep indices into the K
    input masks, bool masks [K',H,W]) — `final_keep` / `torch.stack(final_mask)` of the reference loop
    (indices refer to the unfiltered masks; the reference's are relative to `keep`)."""
    label, fk, _ = partition_labels(mask_pred_lowres[None], scores[None], mask_shape, score_thresh)
    idx = torch.nonzero(fk[0]).flatten()
    masks = label[0][None].to(torch.int64) == idx.view(-1, 1, 1)
    return idx, masks


def masks_at_points_lowres(mask_pred_lowres: torch.Tensor, mask_shape: Sequence[int], rowcol: torch.Tensor,
                           seg_off: torch.Tensor, mode: str = "sigmoid_gt0.5", cap: Optional[int] = None):
    """`F.interpolate(mask)[:, x_label, y_label].sigmoid() > 0.5` (models/xmask3d.py:326-331, 356-357)
    for a batch of views -> member words of every visible point."""
    pm = ops.mask_prep(mask_pred_lowres, mask_shape, mode=mode, want_bits=True)
    member, _ = ops.point_bits(pm, rowcol, seg_off, cap=cap)
    return member
