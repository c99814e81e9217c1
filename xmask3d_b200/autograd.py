"""Autograd wrappers so the fused ops drop into the reference's TRAINING loop as well.

The reference back-propagates through `cal_pred_logits` (models/xmask3d.py:230-235: gradients to
mask_embed, text_embed, null_embed and the learnable logit_scale) and through `mask_mapper`
(models/utils/fuser.py:22-34: gradients to mask_embed through the scatter-mean).  Forward passes run
on libxm3d; the backward of the scatter-mean is the segmented pooling kernel (the adjoint of a
scatter is a gather-sum), the backward of the logits is expressed with torch matmuls (cuBLAS) —
it is not on the inference hot path.
"""
from __future__ import annotations

from typing import List, Sequence

import torch

from . import ops


class _ScatterMean(torch.autograd.Function):
    """out[i] = sum_{m ∋ i} emb[seg(i), m] / counter_i   (counter 0 -> 1e-5)."""

    @staticmethod
    def forward(ctx, emb, seg_off, member, n_pts):
        out, counter = ops.scatter(emb.detach(), seg_off, n_pts, member=member)
        ctx.save_for_backward(seg_off, member, counter)
        ctx.k = emb.shape[1]
        ctx.mark_non_differentiable(counter)
        return out, counter

    @staticmethod
    def backward(ctx, grad_out, _grad_counter):
        seg_off, member, counter = ctx.saved_tensors
        # d emb[s,m] = sum over points i of segment s inside mask m of grad_out[i] / counter_i
        g = (grad_out / counter.unsqueeze(1)).contiguous()
        grad_emb, _, _ = ops.pool(g, seg_off, ctx.k, member=member, want_mean=False)
        return grad_emb, None, None, None


def scatter_mean(emb: torch.Tensor, seg_off: torch.Tensor, member: torch.Tensor, n_pts: int):
    """Differentiable (w.r.t. emb) mask -> point scatter-mean.  Returns (out [n_pts,c], counter)."""
    return _ScatterMean.apply(emb, seg_off, member, int(n_pts))


class _PoolMean(torch.autograd.Function):
    """mean[s,m] = sum of feat[i] over the points i of segment s inside mask m / cnt[s,m]  (0 where cnt = 0) —
    `feature[mask].mean(0)` of models/utils/criterion.py:148-157 and `feature.mean(0)` of models/xmask3d.py:239-258."""

    @staticmethod
    def forward(ctx, feat, seg_off, member, k, cap_pairs):
        _, cnt, mean = ops.pool(feat.detach(), seg_off, k, member=member, cap_pairs=cap_pairs)
        ctx.save_for_backward(seg_off, member, cnt)
        ctx.n = feat.shape[0]
        ctx.mark_non_differentiable(cnt)
        return mean, cnt

    @staticmethod
    def backward(ctx, grad_mean, _grad_cnt):
        seg_off, member, cnt = ctx.saved_tensors
        # d feat[i] = sum over masks m containing i of grad_mean[s,m] / cnt[s,m]: the scatter kernel returns that sum
        # divided by the number of masks containing i (its `counter`, 1e-5 where none does and the sum is 0)
        g = (grad_mean / cnt.clamp(min=1).unsqueeze(-1).to(grad_mean.dtype)).contiguous()
        out, counter = ops.scatter(g, seg_off, ctx.n, member=member)
        return out * counter.unsqueeze(1), None, None, None, None


def pool_mean(feat: torch.Tensor, seg_off: torch.Tensor, member: torch.Tensor, k: int, cap_pairs=None):
    """Differentiable (w.r.t. feat) masked mean pooling.  Returns (mean [n_seg,k,c], cnt int32 [n_seg,k]).
    cap_pairs: host-known bound on the memberships (None: counted on the device, one host sync)."""
    return _PoolMean.apply(feat, seg_off, member, int(k), cap_pairs)


class _PredLogits(torch.autograd.Function):
    """cal_pred_logits with max / mean ensembling; backward in torch from the saved operands."""

    @staticmethod
    def forward(ctx, mask_embed, text_embed, null_embed, logit_scale, sizes, ensemble):
        scale = float(logit_scale.item()) if torch.is_tensor(logit_scale) else float(logit_scale)
        out = ops.logits(mask_embed.detach(), text_embed.detach(), null_embed.detach(), sizes, scale, ensemble=ensemble)
        ctx.save_for_backward(mask_embed, text_embed, null_embed,
                              logit_scale if torch.is_tensor(logit_scale) else torch.tensor(scale))
        ctx.sizes, ctx.ensemble = list(sizes), ensemble
        ctx.scale_is_tensor = torch.is_tensor(logit_scale)
        return out

    @staticmethod
    def backward(ctx, grad_out):
        me, te, ne, ls = ctx.saved_tensors
        with torch.enable_grad():
            me_, te_, ne_ = me.detach().requires_grad_(True), te.detach().requires_grad_(True), ne.detach().requires_grad_(True)
            ls_ = ls.detach().to(me.device).requires_grad_(True)
            ref = _torch_pred_logits(me_, te_, ne_, ls_, ctx.sizes, ctx.ensemble)
            gm, gt, gn, gs = torch.autograd.grad(ref, (me_, te_, ne_, ls_), grad_out)
        return gm, gt, gn, (gs.reshape(ls.shape).to(ls.device) if ctx.scale_is_tensor else None), None, None


def _torch_pred_logits(me, te, ne, scale, sizes: Sequence[int], ensemble: str):
    import torch.nn.functional as F
    m = F.normalize(me, dim=-1)
    pred = scale * (m @ F.normalize(te, dim=-1).t())
    pieces = torch.split(pred, list(sizes), dim=-1)
    cols = [p.max(dim=-1).values if ensemble == "max" else p.mean(dim=-1) for p in pieces]
    null = scale * (m @ F.normalize(ne, dim=-1).t())
    return torch.cat([torch.stack(cols, dim=-1), null], dim=-1)


def pred_logits(mask_embed, text_embed, null_embed, logit_scale, sizes: List[int], ensemble: str = "max"):
    """Differentiable XMASK3d.cal_pred_logits (models/xmask3d.py:129-143)."""
    return _PredLogits.apply(mask_embed, text_embed, null_embed, logit_scale, sizes, ensemble)
