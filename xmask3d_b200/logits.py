"""Drop-in for `XMASK3d.cal_pred_logits` (reference models/xmask3d.py:129-143; twin
models/modeling/meta_arch/odise.py:170-194) and `ensemble_logits_with_labels`
(models/modeling/meta_arch/helper.py:72-97) on libxm3d's tcgen05 kernel.
"""
from __future__ import annotations

from typing import List

import torch

from . import ops


def ensemble_logits_with_labels(logits: torch.Tensor, labels: List[List[str]], ensemble_method: str = "max"):
    """helper.py:72-97 — reduce every synonym group (consecutive columns) to one column.  Kept in
    torch: it only regroups an existing logits tensor (`cal_pred_logits` below fuses the same
    reduction into the GEMM epilogue)."""
    sizes = [len(group) for group in labels]
    assert logits.shape[-1] == sum(sizes), f"{logits.shape[-1]} != {sum(sizes)}"
    assert ensemble_method in ["mean", "max"]
    pieces = torch.split(logits, sizes, dim=-1)
    if ensemble_method == "max":
        cols = [p.max(dim=-1).values for p in pieces]
    else:
        cols = [p.mean(dim=-1) for p in pieces]
    return torch.stack(cols, dim=-1)


def cal_pred_logits(outputs, ensemble_method: str = "max", want_argmax: bool = False):
    """outputs: dict with mask_embed [B,K,C], text_embed [T-1,C], null_embed [1,C], labels
    (list of synonym lists), logit_scale (tensor or float).  Returns [B,K,T] float32."""
    scale = outputs["logit_scale"]
    sizes = [len(l) for l in outputs["labels"]]
    needs_grad = torch.is_grad_enabled() and any(
        torch.is_tensor(t) and t.requires_grad
        for t in (outputs["mask_embed"], outputs["text_embed"], outputs["null_embed"], scale))
    if needs_grad and not want_argmax:
        from .autograd import pred_logits
        return pred_logits(outputs["mask_embed"], outputs["text_embed"], outputs["null_embed"], scale, sizes,
                           ensemble_method)
    scale = float(scale.item()) if torch.is_tensor(scale) else float(scale)
    return ops.logits(outputs["mask_embed"], outputs["text_embed"], outputs["null_embed"], sizes, scale,
                      ensemble=ensemble_method, want_argmax=want_argmax)
