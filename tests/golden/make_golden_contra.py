"""Golden fixture of loss_contra (tests/golden/contra.npz) produced by EXECUTING THE REFERENCE'S OWN
`Criterion.loss_contra` (models/utils/criterion.py:39-182, imported unmodified through refimport.py; `self` only
supplies cfg.mask_shape).  Build-container only.

    python tests/golden/make_golden_contra.py

Stored: every input and, per scene, the ORIGINAL indices of the masks the reference pooled (recovered by matching the
returned mask planes), plus the loss.
"""
import os
import sys
import types

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))

import refimport  # noqa: E402

R = refimport.load()


def make_inputs(seed, b, k, h, w, c, n_lo, n_hi):
    g = torch.Generator().manual_seed(seed)
    pm = torch.randn(b, k, h // 2, w // 2, generator=g)
    pm = torch.nn.functional.avg_pool2d(pm, 5, stride=1, padding=2) * 6 - torch.rand(b, k, 1, 1, generator=g) * 2
    xs, ys, gts, ff, f3 = [], [], [], [], []
    for _ in range(b):
        n = int(torch.randint(n_lo, n_hi, (1,), generator=g))
        x, y = torch.randint(0, h, (n,), generator=g), torch.randint(0, w, (n,), generator=g)
        split = int(torch.randint(w // 4, 3 * w // 4, (1,), generator=g))
        gt = (y >= split).float()
        gt[torch.rand(n, generator=g) < 0.1] = 2.0
        xs.append(x); ys.append(y); gts.append(gt)
        ff.append(torch.randn(n, c, generator=g)); f3.append(torch.randn(n, c, generator=g))
    return pm, xs, ys, gts, ff, f3, torch.randn(b, k, c, generator=g), torch.randn(b, k, c, generator=g)


def main():
    h, w, c = 48, 64, 32
    out = {}
    for tag, (seed, b, k, n_lo, n_hi) in {"a": (5, 4, 20, 1500, 4000), "b": (6, 3, 40, 10, 40), "c": (7, 2, 9, 3000, 3001)}.items():
        pm, xs, ys, gts, ff, f3, me, ce = make_inputs(seed, b, k, h, w, c, n_lo, n_hi)
        self = types.SimpleNamespace(cfg=types.SimpleNamespace(mask_shape=[h, w]))
        outputs = {"pred_masks": pm, "mask_embed": me, "mask_embed_clip": ce, "fused_pred_feature": ff, "pure3d_pred_feature": f3}
        loss, fm = R.Criterion.loss_contra(self, xs, ys, None, None, gts, outputs)
        up = torch.nn.functional.interpolate(pm, size=(h, w), mode="bilinear", align_corners=False)
        sel = -np.ones((b, 5), np.int64)
        for s, planes in fm:
            for j, pl in enumerate(planes):
                hit = [m for m in range(k) if torch.equal(up[s, m], pl)]
                assert len(hit) == 1
                sel[s, j] = hit[0]
        print(tag, "loss", float(loss["loss_3d_contra"]), "selected", sel.tolist())
        off = np.concatenate([[0], np.cumsum([len(x) for x in xs])])
        out.update({f"{tag}_pred_masks": pm.numpy(), f"{tag}_x": torch.cat(xs).numpy(), f"{tag}_y": torch.cat(ys).numpy(),
                    f"{tag}_gt": torch.cat(gts).numpy(), f"{tag}_off": off, f"{tag}_f3": torch.cat(f3).numpy(),
                    f"{tag}_clip": ce.numpy(), f"{tag}_me": me.numpy(), f"{tag}_sel": sel,
                    f"{tag}_loss": loss["loss_3d_contra"].numpy()})
    path = os.path.join(HERE, "contra.npz")
    np.savez_compressed(path, **out)
    print(f"contra: {os.path.getsize(path) / 1024:.0f} KiB")


if __name__ == "__main__":
    main()
