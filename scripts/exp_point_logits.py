"""Experiment: per-point logits kernel (features read once) at the bench size, ring depth forced with
XM3D_PL_STAGES (2 = two CTAs per SM, >= 3 = one CTA per SM with a deeper raw-tile ring)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from xmask3d_b200 import ops
dev = torch.device("cuda", 0)
rows, c = int(os.environ.get("ROWS", 2339470)), 768
feat = torch.empty((rows, c), dtype=torch.float32, device=dev).normal_()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
for t in (19, 200):
    te = torch.randn(t, c, device=dev)
    ref = None
    for st in os.environ.get("STAGES", "0,2,3,4,5").split(","):
        if int(st):
            os.environ["XM3D_PL_STAGES"] = st
        else:
            os.environ.pop("XM3D_PL_STAGES", None)
        try:
            for _ in range(2):
                out = ops.point_logits(feat, te, 1 / 0.07, want_logits=False)
            torch.cuda.synchronize()
            e0.record()
            for _ in range(5):
                out = ops.point_logits(feat, te, 1 / 0.07, want_logits=False)
            e1.record()
            torch.cuda.synchronize()
        except Exception as e:
            print(t, st, "failed", str(e)[:100]); continue
        am = out[1] if isinstance(out, tuple) else out
        if ref is None:
            ref = am.clone()
        ms = e0.elapsed_time(e1) / 5
        print(f"classes {t:3d} stages {st}: {ms:.3f} ms, {4.0 * c * rows / ms / 1e6:.0f} GB/s of features, "
              f"same argmax {bool(torch.equal(am, ref))}")
