"""Profiling target: per-point logits, 200 classes, fused-stream ensemble epilogue (argv[1] = rows, default 600k)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from xmask3d_b200 import ops
dev = torch.device("cuda", 0)
g = torch.Generator(device=dev).manual_seed(1)
n, c, t, kf = int(sys.argv[1]) if len(sys.argv) > 1 else 600_000, 768, 200, 40
feat = torch.randn(n, c, device=dev, generator=g)
te = torch.randn(t, c, device=dev, generator=g)
kw = dict(binary=(torch.rand(n, device=dev, generator=g) > 0.5).float(), is_base=torch.arange(t, device=dev) < (3 * t) // 4,
          mask_label=torch.randint(-1, kf, (n,), device=dev, generator=g, dtype=torch.int32),
          mask_probs=torch.rand(kf, t, device=dev, generator=g).softmax(-1), base_ratio=0.65, novel_ratio=0.35)
for _ in range(3):
    ops.point_logits(feat, te, 1 / 0.07, want_logits=False, **kw)
torch.cuda.synchronize()
print("done")
