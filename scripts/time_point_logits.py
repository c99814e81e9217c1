"""Time the per-point logits kernel: plain argmax and the fused stream's ensemble epilogue."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from xmask3d_b200 import ops
dev = torch.device("cuda", 0)
g = torch.Generator(device=dev).manual_seed(1)
n, c = 2_339_470, 768
feat = torch.randn(n, c, device=dev, generator=g)
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
for t in (19, 200):
    te = torch.randn(t, c, device=dev, generator=g)
    kf = 40
    kw = dict(binary=(torch.rand(n, device=dev, generator=g) > 0.5).float(), is_base=torch.arange(t, device=dev) < (3 * t) // 4,
              mask_label=torch.randint(-1, kf, (n,), device=dev, generator=g, dtype=torch.int32),
              mask_probs=torch.rand(kf, t, device=dev, generator=g).softmax(-1), base_ratio=0.65, novel_ratio=0.35)
    for name, args in (("plain argmax", {}), ("ensemble + blend + argmax", kw)):
        for _ in range(2):
            ops.point_logits(feat, te, 1 / 0.07, want_logits=False, **args)
        torch.cuda.synchronize()
        e0.record()
        for _ in range(5):
            ops.point_logits(feat, te, 1 / 0.07, want_logits=False, **args)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 5
        print(f"T={t:3d} {name:28s}: {ms:.3f} ms = {4.0 * c * n / ms / 1e6:.0f} GB/s of feature reads")
