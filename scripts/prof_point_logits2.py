"""Profiling target: per-point logits, plain argmax, 19 and 200 classes on 1 M rows x 768."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from xmask3d_b200 import ops
dev = torch.device("cuda", 0)
g = torch.Generator(device=dev).manual_seed(1)
n, c = 1_000_000, 768
feat = torch.randn(n, c, device=dev, generator=g)
for t in (19, 200):
    te = torch.randn(t, c, device=dev, generator=g)
    for _ in range(2):
        ops.point_logits(feat, te, 1 / 0.07, want_logits=False)
torch.cuda.synchronize()
print("done")
