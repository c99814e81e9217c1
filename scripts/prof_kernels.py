"""Profiling target (run under ncu): one launch each of the projection, the mask-level logits
(configs[1] and configs[2] shapes) and the per-point logits kernels on bench-sized inputs."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import bench
from xmask3d_b200 import ops
from xmask3d_b200.pipeline import CorrespondencePipeline

sys.argv = [sys.argv[0]] + [a for a in sys.argv[1:]]
args = bench.parse()
dev = torch.device("cuda", 0)
batch, scenes = bench.build_batch(args, 0)
pipe = CorrespondencePipeline(batch, args.k, args.c, dev)
pipe.upload(torch.from_numpy(batch.xyz).pin_memory(), torch.from_numpy(batch.depth_mm.view(np.int16)).pin_memory())
for _ in range(2):
    pr = pipe.project()
torch.cuda.synchronize()
g = torch.Generator(device=dev).manual_seed(1)
for rows, t in ((160 * 50, 20), (160 * 100, 201)):
    me = torch.randn(rows, 768, device=dev, generator=g)
    te = torch.randn(t - 1, 768, device=dev, generator=g)
    ne = torch.randn(1, 768, device=dev, generator=g)
    for _ in range(2):
        ops.logits(me, te, ne, [1] * (t - 1), 1 / 0.07)
torch.cuda.synchronize()
n = 1_000_000
feat = torch.randn(n, 768, device=dev, generator=g)
for t in (19, 200):
    te = torch.randn(t, 768, device=dev, generator=g)
    for _ in range(2):
        ops.point_logits(feat, te, 1 / 0.07, want_logits=False)
torch.cuda.synchronize()
print("prof_kernels done")
