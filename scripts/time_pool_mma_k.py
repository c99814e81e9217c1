"""Time the tensor-core pooling kernel for several mask counts K (bench-sized rows, ~15 % membership density):
K <= 64 runs MMAs with M = 64 and the 16-lane operand stores, K > 64 with M = 128."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from xmask3d_b200 import ops
dev = torch.device("cuda", 0)
total, c, nseg = 2_339_470, 768, 160
feat = torch.randn(total, c, device=dev)
bounds = np.sort(np.random.default_rng(0).choice(np.arange(1, total), nseg - 1, replace=False))
seg = torch.from_numpy(np.concatenate([[0], bounds, [total]]).astype(np.int64)).to(dev)
for k in [int(a) for a in sys.argv[1:]] or [32, 50, 64, 100, 128]:
    words = (k + 31) // 32
    member = torch.zeros(total, words, dtype=torch.int32, device=dev)
    for w in range(words):
        nb = min(32, k - 32 * w)
        bits = (torch.rand(total, nb, device=dev) < 0.15)
        val = (bits.long() << torch.arange(nb, device=dev)).sum(1)
        member[:, w] = torch.where(val >= 2 ** 31, val - 2 ** 32, val).to(torch.int32)
    pairs = int(ops._popcount32(member).sum().item())
    for path in ("mma", "rows") if k <= 96 else ("mma",):
        for _ in range(2):
            out = ops.pool(feat, seg, k, member=member, cap=total, cap_pairs=pairs + 2, path=path)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(10):
            out = ops.pool(feat, seg, k, member=member, cap=total, cap_pairs=pairs + 2, path=path)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 10
        print(f"K={k:4d} {path:5s}: {ms:.3f} ms = {4.0 * c * total / ms / 1e6:.0f} GB/s of feature reads ({pairs / total:.1f} memberships per point)")
