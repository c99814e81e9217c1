"""Profiling target: the tensor-core pooling kernel on bench-sized inputs (2.34 M rows x 768, 160 segments, K = 50,
7.7 memberships per point).  argv[1] = tune word (0 = default)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from xmask3d_b200 import ops
dev = torch.device("cuda", 0)
tune = int(sys.argv[1], 0) if len(sys.argv) > 1 else 0
path = sys.argv[2] if len(sys.argv) > 2 else "mma"
total, c, k, nseg = 2_339_470, 768, 50, 160
feat = torch.randn(total, c, device=dev)
bounds = np.sort(np.random.default_rng(0).choice(np.arange(1, total), nseg - 1, replace=False))
seg = torch.from_numpy(np.concatenate([[0], bounds, [total]]).astype(np.int64)).to(dev)
bits = (torch.rand(total, 64, device=dev) < 7.7 / 50)
bits[:, 50:] = False
w0 = (bits[:, :32].long() << torch.arange(32, device=dev)).sum(1)
w1 = (bits[:, 32:].long() << torch.arange(32, device=dev)).sum(1)
member = torch.stack([w0, w1], 1)
member = torch.where(member >= 2 ** 31, member - 2 ** 32, member).to(torch.int32)
pairs = int(bits.sum().item())
for _ in range(3):
    ops.pool(feat, seg, k, member=member, cap=total, cap_pairs=pairs + 2, path=path, _tune=tune)
torch.cuda.synchronize()
print("prof_pool_mma done")
