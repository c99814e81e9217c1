"""Soak test of the two persistent tcgen05 kernels: random shapes for a fixed time, every result checked
(tensor-core pooling against the pair-list kernel, per-point logits against float64 torch)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import torch.nn.functional as F
from xmask3d_b200 import ops
dev = torch.device("cuda", 0)
budget = float(sys.argv[1]) if len(sys.argv) > 1 else 60.0
rng = np.random.default_rng(int(sys.argv[2]) if len(sys.argv) > 2 else 0)
t_end = time.time() + budget
it = 0
while time.time() < t_end:
    # ---- pooling
    n_seg = int(rng.integers(1, 400))
    k = int(rng.choice([1, 7, 20, 33, 50, 64, 65, 100, 128]))
    c = int(rng.choice([128, 256, 768, 1024]))
    max_n = int(rng.choice([10, 200, 3000, 30000]))
    n = rng.integers(0, max_n + 1, n_seg)
    n[rng.random(n_seg) < 0.1] = 0
    total = int(n.sum())
    if total == 0:
        continue
    if total * c > 600_000_000:
        continue
    off = torch.from_numpy(np.concatenate([[0], np.cumsum(n)]).astype(np.int64)).to(dev)
    g = torch.Generator(device=dev).manual_seed(int(rng.integers(1 << 30)))
    f = torch.randn(total, c, device=dev, generator=g)
    words = (k + 31) // 32
    mem = torch.randint(-2 ** 31, 2 ** 31 - 1, (total, words), device=dev, generator=g, dtype=torch.int64).to(torch.int32)
    if rng.random() < 0.5:
        mem &= torch.randint(-2 ** 31, 2 ** 31 - 1, (total, words), device=dev, generator=g, dtype=torch.int64).to(torch.int32)
    a_sum, a_cnt, a_mean = ops.pool(f, off, k, member=mem, cap_pairs=total * k, path="mma")
    b_sum, b_cnt, b_mean = ops.pool(f, off, k, member=mem, cap_pairs=total * k, path="pair_lists")
    assert torch.equal(a_cnt, b_cnt), ("pool counts", n_seg, k, c)
    err = float(((a_sum - b_sum).abs().amax(-1) / b_sum.abs().amax(-1).clamp_min(1e-30)).max())
    assert err < 1e-5, ("pool sums", err, n_seg, k, c, max_n)
    # ---- per-point logits
    rows = int(rng.choice([1, 100, 129, 5000, 70000, 300000]))
    t = int(rng.choice([3, 19, 64, 65, 150, 200, 255]))
    cc = int(rng.choice([64, 100, 768]))
    feat = torch.randn(rows, cc, device=dev, generator=g) * 2
    te = torch.randn(t, cc, device=dev, generator=g)
    out, amax = ops.point_logits(feat, te, 1 / 0.07)
    ref = (1 / 0.07) * (F.normalize(feat.double(), dim=-1) @ F.normalize(te.double(), dim=-1).t())
    # error of the cosine similarity itself (relative to max|logit| is meaningless when a tiny problem has only small logits)
    assert float((out.double() - ref).abs().max()) * 0.07 < 5e-6, ("logits", rows, t, cc)
    top2 = ref.topk(min(2, t), -1).values
    clear = (top2[:, 0] - top2[:, -1]) > 1e-4
    assert torch.equal(amax.long()[clear], ref.argmax(-1)[clear]), ("argmax", rows, t, cc)
    it += 1
torch.cuda.synchronize()
print(f"soak ok: {it} iterations in {budget:.0f} s")
