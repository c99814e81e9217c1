"""Bring-up diagnostics of the tensor-core pooling kernel (pool_mma.cu): structured inputs whose wrong
placement is readable, then random inputs against a float64 torch contraction, then timing."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from xmask3d_b200 import ops

dev = torch.device("cuda", 0)
torch.manual_seed(0)


def pack(member_bool):            # [k, n] bool -> [n, words] int32
    k, n = member_bool.shape
    words = (k + 31) // 32
    out = torch.zeros(n, words, dtype=torch.int64, device=member_bool.device)
    for m in range(k):
        out[:, m // 32] |= member_bool[m].long() << (m % 32)
    out = torch.where(out >= 2 ** 31, out - 2 ** 32, out)
    return out.to(torch.int32)


TUNE = int(os.environ.get("PM_TUNE", "0"), 0)


def run(feat, seg, k, member, path):
    cap = feat.shape[0]
    return ops.pool(feat, seg, k, member=member, cap=cap, cap_pairs=cap * 16 + 2, path=path, _tune=(TUNE if path == "mma" else 0))


def report(tag, got, ref):
    err = (got.double() - ref).abs()
    scale = ref.abs().amax(-1).clamp_min(1e-30)
    rel = (err.amax(-1) / scale).max().item()
    print(f"{tag}: max abs err {err.max().item():.3e}  vector-rel {rel:.3e}  ref max {ref.abs().max().item():.3e}", flush=True)
    return rel


# 1. structured: n = 64, k = 64, c = 128; point p in mask p only; feat[p, c] = p + 1 at c == p
n, k, c = 64, 64, 128
feat = torch.zeros(n, c, device=dev)
feat[torch.arange(n), torch.arange(n)] = torch.arange(1, n + 1, device=dev).float()
mb = torch.eye(k, n, dtype=torch.bool, device=dev)
seg = torch.tensor([0, n], dtype=torch.int64, device=dev)
s, cnt, mean = run(feat, seg, k, pack(mb), "mma")
torch.cuda.synchronize()
ref = mb.double() @ feat.double()
r = report("structured one-hot", s[0], ref)
if r > 1e-6:
    nz = s[0].nonzero()
    print("nonzero (mask, channel, value) first 40:", [(int(a), int(b), float(s[0][a, b])) for a, b in nz[:40]])
print("cnt", cnt[0][:8].tolist(), "...", "ok" if torch.equal(cnt[0].long(), mb.sum(1)) else "COUNT MISMATCH")

# 2. structured, dense channels: feat[p, c] = p * 128 + c (exact in tf32 up to 2^11? no: use small values)
feat = (torch.arange(n, device=dev).view(n, 1) % 8 * 16 + torch.arange(c, device=dev).view(1, c) % 16).float()
s, cnt, mean = run(feat, seg, k, pack(mb), "mma")
report("structured dense (exact small ints)", s[0], mb.double() @ feat.double())

# 3. random features (hi + lo), random members, several tiles, ragged
for (ns, k, c, dens) in (([64], 64, 128, 0.3), ([200], 50, 128, 0.2), ([3000, 0, 1777, 1, 5200, 320], 50, 768, 0.12),
                         ([3000, 0, 1777, 1, 5200, 320], 96, 128, 0.12), ([3000, 64, 65], 33, 256, 0.5),
                         ([70000], 50, 256, 0.3), ([5000, 4000], 100, 384, 0.1), ([1000], 1, 128, 1.0)):
    off = np.concatenate([[0], np.cumsum(ns)]).astype(np.int64)
    total = int(off[-1])
    feat = torch.randn(total, c, device=dev)
    mbs = [torch.rand(k, n_, device=dev) < dens for n_ in ns]
    member = torch.cat([pack(m) for m in mbs])
    seg = torch.from_numpy(off).to(dev)
    got = {}
    for path in ("mma", "rows" if k <= 96 else "pair_lists"):
        s, cnt, mean = run(feat, seg, k, member, path)
        got[path] = s
        worst = 0.0
        for i, n_ in enumerate(ns):
            ref = mbs[i].double() @ feat[off[i]:off[i + 1]].double()
            err = (s[i].double() - ref).abs().amax(-1) / ref.abs().amax(-1).clamp_min(1e-30)
            worst = max(worst, err.max().item())
            assert torch.equal(cnt[i].long(), mbs[i].sum(1)), (path, i)
        print(f"ns={ns} k={k} c={c} path={path}: worst vector-rel err {worst:.3e}", flush=True)
    s2, _, _ = run(feat, seg, k, member, "mma")
    print("   deterministic:", torch.equal(s2, got["mma"]))

# 4. neighbours: huge values in the rows after a segment must not leak (the converters zero them)
ns = [100, 100]
feat = torch.randn(200, 128, device=dev)
feat[100:] = 1e30
mbs = [torch.rand(50, 100, device=dev) < 0.3 for _ in ns]
s, cnt, mean = run(feat, torch.tensor([0, 100, 200], device=dev), 50, torch.cat([pack(m) for m in mbs]), "mma")
report("segment followed by 1e30 rows", s[0], mbs[0].double() @ feat[:100].double())

# 5. timing at bench size: 2.34 M rows x 768, 160 segments, K = 50, ~7.7 memberships per point
total, c, k, nseg = 2_339_470, 768, 50, 160
feat = torch.randn(total, c, device=dev)
bounds = np.sort(np.random.default_rng(0).choice(np.arange(1, total), nseg - 1, replace=False))
off = np.concatenate([[0], bounds, [total]]).astype(np.int64)
seg = torch.from_numpy(off).to(dev)
bits = (torch.rand(total, 64, device=dev) < 7.7 / 50)
bits[:, 50:] = False
w0 = (bits[:, :32].long() << torch.arange(32, device=dev)).sum(1)
w1 = (bits[:, 32:].long() << torch.arange(32, device=dev)).sum(1)
member = torch.stack([w0, w1], 1)
member = torch.where(member >= 2 ** 31, member - 2 ** 32, member).to(torch.int32)
pairs = int(bits.sum().item())
def timeit(path, tune=0):
    for _ in range(2):
        out = ops.pool(feat, seg, k, member=member, cap=total, cap_pairs=pairs + 2, path=path, _tune=tune)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5):
        out = ops.pool(feat, seg, k, member=member, cap=total, cap_pairs=pairs + 2, path=path, _tune=tune)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 5
    print(f"bench-size pooling path={path} tune={tune:#x}: {ms:.3f} ms = {total * c * 4 / ms / 1e6:.0f} GB/s of feature reads ({pairs / total:.2f} memberships/point)", flush=True)
    return out[0]


ref_s = timeit("mma")
for tn in (3, 2, 0x100):
    timeit("mma", tn)
o = timeit("rows")
d = (o - ref_s).abs().amax(-1) / o.abs().amax(-1).clamp_min(1e-30)
print(f"   mma vs rows: worst vector-rel diff {d.max().item():.3e}")
print("dbg_pool_mma done")
