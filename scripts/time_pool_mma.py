"""Time the pooling kernels on bench-sized inputs: argv = "rows" | "pair_lists" | a tune word for the tensor-core kernel
(0 = default; low nibble = depth of the raw ring, bit 9 = no size ordering, bit 10 = M = 128 for every K; the "switch a role
off" bits 14-17 only act in the instrumented build: make -C xmask3d_b200/csrc BUILD=build_dbg OUT=../libxm3d_dbg.so
EXTRA=-DXM3D_PM_TIMING, XM3D_SO=xmask3d_b200/libxm3d_dbg.so).  Timed-alone numbers differ by ~4 % from box to box: compare
two libraries in ONE gpurun call."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from xmask3d_b200 import ops
dev = torch.device("cuda", 0)
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
ref = None
for arg in sys.argv[1:]:
    path, tune = ("mma", int(arg, 0)) if arg not in ("rows", "pair_lists") else (arg, 0)
    for _ in range(2):
        out = ops.pool(feat, seg, k, member=member, cap=total, cap_pairs=pairs + 2, path=path, _tune=tune)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10):
        out = ops.pool(feat, seg, k, member=member, cap=total, cap_pairs=pairs + 2, path=path, _tune=tune)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 10
    if ref is None:
        ref = out[0].clone()
    d = ((out[0] - ref).abs().amax(-1) / ref.abs().amax(-1).clamp_min(1e-30)).max().item()
    print(f"{arg:>10s}: {ms:.3f} ms = {total * c * 4 / ms / 1e6:.0f} GB/s of feature reads; max vector-rel diff vs first {d:.2e}", flush=True)
