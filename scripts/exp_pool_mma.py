"""Experiment: per-role wait / work cycles of pool_mma_kernel (needs the -DXM3D_PM_TIMING build:
make -C xmask3d_b200/csrc BUILD=build_dbg OUT=../libxm3d_dbg.so EXTRA=-DXM3D_PM_TIMING; XM3D_SO=xmask3d_b200/libxm3d_dbg.so)."""
import os, sys, ctypes as C
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from xmask3d_b200 import ops

dev = torch.device("cuda", 0)
lib = ops.L.lib()
lib.xm3d_pool_mma_debug.restype = C.c_int
lib.xm3d_pool_mma_debug.argtypes = [C.c_void_p]
total, c, k, nseg = 2_339_470, 768, int(os.environ.get("PM_K", "50")), 160
feat = torch.randn(total, c, device=dev)
bounds = np.sort(np.random.default_rng(0).choice(np.arange(1, total), nseg - 1, replace=False))
off = np.concatenate([[0], bounds, [total]]).astype(np.int64)
seg = torch.from_numpy(off).to(dev)
words = (k + 31) // 32
member = torch.zeros(total, words, dtype=torch.int32, device=dev)
for w in range(words):
    nb = min(32, k - 32 * w)
    bits = (torch.rand(total, nb, device=dev) < 0.154)
    val = (bits.long() << torch.arange(nb, device=dev)).sum(1)
    member[:, w] = torch.where(val >= 2 ** 31, val - 2 ** 32, val).to(torch.int32)
pairs = int(ops._popcount32(member).sum().item())
tunes = [int(x, 0) for x in sys.argv[1:]] or [0]
for tune in tunes:
    for _ in range(2):
        ops.pool(feat, seg, k, member=member, cap=total, cap_pairs=pairs + 2, path="mma", _tune=tune)
    torch.cuda.synchronize()
    dbg = torch.zeros(148 * 8 * 4, dtype=torch.int64, device=dev)
    assert lib.xm3d_pool_mma_debug(dbg.data_ptr()) == 0
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    ops.pool(feat, seg, k, member=member, cap=total, cap_pairs=pairs + 2, path="mma", _tune=tune)
    e1.record()
    torch.cuda.synchronize()
    assert lib.xm3d_pool_mma_debug(None) == 0
    d = dbg.cpu().numpy().reshape(148, 8, 4).astype(np.float64)
    tiles = d[:, 0, 3]
    print(f"tune {tune:#x}: {e0.elapsed_time(e1):.3f} ms; tiles per CTA {tiles.mean():.0f} (min {tiles.min():.0f} max {tiles.max():.0f})")
    names = {0: ("TMA producer", ["wait raw_empty", "-", "loop total", "tiles"]),
             1: ("MMA issuer", ["wait raw_full", "wait conv_full", "wait tmem_free", "loop total"]),
             2: ("converter", ["wait raw_full", "wait conv_empty", "loop total", "fence.proxy.async"]),
             3: ("builder", ["wait conv_empty", "expand + tcgen05.st", "loop total", "tcgen05.wait::st"]),
             5: ("builder (2)", ["wait transposed words", "-", "-", "-"]),
             6: ("preparation", ["cp.async.wait_group", "LDS + transposes", "issue cp.async", "wait free slot"]),
             4: ("epilogue", ["wait tile_done", "-", "loop total", "-"])}
    for r, (nm, cols) in names.items():
        per_tile = d[:, r, :].sum(0) / tiles.sum()
        print(f"  {nm:14s} " + "  ".join(f"{cn}: {v:8.0f} cyc/tile" for cn, v in zip(cols, per_tile) if cn not in ("-", "tiles")))
