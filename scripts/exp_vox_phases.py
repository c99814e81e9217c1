"""Experiment: per-phase clock64 stamps of vox_fast_kernel (needs the -DXM3D_FV_TIMING build:
XM3D_SO=xmask3d_b200/libxm3d_dbg.so)."""
import sys, os, ctypes as C
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import bench
from xmask3d_b200 import ops
from xmask3d_b200.pipeline import CorrespondencePipeline
args = bench.parse()
dev = torch.device("cuda", 0)
batch, scenes = bench.build_batch(args, 0)
pipe = CorrespondencePipeline(batch, args.k, args.c, dev)
pipe.upload(torch.from_numpy(batch.xyz).pin_memory(), torch.from_numpy(batch.depth_mm.view(np.int16)).pin_memory())
pr = pipe.project()
total_vis = int(pr.n_vis.sum().item())
lib = ops.L.lib()
lib.xm3d_voxel_debug.restype = C.c_int
lib.xm3d_voxel_debug.argtypes = [C.c_void_p]
dbg = torch.zeros(4096 * 32, dtype=torch.int64, device=dev)
ws = ops._ws(lib.xm3d_voxelize_ws_bytes(pipe.n_views, total_vis), dev)
for _ in range(3):
    u = ops.voxelize_batch(pr.xyz_vis, pr.vis_off, pipe.rt, cap=total_vis, ws=ws)
torch.cuda.synchronize()
assert lib.xm3d_voxel_debug(dbg.data_ptr()) == 0
u = ops.voxelize_batch(pr.xyz_vis, pr.vis_off, pipe.rt, cap=total_vis, ws=ws)
torch.cuda.synchronize()
d = dbg.cpu().numpy().reshape(-1, 32)
d = d[d[:, 8] != 0]
print("units", len(d))
seq = [0, 1, 2, 3, 9, 10, 11, 12, 13, 14, 4, 5, 6, 7, 8]     # stamp order inside the kernel
ph = np.diff(d[:, seq], axis=1) / 1000.0          # kcycles
names = ["partition", "clear", "insert", "uniq-scan", "rank:sort", "rank:search", "rank:scan", "rank:scatter",
         "rank:count", "rank:write", "lookback", "first-init", "inverse", "output"]
print("phase kcycles: mean / median / max")
for i, nm in enumerate(names):
    print(f"  {nm:12s} {ph[:, i].mean():8.1f} {np.median(ph[:, i]):8.1f} {ph[:, i].max():8.1f}")
tot = (d[:, 8] - d[:, 0]) / 1000.0
print(f"  total      {tot.mean():8.1f} {np.median(tot):8.1f} {tot.max():8.1f}")
t0 = d[:, 28].min()
st, en = (d[:, 28] - t0) / 1000.0, (d[:, 29] - t0) / 1000.0
print(f"wall: first start 0, last end {en.max():.1f} us; unit duration mean {np.mean(en - st):.1f} us, max {np.max(en - st):.1f} us")
order = np.argsort(st)
for q in (0, 100, 147, 148, 200, 300, 400, len(d) - 1):
    if q < len(d):
        i = order[q]
        print(f"  unit#{q:3d} (blk {i}) start {st[i]:7.1f} end {en[i]:7.1f} n={d[i, 24]} M={d[i, 25]} P={d[i, 26]} sm={d[i, 27]}")
# per-SM busy time
sm = d[:, 27]
busy = np.array([np.sum((en - st)[sm == k]) for k in np.unique(sm)])
print(f"per-SM busy us: mean {busy.mean():.1f} max {busy.max():.1f} (SMs used {len(busy)})")
