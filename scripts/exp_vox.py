"""Experiment: the voxelization stage alone on the bench workload — both paths timed with CUDA
events, segment / unit statistics (used for the ncu capture of vox_fast_kernel as well)."""
import sys, os
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
n = pr.n_vis.cpu().numpy()
print(f"segments {len(n)}: visible {total_vis}, per segment min {n.min()} mean {n.mean():.0f} max {n.max()}")
reps = int(os.environ.get("REPS", "20"))
for name, mode, unit in (("fast", 0, 0), ("slow", 1, 0)) if not os.environ.get("ONLY_FAST") else (("fast", 0, 0),):
    if mode == 0:
        units = np.maximum(1, -(-n // 7000))
        print(f"units {units.sum()} ({units.sum() / 148:.2f} waves), point visits {(units * n).sum()}")
    ws = ops._ws(ops.L.lib().xm3d_voxelize_ws_bytes(len(n), total_vis), dev)
    for _ in range(3):
        u = ops.voxelize_batch(pr.xyz_vis, pr.vis_off, pipe.rt, cap=total_vis, ws=ws, mode=mode, unit_pts=unit)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        u = ops.voxelize_batch(pr.xyz_vis, pr.vis_off, pipe.rt, cap=total_vis, ws=ws, mode=mode, unit_pts=unit)
    e1.record()
    torch.cuda.synchronize()
    print(f"{name}: {e0.elapsed_time(e1) / reps * 1000:.1f} us per voxelize_batch (eager, includes torch allocs), "
          f"voxels {int(u.m.sum())}, path {ops.voxel_path_info(u)}")
