"""Experiment (not part of the product): decompose the captured step into its stage graphs."""
import sys, os, types
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
pipe.set_cap(total_vis)
masks, mode, _ = bench.make_masks(args, batch.n_views, dev, 4242)
feat = torch.empty((total_vis, args.c), dtype=torch.float32, device=dev).normal_()
for _ in range(3):
    out = pipe.run(masks, feat, mode)
torch.cuda.synchronize()

def timeit(fn, name, n=50):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        fn()
    for _ in range(5): g.replay()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): g.replay()
    e1.record(); torch.cuda.synchronize()
    print(f"{name:40s} {e0.elapsed_time(e1)/n*1000:8.1f} us", flush=True)

pr = pipe.project()
torch.cuda.synchronize()
def f_proj(): pipe.project()
def f_vox(): ops.voxelize_batch(pr.xyz_vis, pr.vis_off, pipe.rt, cap=pipe.cap_vis, collate=True, ws=pipe.ws_vox)
def f_gather(): return ops.gather_masks(masks, pr.rowcol, pr.vis_off, mode=mode, cap=pipe.cap_vis, ws=pipe.ws_gather)
member, _ = f_gather()
def f_pool(): ops.pool(feat, pr.vis_off, pipe.k, member=member, cap=pipe.cap_vis, cap_pairs=pipe.cap_pairs, ws=pipe.ws_pool, status=pipe._side_status)
def f_gp():
    m, _ = f_gather()
    ops.pool(feat, pr.vis_off, pipe.k, member=m, cap=pipe.cap_vis, cap_pairs=pipe.cap_pairs, ws=pipe.ws_pool, status=pipe._side_status)
side = torch.cuda.Stream()
def f_gp_vox():
    main = torch.cuda.current_stream()
    side.wait_stream(main)
    with torch.cuda.stream(side):
        f_gp()
    f_vox()
    main.wait_stream(side)
hi = torch.cuda.Stream(priority=-1)
def f_gp_vox_prio():
    main = torch.cuda.current_stream()
    hi.wait_stream(main)
    with torch.cuda.stream(hi):
        f_vox()
    f_gp()
    main.wait_stream(hi)
def f_full_prio():
    main = torch.cuda.current_stream()
    prj = pipe.project()
    hi.wait_stream(main)
    with torch.cuda.stream(hi):
        ops.voxelize_batch(prj.xyz_vis, prj.vis_off, pipe.rt, cap=pipe.cap_vis, collate=True, ws=pipe.ws_vox)
    m, _ = ops.gather_masks(masks, prj.rowcol, prj.vis_off, mode=mode, cap=pipe.cap_vis, ws=pipe.ws_gather)
    ops.pool(feat, prj.vis_off, pipe.k, member=m, cap=pipe.cap_vis, cap_pairs=pipe.cap_pairs, ws=pipe.ws_pool, status=pipe._side_status)
    main.wait_stream(hi)
def f_full(): pipe.run(masks, feat, mode)
pipe2 = CorrespondencePipeline(batch, args.k, args.c, dev, overlap=False); pipe2.xyz = pipe.xyz; pipe2.depth = pipe.depth; pipe2.set_cap(total_vis)
pipe2.pairs_per_point = pipe.pairs_per_point; pipe2._size_pool_ws()
def f_serial(): pipe2.run(masks, feat, mode)
import os
QUICK = os.environ.get("QUICK")
if not QUICK: timeit(f_proj, "project")
if not QUICK: timeit(f_vox, "voxelize")
if not QUICK: timeit(f_gather, "gather")
if not QUICK or QUICK == "2": timeit(f_pool, "pool")
if not QUICK: timeit(f_gp, "gather+pool")
if not QUICK: timeit(f_gp_vox, "gather+pool || voxelize")
if not QUICK: timeit(f_gp_vox_prio, "gather+pool || voxelize(high prio)")
if not QUICK: timeit(f_full_prio, "full, voxelize on high-prio stream")
if not QUICK: timeit(f_serial, "full serial")
if not QUICK or QUICK == "2": timeit(f_full, "full overlap")

if QUICK == '2': sys.exit(0)
# ---- does voxelize slow down next to an HBM-saturating copy that uses no SMs (copy engine)?
import ctypes
rt_ = ctypes.CDLL('libcudart.so.12')
rt_.cudaMemcpyAsync.argtypes=[ctypes.c_void_p,ctypes.c_void_p,ctypes.c_size_t,ctypes.c_int,ctypes.c_void_p]
big_a = torch.empty(3 << 30, dtype=torch.uint8, device=dev); big_b = torch.empty_like(big_a)
cs = torch.cuda.Stream()
def ev(): return torch.cuda.Event(enable_timing=True)
for trial in range(3):
    torch.cuda.synchronize()
    a0, a1, b0, b1 = ev(), ev(), ev(), ev()
    with torch.cuda.stream(cs):
        b0.record()
        assert rt_.cudaMemcpyAsync(big_b.data_ptr(), big_a.data_ptr(), big_a.numel(), 3, cs.cuda_stream) == 0
        b1.record()
    a0.record()
    for _ in range(2): f_vox()
    a1.record()
    torch.cuda.synchronize()
    print(f"vox x2 next to a CE copy: {a0.elapsed_time(a1)*500:.1f} us per vox;  copy {b0.elapsed_time(b1)*1000:.0f} us ({2*big_a.numel()/b0.elapsed_time(b1)/1e6:.0f} GB/s r+w)")
torch.cuda.synchronize()
a0, a1 = ev(), ev(); a0.record()
for _ in range(2): f_vox()
a1.record(); torch.cuda.synchronize()
print(f"vox x2 alone (eager): {a0.elapsed_time(a1)*500:.1f} us per vox")
