"""Experiment: warm per-kernel durations (CUPTI via torch.profiler) of the eager step and of the graph replay."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch, collections
import bench
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
for _ in range(3): pipe.run(masks, feat, mode)
torch.cuda.synchronize()
pipe.capture(masks, feat, mode)
for _ in range(3): pipe.replay()
torch.cuda.synchronize()
from torch.profiler import profile, ProfilerActivity
for label, fn in (("graph replay", pipe.replay), ("eager serial", lambda: pipe.run(masks, feat, mode, times=__import__('xmask3d_b200.pipeline', fromlist=['StageTimes']).StageTimes()))):
    with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
        for _ in range(10): fn()
        torch.cuda.synchronize()
    acc = collections.OrderedDict()
    t0 = None; t1 = 0
    for e in prof.events():
        if e.device_type == torch.autograd.DeviceType.CUDA:
            acc.setdefault(e.name[:60], []).append(e.device_time if hasattr(e, "device_time") else e.cuda_time)
            st = e.time_range.start; en = e.time_range.end
            t0 = st if t0 is None else min(t0, st); t1 = max(t1, en)
    print(f"== {label}: span {(t1 - t0) / 10:.1f} us per step")
    tot = 0
    for k, v in acc.items():
        print(f"  {k:60s} n={len(v)//10:3d} {sum(v)/10:8.1f} us")
        tot += sum(v) / 10
    print(f"  sum of kernel times {tot:.1f} us")
