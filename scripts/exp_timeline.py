"""Experiment: kernel timeline of ONE graph replay of the headline step (CUPTI via torch.profiler): start offset,
duration and stream of every kernel — where the step's time goes and what overlaps."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import bench
args = bench.parse([a for a in sys.argv[1:]])
dev = torch.device("cuda", 0)
batch, scenes = bench.build_batch(args, 0)
pipe, masks, mode, mask_bytes, feat, xyz_h, depth_h, n_vis, total_vis, total_pairs = bench.prepare_pipeline(
    batch, args.k, args.c, dev, args.masks, 4242, vox_mode=2)
for _ in range(3):
    pipe.run(masks, feat, mode)
torch.cuda.synchronize()
pipe.capture(masks, feat, mode)
for _ in range(5):
    pipe.replay()
torch.cuda.synchronize()
from torch.profiler import profile, ProfilerActivity
with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
    for _ in range(4):
        pipe.replay()
    torch.cuda.synchronize()
evs = [e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA]
evs.sort(key=lambda e: e.time_range.start)
# split into replays by large gaps
names = [e.name for e in evs]
first = names[0]
starts = [i for i, e in enumerate(evs) if e.name == first]
per = len(evs) // 4
rep = evs[2 * per:3 * per]
t0 = rep[0].time_range.start
print(f"kernels per replay {per}; span {(rep[-1].time_range.end - t0):.1f} us")
for e in rep:
    print(f"{e.time_range.start - t0:9.1f} +{e.time_range.end - e.time_range.start:8.1f} us  {e.name[:70]}")
