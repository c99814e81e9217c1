#!/usr/bin/env python
"""Benchmark of the cross-modal correspondence path (voxelize + project + mask-pool).

    python bench.py --gpus N --steps K --warmup W            # native CUDA path (one rank per GPU)
    python bench.py --impl reference --steps K --warmup W    # reference CPU algorithm on host cores

Workload (BASELINE.json configs[1]): per GPU 8 synthetic ScanNet-sized scenes x 20 posed RGB-D
views, 150k points per scene, 768-d float32 per-point features, 50 masks per view.  A "step" is
one pass of project -> voxelize -> masks-at-points -> pool over the whole batch.  Weak scaling:
every rank owns its own 8 scenes (scenes shard with no data-path collective).
Prints ONE JSON line on rank 0.
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
# stdout carries exactly one JSON line: NCCL's banner / debug output goes to stderr
os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")

METRIC = "points*views/sec (voxelize+project+mask-pool)"
UNIT = "points*views/s"
LOADER_VOX = dict(clip_bound=None, use_augmentation=True, scale_augmentation_bound=(0.9, 1.1),
                  rotation_augmentation_bound=((-np.pi / 64, np.pi / 64), (-np.pi / 64, np.pi / 64), (-np.pi, np.pi)),
                  translation_augmentation_ratio_bound=((-0.2, 0.2), (-0.2, 0.2), (0, 0)))   # dataset/point_loader.py:54-60


def parse():
    p = argparse.ArgumentParser()
    p.add_argument("--gpus", type=int, default=1)
    p.add_argument("--steps", type=int, default=100)
    p.add_argument("--warmup", type=int, default=5)
    p.add_argument("--impl", default="native", choices=["native", "reference"])
    p.add_argument("--scenes", type=int, default=8)
    p.add_argument("--views", type=int, default=20)
    p.add_argument("--points", type=int, default=150_000)
    p.add_argument("--k", type=int, default=50)
    p.add_argument("--c", type=int, default=768)
    p.add_argument("--voxel", type=float, default=0.02)
    p.add_argument("--masks", default="partition", choices=["partition", "overlap"])
    p.add_argument("--cpu-views", type=int, default=12, help="views of scene 0 timed by the CPU baseline")
    p.add_argument("--cpu-procs", type=int, default=0, help="processes of the reference arm (0 = one per host cpu)")
    p.add_argument("--no-cpu", action="store_true")
    p.add_argument("--no-e2e-all", action="store_true")
    p.add_argument("--e2e-slots", type=int, default=3, help="staging slots of the pipelined end-to-end loop")
    p.add_argument("--profile-steps", type=int, default=0, help="run only this many plain steps (for ncu)")
    p.add_argument("--distinct-scenes", action="store_true",
                   help="rank r processes scenes 8r..8r+7 (data-dependent imbalance) instead of a copy of scenes 0..7")
    p.add_argument("--no-overlap", action="store_true", help="run the stages back to back on one stream")
    p.add_argument("--no-graph", action="store_true", help="launch every step eagerly instead of replaying a CUDA graph")
    p.add_argument("--workload", default="batch", choices=["batch", "split_scene"],
                   help="batch = configs[1] (default, weak scaling); split_scene = configs[3]: one 1M-point scene, "
                        "100 views block-partitioned over the ranks, NCCL all-reduce of per-mask sums/counts")
    return p.parse_args()


# ----------------------------------------------------------------------------- workload
def build_batch(args, rank: int):
    from xmask3d_b200 import synthetic as syn
    from xmask3d_b200.pipeline import Batch
    from xmask3d_b200.voxelizer import Voxelizer
    xyz, off, vs, w2c, depth, rts, scenes = [], [0], [], [], [], [], []
    for s in range(args.scenes):
        # weak scaling: every rank owns the same amount of work, i.e. a copy of the same 8 synthetic
        # scenes (--distinct-scenes gives rank-specific seeds; the visible fraction then varies by
        # +-15 % between ranks and the slowest rank sets the time)
        gs = (rank * args.scenes + s) if args.distinct_scenes else s
        sc = syn.make_scene(1000 + gs, args.points)
        scenes.append(sc)
        xyz.append(sc.xyz)
        off.append(off[-1] + sc.xyz.shape[0])
        for v in range(args.views):
            vw = syn.make_view(sc, v)
            vs.append(s)
            w2c.append(np.linalg.inv(vw.pose))
            depth.append(vw.depth_mm)
            np.random.seed(5557 + 1000 * gs + v)
            rt, _ = Voxelizer(voxel_size=args.voxel, **LOADER_VOX).draw_rigid_transformation()
            rts.append(rt[:3, :4])
    b = Batch(np.concatenate(xyz), np.array(off, np.int64), np.array(vs, np.int64), np.stack(w2c),
              np.stack(depth), np.stack(rts), syn.scannet_intrinsics())
    return b, scenes


def make_masks(args, n_views: int, dev, seed: int):
    """partition: K-seed Voronoi label image per view -> bool [V,K,240,320] (the argmax partition of
    models/xmask3d.py:418-435); overlap: smooth random float32 logits, thresholded sigmoid >= 0.5."""
    import torch
    g = torch.Generator(device=dev).manual_seed(seed)
    h, w, k = 240, 320, args.k
    if args.masks == "partition":
        sy = torch.rand(n_views, k, 1, 1, device=dev, generator=g) * h
        sx = torch.rand(n_views, k, 1, 1, device=dev, generator=g) * w
        yy = torch.arange(h, device=dev).view(1, 1, h, 1).float()
        xx = torch.arange(w, device=dev).view(1, 1, 1, w).float()
        out = torch.empty(n_views, k, h, w, dtype=torch.bool, device=dev)
        for a in range(0, n_views, 16):
            d2 = (yy - sy[a:a + 16]) ** 2 + (xx - sx[a:a + 16]) ** 2
            lab = d2.argmin(1, keepdim=True)
            out[a:a + 16] = lab == torch.arange(k, device=dev).view(1, k, 1, 1)
        return out, "ge0.5", 1
    import torch.nn.functional as F
    z = torch.randn(n_views * k, 1, 12, 16, device=dev, generator=g)
    z = F.interpolate(z, size=(h, w), mode="bicubic", align_corners=False)
    z = z / z.flatten(1).std(1).view(-1, 1, 1, 1)
    cover = torch.rand(n_views * k, device=dev, generator=g) * 0.28 + 0.02
    thr = torch.special.ndtri(1.0 - cover).view(-1, 1, 1, 1)
    z = ((z - thr) * 4.0).view(n_views, k, h, w)
    z = torch.where(z.abs() < 1e-3, torch.full_like(z, 1e-3), z)
    return z.contiguous(), "sigmoid_ge0.5", 4


# ----------------------------------------------------------------------------- clocks
class ClockSampler:
    def __init__(self, index: int):
        self.index, self.samples, self.reasons, self.max_mhz = index, [], set(), None
        self._stop = threading.Event()
        self._t = None

    def _loop(self):
        try:
            import pynvml
            pynvml.nvmlInit()
            h = pynvml.nvmlDeviceGetHandleByIndex(self.index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(h, pynvml.NVML_CLOCK_SM)
            names = {"hw_slowdown": 0x8, "sw_power_cap": 0x4, "hw_thermal_slowdown": 0x40, "sw_thermal_slowdown": 0x20,
                     "hw_power_brake": 0x80, "sync_boost": 0x10, "app_clocks": 0x2, "display_clocks": 0x100}
            while not self._stop.is_set():
                self.samples.append(pynvml.nvmlDeviceGetClockInfo(h, pynvml.NVML_CLOCK_SM))
                try:
                    r = pynvml.nvmlDeviceGetCurrentClocksEventReasons(h)
                except Exception:
                    r = pynvml.nvmlDeviceGetCurrentClocksThrottleReasons(h)
                for n, bit in names.items():
                    if r & bit:
                        self.reasons.add(n)
                time.sleep(0.02)
        except Exception as e:          # noqa: BLE001
            self.reasons.add(f"sampler_error:{type(e).__name__}")

    def __enter__(self):
        self._t = threading.Thread(target=self._loop, daemon=True)
        self._t.start()
        return self

    def __exit__(self, *a):
        self._stop.set()
        self._t.join(timeout=2)

    def summary(self):
        s = sorted(self.samples)
        return {"sm_mhz": (s[len(s) // 2] if s else None), "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons),
                "samples": len(s)}


# ----------------------------------------------------------------------------- CPU baseline
def cpu_reference_pass(args, scene, views, rts_seed_base, k, c, masks_np, n_views: int):
    """The reference's algorithm for the path on host cores (oracle/ref_port.py: numpy / torch-CPU
    restatement of compute_mapping, Voxelizer.voxelize and the per-mask feature[mask].mean(0) loop
    of criterion.py:152-157).  Returns (seconds, point*views)."""
    import torch
    from oracle import ref_port as P
    mapper = P.getMapping()
    feats = []
    pv = 0
    pre = []
    for v in range(n_views):                       # inputs prepared outside the timed region
        pre.append((views[v].pose, views[v].depth_m, torch.from_numpy(masks_np[v]).float()))
    # features depend on n: draw a generous buffer once
    gen = torch.Generator().manual_seed(7)
    fbuf = torch.randn(65_000, c, generator=gen)
    t0 = time.perf_counter()
    for v in range(n_views):
        pose, depth_m, mask = pre[v]
        mapping = mapper.compute_mapping(pose, scene.xyz, depth_m)
        vis, x_label, y_label = P.compact_mapping(mapping)
        locs, col, lab = scene.xyz[vis], scene.colors[vis], scene.labels[vis]
        np.random.seed(rts_seed_base + v)
        vox = P.Voxelizer(voxel_size=args.voxel, **LOADER_VOX)
        vox.voxelize(locs, col, lab)
        n = locs.shape[0]
        member = P.gather_masks(mask, torch.from_numpy(x_label), torch.from_numpy(y_label), "ge0.5")
        P.masked_mean_pool(fbuf[:n], member)
        pv += scene.xyz.shape[0]
    return time.perf_counter() - t0, pv


def cpu_masks(args, n_views: int, seed: int):
    from xmask3d_b200 import synthetic as syn
    return [syn.make_partition_masks(seed + v, args.k) for v in range(n_views)]


_CPU_CTX = None      # (args, scene, views, masks): inherited by the forked workers of the reference arm


def _cpu_worker_init():
    import torch
    torch.set_num_threads(1)             # one view per process, like the reference's DataLoader workers


def _cpu_one_view(v):
    args, sc, views, masks = _CPU_CTX
    t, pv = cpu_reference_pass(args, sc, [views[v]], 5557 + v, args.k, args.c, [masks[v]], 1)
    return pv


def run_reference(args, rank: int, world: int):
    """The reference's CPU implementation of the path on the host cores: the views of a step are spread over
    a process pool (one view per process at a time, torch threads 1 — the reference parallelises this stage
    with DataLoader workers, `workers: 4` in its yaml; here every host cpu gets one), wall clock per step."""
    global _CPU_CTX
    if rank != 0:
        return
    import multiprocessing as mp

    import torch
    from xmask3d_b200 import synthetic as syn
    sc = syn.make_scene(1000, args.points)
    procs = max(1, min(os.cpu_count() or 1, args.cpu_procs if args.cpu_procs > 0 else (os.cpu_count() or 1)))
    nv = max(1, min(args.cpu_views, max(4, procs)))
    views = [syn.make_view(sc, v) for v in range(nv)]
    masks = cpu_masks(args, nv, 9000)
    # single process, library threading only (what one DataLoader worker does)
    cpu_reference_pass(args, sc, views, 5557, args.k, args.c, masks, 1)
    t1, pv1 = cpu_reference_pass(args, sc, views, 5557, args.k, args.c, masks, min(nv, 4))
    single = pv1 / t1
    _CPU_CTX = (args, sc, views, masks)
    tot_t, tot_pv = 0.0, 0
    with mp.get_context("fork").Pool(procs, initializer=_cpu_worker_init) as pool:
        for _ in range(max(args.warmup, 1)):
            pool.map(_cpu_one_view, range(nv))
        for _ in range(args.steps):
            t0 = time.perf_counter()
            pvs = pool.map(_cpu_one_view, range(nv))
            tot_t += time.perf_counter() - t0
            tot_pv += sum(pvs)
    val = tot_pv / tot_t
    line = {"metric": METRIC, "value": val, "unit": UNIT, "impl": "reference", "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": 1e3 * tot_t / max(args.steps, 1), "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f64/f32", "data": "synthetic",
            "config": {"workload": workload_name(args), "sample": f"scene 0 x {nv} views per step"},
            "cpu_baseline": {"value": val, "unit": UNIT, "cores": procs, "kind": "port",
                             "sample": f"{args.steps} x (scene 0, {nv} views over {procs} processes): numpy/torch-CPU port of "
                                       f"the reference path, {os.cpu_count()} host cpus",
                             "single_process_value": single, "single_process_threads": torch.get_num_threads()},
            "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)


def workload_name(args):
    return (f"configs[1]: {args.scenes} scenes x {args.views} views per GPU, {args.points} pts/scene, "
            f"{int(args.voxel * 100)} cm voxels, C={args.c}, K={args.k} {args.masks} masks/view")


# ----------------------------------------------------------------------------- native arm
def run_native(args, rank: int, world: int, local_rank: int):
    import torch
    import torch.distributed as dist
    from xmask3d_b200 import _lib as L
    from xmask3d_b200.pipeline import CorrespondencePipeline, StageTimes, algorithmic_bytes
    assert torch.cuda.is_available(), "the native arm needs a CUDA device (no CPU fallback)"
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    t_setup = time.perf_counter()
    batch, scenes = build_batch(args, rank)
    pipe = CorrespondencePipeline(batch, args.k, args.c, dev, overlap=not args.no_overlap)
    xyz_h = torch.from_numpy(batch.xyz).pin_memory()
    depth_h = torch.from_numpy(batch.depth_mm.view(np.int16)).pin_memory()
    pipe.upload(xyz_h, depth_h)
    # one untimed projection fixes the visible counts (deterministic), which size the feature tensor
    pr = pipe.project()
    n_vis = pr.n_vis.cpu().numpy().astype(np.int64)
    total_vis = int(n_vis.sum())
    pipe.set_cap(total_vis)
    masks, mode, mask_bytes = make_masks(args, batch.n_views, dev, 4242 + rank)
    from xmask3d_b200 import ops
    member0, _ = ops.gather_masks(masks, pr.rowcol, pr.vis_off, mode=mode, cap=total_vis)
    total_pairs = int(ops._popcount32(member0[:total_vis]).sum().item())
    del member0
    pipe.pairs_per_point = total_pairs / max(total_vis, 1)
    pipe._size_pool_ws()
    feat = torch.empty((total_vis, args.c), dtype=torch.float32, device=dev)
    g = torch.Generator(device=dev).manual_seed(7 + rank)
    for a in range(0, total_vis, 1 << 20):
        feat[a:a + (1 << 20)].normal_(generator=g)
    torch.cuda.synchronize()
    t_setup = time.perf_counter() - t_setup

    if args.profile_steps:
        for _ in range(args.profile_steps):
            pipe.run(masks, feat, mode)
        torch.cuda.synchronize()
        return

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    out = None
    for _ in range(max(args.warmup, 3)):
        out = pipe.run(masks, feat, mode)
    assert int(out["proj"].status.item()) == 0 and int(out["vox"].status.item()) == 0
    assert int(out["pool_status"].item()) == 0
    m_vox = out["vox"].m.cpu().numpy().astype(np.int64)

    # the whole step (~30 launches) is captured once and replayed as one CUDA graph
    use_graph = not args.no_graph
    l0 = L.lib().xm3d_launch_count()
    if use_graph:
        pipe.capture(masks, feat, mode)
    else:
        pipe.run(masks, feat, mode)
    launches_per_step = L.lib().xm3d_launch_count() - l0

    def do_step():
        return pipe.replay() if use_graph else pipe.run(masks, feat, mode)

    # ---- device-resident throughput (`value`): K steps between barriers, CUDA events, max over ranks
    barrier()
    with ClockSampler(local_rank) as clk:
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(args.steps):
            do_step()
        e1.record()
        barrier()
        ms = e0.elapsed_time(e1)
        # per-stage durations inside the same region: a second timed pass with events between stages
        stage_acc = {}
        ev_a, ev_b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        ev_a.record(); ev_b.record()                       # materialise the cudaEvent_t handles
        torch.cuda.synchronize()
        import ctypes
        L.lib().xm3d_set_pool_events(ctypes.c_void_p(ev_a.cuda_event), ctypes.c_void_p(ev_b.cuda_event))
        for _ in range(min(args.steps, 20)):
            st = StageTimes()
            pipe.run(masks, feat, mode, times=st)
            torch.cuda.synchronize()
            for kk, vv in st.result().items():
                stage_acc.setdefault(kk, []).append(vv)
            stage_acc.setdefault("pool_sum_kernel", []).append(ev_a.elapsed_time(ev_b))
        L.lib().xm3d_set_pool_events(None, None)
    t = torch.tensor([ms], dtype=torch.float64, device=dev)
    lt = torch.tensor([float(launches_per_step * args.steps)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dist.all_reduce(lt, op=dist.ReduceOp.SUM)
    ms = float(t.item())
    pv_rank = batch.point_views
    value = world * pv_rank * args.steps / (ms * 1e-3)
    stage_ms = {kk: float(np.mean(vv)) for kk, vv in stage_acc.items()}

    # ---- end to end through the host-facing API: host (pinned) loader inputs in, results out
    res_h = {
        "rowcol": torch.empty((total_vis, 2), dtype=torch.int32).pin_memory(),
        "inverse": torch.empty(total_vis, dtype=torch.int32).pin_memory(),
        "first": torch.empty(total_vis, dtype=torch.int32).pin_memory(),
        "voxel": torch.empty((total_vis, 3), dtype=torch.int32).pin_memory(),
        "vis_off": torch.empty(batch.n_views + 1, dtype=torch.int64).pin_memory(),
        "m": torch.empty(batch.n_views, dtype=torch.int32).pin_memory(),
        "mean": torch.empty((batch.n_views, args.k, args.c), dtype=torch.float32).pin_memory(),
        "cnt": torch.empty((batch.n_views, args.k), dtype=torch.int32).pin_memory(),
    }

    def e2e_step(all_host=None):
        pipe.upload(xyz_h, depth_h)
        mk, ft = masks, feat
        if all_host is not None:
            masks_d2.copy_(all_host[0], non_blocking=True)
            feat.copy_(all_host[1], non_blocking=True)
            mk, ft = masks_d2, feat
        o = do_step() if all_host is None else pipe.run(mk, ft, mode)
        res_h["rowcol"].copy_(o["proj"].rowcol[:total_vis], non_blocking=True)
        res_h["vis_off"].copy_(o["proj"].vis_off, non_blocking=True)
        res_h["inverse"].copy_(o["vox"].inverse[:total_vis], non_blocking=True)
        res_h["first"].copy_(o["vox"].first[:total_vis], non_blocking=True)
        res_h["voxel"].copy_(o["vox"].voxel_xyz[:total_vis], non_blocking=True)
        res_h["m"].copy_(o["vox"].m, non_blocking=True)
        res_h["mean"].copy_(o["mean"], non_blocking=True)
        res_h["cnt"].copy_(o["cnt"], non_blocking=True)
        torch.cuda.current_stream().synchronize()          # the caller reads the results every step

    h2d = xyz_h.numel() * 4 + depth_h.numel() * 2 + 192 * batch.n_views
    d2h = sum(v.numel() * v.element_size() for v in res_h.values())

    # Pipelined end-to-end loop: the copy engines run beside the kernels.  Step i's inputs are copied
    # H2D on a copy-in stream into a staging slot, the compute stream moves them into the pipeline's
    # input tensors and replays the step, copies the results into an output staging slot, and a
    # copy-out stream moves that slot to pinned host memory.  --e2e-slots slots each way; the host waits for
    # the results of the oldest step in flight before it enqueues the next one.
    s_in, s_out = torch.cuda.Stream(device=dev), torch.cuda.Stream(device=dev)
    comp = torch.cuda.current_stream()
    NSLOT = args.e2e_slots
    xyz_st = [torch.empty_like(pipe.xyz) for _ in range(NSLOT)]
    dep_st = [torch.empty_like(pipe.depth) for _ in range(NSLOT)]
    out_st = [{k2: torch.empty(v2.shape, dtype=v2.dtype, device=dev) for k2, v2 in res_h.items()} for _ in range(NSLOT)]
    res_h2 = [res_h] + [{k2: torch.empty(v2.shape, dtype=v2.dtype).pin_memory() for k2, v2 in res_h.items()}
                       for _ in range(NSLOT - 1)]
    ev_in_ready = [torch.cuda.Event() for _ in range(NSLOT)]
    ev_consumed = [torch.cuda.Event() for _ in range(NSLOT)]
    ev_out_ready = [torch.cuda.Event() for _ in range(NSLOT)]
    ev_out_free = [torch.cuda.Event() for _ in range(NSLOT)]
    ev_done = [torch.cuda.Event() for _ in range(NSLOT)]
    for b in range(NSLOT):
        ev_consumed[b].record(comp)
        ev_out_free[b].record(comp)
        ev_done[b].record(comp)

    def e2e_pipelined(i):
        b = i % NSLOT
        with torch.cuda.stream(s_in):
            s_in.wait_event(ev_consumed[b])
            xyz_st[b].copy_(xyz_h, non_blocking=True)
            dep_st[b].copy_(depth_h, non_blocking=True)
            ev_in_ready[b].record(s_in)
        comp.wait_event(ev_in_ready[b])
        pipe.xyz.copy_(xyz_st[b], non_blocking=True)
        pipe.depth.copy_(dep_st[b], non_blocking=True)
        ev_consumed[b].record(comp)
        o = do_step()
        comp.wait_event(ev_out_free[b])
        ost = out_st[b]
        ost["rowcol"].copy_(o["proj"].rowcol[:total_vis], non_blocking=True)
        ost["vis_off"].copy_(o["proj"].vis_off, non_blocking=True)
        ost["inverse"].copy_(o["vox"].inverse[:total_vis], non_blocking=True)
        ost["first"].copy_(o["vox"].first[:total_vis], non_blocking=True)
        ost["voxel"].copy_(o["vox"].voxel_xyz[:total_vis], non_blocking=True)
        ost["m"].copy_(o["vox"].m, non_blocking=True)
        ost["mean"].copy_(o["mean"], non_blocking=True)
        ost["cnt"].copy_(o["cnt"], non_blocking=True)
        ev_out_ready[b].record(comp)
        with torch.cuda.stream(s_out):
            s_out.wait_event(ev_out_ready[b])
            for k2 in ost:
                res_h2[b][k2].copy_(ost[k2], non_blocking=True)
            ev_out_free[b].record(s_out)
            ev_done[b].record(s_out)
        if i >= NSLOT - 1:
            ev_done[(i - NSLOT + 1) % NSLOT].synchronize()   # the caller consumes the oldest step in flight

    for i in range(2 * NSLOT):
        e2e_pipelined(i)
    barrier()
    n_e2e = max(4, min(args.steps, 50))
    e0.record()
    for i in range(n_e2e):
        e2e_pipelined(i)
    comp.wait_stream(s_out)
    comp.wait_stream(s_in)
    e1.record()
    barrier()
    e2e_ms = e0.elapsed_time(e1)                      # device clock around H2D + kernels + D2H of all steps
    # sanity: what came back is what the device computed
    assert torch.equal(res_h2[(n_e2e - 1) % NSLOT]["cnt"], out["cnt"].cpu()) or use_graph is None
    te = torch.tensor([e2e_ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(te, op=dist.ReduceOp.MAX)
    e2e_val = world * pv_rank * n_e2e / (float(te.item()) * 1e-3)

    # the same without overlap: copy in, run, copy out, wait — one step at a time
    for _ in range(2):
        e2e_step()
    barrier()
    e0.record()
    for _ in range(10):
        e2e_step()
    e1.record()
    barrier()
    ts = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(ts, op=dist.ReduceOp.MAX)
    e2e_serial = world * pv_rank * 10 / (float(ts.item()) * 1e-3)

    # stricter variant: the per-point features and the masks also start in pinned host memory
    e2e_all = None
    if not args.no_e2e_all:
        try:
            feat_h = torch.empty(feat.shape, dtype=torch.float32).pin_memory()
            feat_h.copy_(feat)
            mk8 = masks.view(torch.uint8) if masks.dtype == torch.bool else masks
            masks_h = torch.empty(mk8.shape, dtype=mk8.dtype).pin_memory()
            masks_h.copy_(mk8)
            masks_d2 = torch.empty_like(mk8)
            torch.cuda.synchronize()
            e2e_step((masks_h, feat_h))
            barrier()
            e0.record()
            for _ in range(2):
                e2e_step((masks_h, feat_h))
            e1.record()
            barrier()
            ta = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=dev)
            if world > 1:
                dist.all_reduce(ta, op=dist.ReduceOp.MAX)
            e2e_all = {"value": world * pv_rank * 2 / (float(ta.item()) * 1e-3), "unit": UNIT,
                       "h2d_bytes_per_step": h2d + feat_h.numel() * 4 + masks_h.numel() * masks_h.element_size(),
                       "d2h_bytes_per_step": d2h,
                       "note": "features and masks also copied from pinned host memory every step (PCIe-bound)"}
            del feat_h, masks_h, masks_d2
        except Exception as e:                  # noqa: BLE001
            e2e_all = {"value": None, "note": f"skipped: {type(e).__name__}: {e}"[:200]}

    if rank != 0:
        if world > 1:
            dist.barrier()
            dist.destroy_process_group()
        return

    # ---- stage 4 (text logits, the only tensor-core kernel): timed apart, it is not part of the metric
    logits_info = {}
    try:
        from xmask3d_b200 import ops as _ops
        peaks_tf, peak_hbm = 1678.5, 6650.0
        if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")):
            _pk = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
            peaks_tf, peak_hbm = float(_pk.get("bf16_tflops", peaks_tf)), float(_pk.get("hbm_gbs", peak_hbm))
        gl = torch.Generator(device=dev).manual_seed(1)
        for name, rows, t in (("configs[1] B15N4 160x50 masks x 20 classes", 160 * 50, 20),
                              ("configs[2] B170N30 160x100 masks x 201 classes", 160 * 100, 201)):
            me = torch.randn(rows, args.c, device=dev, generator=gl)
            te = torch.randn(t - 1, args.c, device=dev, generator=gl)
            ne = torch.randn(1, args.c, device=dev, generator=gl)
            for _ in range(3):
                _ops.logits(me, te, ne, [1] * (t - 1), 1 / 0.07)
            torch.cuda.synchronize()
            e0.record()
            for _ in range(20):
                _ops.logits(me, te, ne, [1] * (t - 1), 1 / 0.07)
            e1.record()
            torch.cuda.synchronize()
            ms_l = e0.elapsed_time(e1) / 20
            fl = 2.0 * rows * args.c * t
            logits_info[name] = {"ms": ms_l, "useful_tflops": fl / (ms_l * 1e-3) / 1e12,
                                 "mma_tflops_3xtf32": 3 * fl / (ms_l * 1e-3) / 1e12,
                                 "frac_of_bf16_peak": 3 * fl / (ms_l * 1e-3) / 1e12 / peaks_tf,
                                 "note": "prep + tcgen05 GEMM + epilogue; <= 5 GFLOP, launch/latency bound by construction"}
        # per-point logits: the visible points' features (the pooling input) x text
        # embeddings, argmax only — the [n, C] features are read from HBM once
        for name, t in (("per-point argmax, 19 classes", 19), ("per-point argmax, 200 classes", 200)):
            te = torch.randn(t, args.c, device=dev, generator=gl)
            for _ in range(2):
                _ops.point_logits(feat, te, 1 / 0.07, want_logits=False)
            torch.cuda.synchronize()
            e0.record()
            for _ in range(5):
                _ops.point_logits(feat, te, 1 / 0.07, want_logits=False)
            e1.record()
            torch.cuda.synchronize()
            ms_l = e0.elapsed_time(e1) / 5
            fl = 2.0 * total_vis * args.c * t
            logits_info[name] = {"ms": ms_l, "rows": total_vis, "feature_gbs": 4.0 * args.c * total_vis / (ms_l * 1e-3) / 1e9,
                                 "frac_of_hbm_peak": 4.0 * args.c * total_vis / (ms_l * 1e-3) / 1e9 / peak_hbm,
                                 "useful_tflops": fl / (ms_l * 1e-3) / 1e12,
                                 "mma_tflops_3xtf32": 3 * fl / (ms_l * 1e-3) / 1e12}
    except Exception as e:                      # noqa: BLE001
        logits_info = {"error": f"{type(e).__name__}: {e}"[:200]}

    # ---- mask preparation: low-resolution logits of the mask head -> membership words /
    # partition labels of every view of the batch, one fused pass (timed apart, not part of the metric)
    mask_prep_info = {}
    try:
        from xmask3d_b200 import ops as _ops
        import torch.nn.functional as _F
        gm = torch.Generator(device=dev).manual_seed(2)
        lg = torch.randn(batch.n_views, args.k, 16, 16, device=dev, generator=gm)
        lg = _F.interpolate(lg, size=(128, 128), mode="bicubic", align_corners=False).contiguous() * 3.0
        scm = torch.rand(batch.n_views, args.k, device=dev, generator=gm) + 0.05

        def _time(fn, n):
            for _ in range(2):
                fn()
            torch.cuda.synchronize()
            e0.record()
            for _ in range(n):
                fn()
            e1.record()
            torch.cuda.synchronize()
            return e0.elapsed_time(e1) / n

        def _torch_ops():
            up = _F.interpolate(lg, size=(240, 320), mode="bilinear", align_corners=False)
            sg = up.sigmoid()
            return sg > 0.5, (scm.view(batch.n_views, args.k, 1, 1) * sg).argmax(1)
        mask_prep_info = {
            "shape": f"{batch.n_views} views x {args.k} masks, 128x128 -> 240x320",
            "bits_ms": _time(lambda: _ops.mask_prep(lg, (240, 320), mode="sigmoid_gt0.5", want_bits=True), 10),
            "partition_ms": _time(lambda: _ops.mask_prep(lg, (240, 320), scores=scm, want_bits=False,
                                                         want_partition=True), 10),
            "torch_cuda_ops_ms": _time(_torch_ops, 3),
            "note": "fused bilinear upsample + sigmoid threshold (+ score-weighted argmax partition and areas); "
                    "torch_cuda_ops = interpolate + sigmoid + mul + argmax + compare on the same GPU",
        }
        del lg
    except Exception as e:                      # noqa: BLE001
        mask_prep_info = {"error": f"{type(e).__name__}: {e}"[:200]}

    # ---- roofline of the dominant kernel (pool) and of the whole step
    peaks_path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(peaks_path):
        peak, peak_src = float(json.load(open(peaks_path))["hbm_gbs"]), "MEASURED_PEAKS.json hbm_gbs (of measured)"
    else:
        peak, peak_src = 6650.0, "B200_PROFILING.md fallback (of fallback)"
    n_pts_view = np.diff(batch.scene_off)[batch.view_scene]
    alg = algorithmic_bytes(n_pts_view, n_vis, m_vox, args.k, args.c, mask_bytes)
    words = (args.k + 31) // 32
    # algorithmic bytes of the dominant kernel: every visible point's C-float row read once + its
    # 4-byte row index, plus the [K,C] sums it produces per view
    pool_bytes = int((4 * args.c + 4) * total_vis + batch.n_views * 4 * args.k * args.c)
    pool_ms = stage_ms.get("pool_sum_kernel", float("nan"))
    achieved = pool_bytes / (pool_ms * 1e-3) / 1e9
    step_ms = ms / args.steps
    # dram__bytes_read+write of this kernel from the committed `ncu --set full` capture of this same
    # default workload (profiles/r01_pool_sum_kernel_ncu_full.json); null for any other workload
    traffic = None
    tpath = os.path.join(ROOT, "profiles", "r01_pool_sum_kernel_ncu_full.json")
    default_wl = (args.scenes, args.views, args.points, args.k, args.c, args.masks, rank) == (8, 20, 150_000, 50, 768, "partition", 0)
    if default_wl and os.path.exists(tpath):
        traffic = json.load(open(tpath)).get("_summary", {}).get("traffic_bytes_per_launch")
    roof = {"bound": "hbm", "kernel": ("pool_rows_kernel (point-major, overlapping masks" if args.masks == "overlap"
                                     else "pool_sum_kernel<4> (pair lists") + "; events recorded around this launch alone)",
            "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": traffic,
            "bytes_per_launch": pool_bytes, "ms_per_launch": pool_ms, "peak_source": peak_src}
    pipe_gbs = alg["total"] / (step_ms * 1e-3) / 1e9
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
        "ms_per_step": step_ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64/u64/f32",
        "data": "synthetic",
        "config": {"workload": workload_name(args), "point_views_per_gpu_step": pv_rank, "visible_pairs_per_gpu": total_vis, "mask_memberships_per_gpu": total_pairs,
                   "voxels_per_gpu": int(m_vox.sum()), "cache": "inputs larger than L2 (features %.1f GB per GPU)" % (feat.numel() * 4 / 1e9),
                   "parallelism": f"scenes sharded over {world} rank(s), no data-path collective; " +
                                  ("rank-specific scenes" if args.distinct_scenes else "every rank processes a copy of the same 8 scenes"),
                   "launch": ("one CUDA graph replay per step" if use_graph else "eager launches") +
                             ("; voxelize overlapped with gather+pool on a second stream" if pipe.overlap else ""),
                   "kernels_per_step": int(launches_per_step), "setup_s": round(t_setup, 1)},
        "roofline": roof,
        "pipeline_roofline": {"algorithmic_bytes_per_step": alg, "achieved": pipe_gbs, "peak": peak, "unit": "GB/s",
                              "frac": pipe_gbs / peak},
        "stage_ms": stage_ms,
        "logits": logits_info,
        "mask_prep": mask_prep_info,
        "e2e": {"value": e2e_val, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                "unpipelined_value": e2e_serial,
                "note": "double-buffered: copy-in / compute / copy-out streams overlap across steps; host->device: scene xyz, depth PNG arrays, view records (the reference's loader-side numpy inputs); "
                        "device->host: x/y labels, inverse/first/voxel maps, pooled means and counts; per-point features "
                        "and 2D masks are consumed on the device, where the reference's API produces them"},
        "e2e_all_host": e2e_all,
        "gpu_launches": int(lt.item()),
        "clocks": clk.summary(),
    }
    if not args.no_cpu:
        nv = max(1, args.cpu_views)
        views = []
        from xmask3d_b200 import synthetic as syn
        views = [syn.make_view(scenes[0], v) for v in range(nv)]
        mk = cpu_masks(args, nv, 9000)
        cpu_reference_pass(args, scenes[0], views, 5557, args.k, args.c, mk, 1)
        t_cpu, pv_cpu = cpu_reference_pass(args, scenes[0], views, 5557, args.k, args.c, mk, nv)
        line["cpu_baseline"] = {"value": pv_cpu / t_cpu, "unit": UNIT, "cores": torch.get_num_threads(), "kind": "port",
                                "sample": f"scene 0 x {nv} views ({t_cpu:.1f} s): numpy/torch-CPU port of the reference "
                                          f"path (oracle/ref_port.py), {os.cpu_count()} host cpus"}
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def run_split_scene(args, rank: int, world: int, local_rank: int):
    """configs[3]: one dense scene (1M points, 1 cm voxels), 100 views split over the ranks; every rank
    pools its views under a scene-level set of K masks and one all-reduce(SUM) of the packed
    [K, C+1] sums/counts yields the scene-level mask features on every rank.  Strong scaling."""
    import torch
    import torch.distributed as dist
    from xmask3d_b200 import dist as xd, ops, synthetic as syn
    from xmask3d_b200.pipeline import Batch, CorrespondencePipeline
    from xmask3d_b200.voxelizer import Voxelizer
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    n_pts, n_views_total, voxel = 1_000_000, 100, 0.01
    sc = syn.make_scene(2000, n_pts, room=(12.0, 10.0, 3.0))
    mine = list(xd.shard_views(n_views_total, world, rank))
    w2c, depth, rts = [], [], []
    for v in mine:
        vw = syn.make_view(sc, v)
        w2c.append(np.linalg.inv(vw.pose))
        depth.append(vw.depth_mm)
        np.random.seed(5557 + v)
        rt, _ = Voxelizer(voxel_size=voxel, **LOADER_VOX).draw_rigid_transformation()
        rts.append(rt[:3, :4])
    batch = Batch(sc.xyz, np.array([0, n_pts], np.int64), np.zeros(len(mine), np.int64), np.stack(w2c), np.stack(depth),
                  np.stack(rts), syn.scannet_intrinsics())
    pipe = CorrespondencePipeline(batch, args.k, args.c, dev)
    pipe.upload(torch.from_numpy(batch.xyz).pin_memory(), torch.from_numpy(batch.depth_mm.view(np.int16)).pin_memory())
    pr = pipe.project()
    total_vis = int(pr.n_vis.sum().item())
    pipe.set_cap(total_vis)
    masks, mode, _ = make_masks(args, len(mine), dev, 777 + rank)
    g = torch.Generator(device=dev).manual_seed(11)          # the scene's per-point features: same on every rank
    feat = torch.empty((n_pts, args.c), dtype=torch.float32, device=dev)
    for a in range(0, n_pts, 1 << 18):
        feat[a:a + (1 << 18)].normal_(generator=g)

    def step():
        o = pipe.run(masks, feat, mode, feat_per_point=True)
        tot, cnt = xd.allreduce_mask_sums(o["sum"], o["cnt"])
        return o, xd.finalize_mean(tot, cnt), cnt

    for _ in range(max(args.warmup, 3)):
        o, mean, cnt = step()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        o, mean, cnt = step()
    e1.record()
    torch.cuda.synchronize()
    t = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.barrier()
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms = float(t.item()) / args.steps
    # cross-rank consistency: every rank must hold the same scene-level result
    chk = torch.stack([mean.double().sum(), cnt.double().sum()])
    lo, hi = chk.clone(), chk.clone()
    if world > 1:
        dist.all_reduce(lo, op=dist.ReduceOp.MIN)
        dist.all_reduce(hi, op=dist.ReduceOp.MAX)
    if rank == 0:
        line = {"metric": METRIC, "value": n_pts * n_views_total / (ms * 1e-3), "unit": UNIT, "n_gpus": world,
                "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": ms, "higher_is_better": True,
                "scaling": "strong", "vs_baseline": None, "dtype": "f64/u64/f32", "data": "synthetic",
                "config": {"workload": f"configs[3]: 1 scene x {n_pts} pts, 1 cm voxels, {n_views_total} views split over "
                                       f"{world} rank(s), K={args.k} scene-level masks, C={args.c}, one all-reduce of "
                                       f"[K,C+1] float64 per step", "visible_pairs_rank0": total_vis,
                           "pooled_pairs_all_ranks": int(cnt.sum().item()),
                           "ranks_agree": bool(torch.equal(lo, hi))}}
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def main():
    args = parse()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank, world)
        return
    if args.workload == "split_scene":
        run_split_scene(args, rank, world, local_rank)
        return
    run_native(args, rank, world, local_rank)


if __name__ == "__main__":
    main()
