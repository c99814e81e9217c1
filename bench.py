#!/usr/bin/env python
"""Benchmark of the cross-modal correspondence path (voxelize + project + mask-pool).

    python bench.py --gpus N --steps K --warmup W            # native CUDA path (one rank per GPU)
    python bench.py --impl reference --steps K --warmup W    # the reference's CPU implementation on host cores

Headline workload (BASELINE.json configs[1]): per GPU 8 synthetic ScanNet-sized scenes x 20 posed RGB-D
views, 150k points per scene, 768-d float32 per-point features, 50 masks per view.  A "step" is one pass of
project -> voxelize -> masks-at-points -> pool over the whole batch.  Weak scaling: the 8 x N scenes of the job
are DISTINCT and assigned to ranks by visible-pair count (longest-processing-time greedy, xmask3d_b200/dist.py;
costs from xmask3d_b200/scene_costs.json) — scenes shard with no data-path collective.
The same JSON line carries the other BASELINE configs as extra keys (`configs`): [0] single view through the
drop-in shims (latency), [2] K = 100 masks / 201 classes, [3] one 1 M-point scene split over the ranks with the
NCCL all-reduce of per-mask sums / counts, [4] the 312-scene sweep sharded over the ranks.
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
POOLED_TOL = "pooled sums / means: max|a-b| / max|b| per mask vector <= 1e-5 (vector-wise, SURVEY 7.6), not element-wise"


def parse(argv=None):
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
    p.add_argument("--pool-path", default="auto", choices=["auto", "pair_lists", "rows", "mma"])
    p.add_argument("--cpu-procs", type=int, default=0, help="processes of the CPU legs (0 = one per host cpu)")
    p.add_argument("--no-cpu", action="store_true")
    p.add_argument("--e2e-slots", type=int, default=3, help="staging slots of the pipelined end-to-end loop")
    p.add_argument("--profile-steps", type=int, default=0, help="run only this many plain steps (for ncu)")
    p.add_argument("--replicas", action="store_true",
                   help="every rank processes a copy of scenes 0..7 (round-1 behaviour) instead of distinct, balanced scenes")
    p.add_argument("--no-overlap", action="store_true", help="run the stages back to back on one stream")
    p.add_argument("--no-graph", action="store_true", help="launch every step eagerly instead of replaying a CUDA graph")
    p.add_argument("--extras", default="c0,c2,ov,c3,c4", help="comma list of extra BASELINE configs to run (none = skip)")
    p.add_argument("--sweep-scenes", type=int, default=312)
    p.add_argument("--workload", default="batch", choices=["batch", "split_scene"],
                   help="batch = configs[1] (default, weak scaling); split_scene = configs[3] alone (strong scaling)")
    return p.parse_args(argv)


# ----------------------------------------------------------------------------- synthetic workload (host side)
def sweep_sizes(n_scenes: int):
    """configs[4]: N ~ round(lognormal(ln 150k, 0.5)) clipped to [30k, 500k]."""
    rng = np.random.default_rng(777)
    return np.clip(np.rint(rng.lognormal(np.log(150_000), 0.5, n_scenes)), 30_000, 500_000).astype(np.int64)


_COSTS = None


def scene_cost(seed: int, n_points: int, n_views: int) -> float:
    """Work estimate of a scene: visible (point, view) pairs (xmask3d_b200/scene_costs.json), else N * V."""
    global _COSTS
    if _COSTS is None:
        path = os.path.join(ROOT, "xmask3d_b200", "scene_costs.json")
        _COSTS = json.load(open(path))["costs"] if os.path.exists(path) else {}
    c = _COSTS.get(f"{seed}:{n_points}")
    return float(c) * n_views / 20.0 if c is not None else 0.1 * n_points * n_views


def gen_scene(job):
    """One scene with its posed views (runs in a worker process: numpy only).  job = (seed, n_points, n_views,
    voxel, room, view_ids or None)."""
    seed, n_points, n_views, voxel, room, view_ids = job
    from xmask3d_b200 import synthetic as syn
    from xmask3d_b200.voxelizer import Voxelizer
    sc = syn.make_scene(seed, n_points, room=room)
    w2c, depth, rts, poses = [], [], [], []
    for v in (range(n_views) if view_ids is None else view_ids):
        vw = syn.make_view(sc, v)
        poses.append(vw.pose)
        w2c.append(np.linalg.inv(vw.pose))
        depth.append(vw.depth_mm)
        np.random.seed(5557 + 1000 * (seed - 1000) + v)
        rt, _ = Voxelizer(voxel_size=voxel, **LOADER_VOX).draw_rigid_transformation()
        rts.append(rt[:3, :4])
    return {"seed": seed, "xyz": sc.xyz, "colors": sc.colors, "labels": sc.labels, "poses": np.stack(poses),
            "w2c": np.stack(w2c), "depth": np.stack(depth), "rt": np.stack(rts)}


def make_batch(scene_data):
    from xmask3d_b200 import synthetic as syn
    from xmask3d_b200.pipeline import Batch
    xyz, off, vs, w2c, depth, rts = [], [0], [], [], [], []
    for s, d in enumerate(scene_data):
        xyz.append(d["xyz"])
        off.append(off[-1] + d["xyz"].shape[0])
        vs += [s] * d["w2c"].shape[0]
        w2c.append(d["w2c"]); depth.append(d["depth"]); rts.append(d["rt"])
    return Batch(np.concatenate(xyz), np.array(off, np.int64), np.array(vs, np.int64), np.concatenate(w2c),
                 np.concatenate(depth), np.concatenate(rts), syn.scannet_intrinsics())


def rank_scene_ids(args, rank: int, world: int):
    """Global scene ids (seed = 1000 + id) of this rank: distinct scenes, balanced by visible pairs."""
    if args.replicas:
        return list(range(args.scenes))
    from xmask3d_b200 import dist as xd
    total = world * args.scenes
    costs = [scene_cost(1000 + g, args.points, args.views) for g in range(total)]
    return xd.shard_scenes(costs, world)[rank]


def build_batch(args, rank: int, world: int = 1, pool=None):
    """(Batch, per-scene dicts) of this rank's headline workload."""
    ids = rank_scene_ids(args, rank, world)
    jobs = [(1000 + g, args.points, args.views, args.voxel, None, None) for g in ids]
    data = pool.map(gen_scene, jobs) if pool is not None else [gen_scene(j) for j in jobs]
    return make_batch(data), data


def make_masks(k: int, kind: str, n_views: int, dev, seed: int):
    """partition: K-seed Voronoi label image per view -> bool [V,K,240,320] (the argmax partition of
    models/xmask3d.py:418-435); overlap: smooth random float32 logits, thresholded sigmoid >= 0.5."""
    import torch
    g = torch.Generator(device=dev).manual_seed(seed)
    h, w = 240, 320
    if kind == "partition":
        sy = torch.rand(n_views, k, 1, 1, device=dev, generator=g) * h
        sx = torch.rand(n_views, k, 1, 1, device=dev, generator=g) * w
        yy = torch.arange(h, device=dev).view(1, 1, h, 1).float()
        xx = torch.arange(w, device=dev).view(1, 1, 1, w).float()
        out = torch.empty(n_views, k, h, w, dtype=torch.bool, device=dev)
        for a in range(0, n_views, 8):
            d2 = (yy - sy[a:a + 8]) ** 2 + (xx - sx[a:a + 8]) ** 2
            lab = d2.argmin(1, keepdim=True)
            out[a:a + 8] = lab == torch.arange(k, device=dev).view(1, k, 1, 1)
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


# ----------------------------------------------------------------------------- CPU legs (the reference on host cores)
_CPU_CTX = None      # inherited by the forked workers: (args, scene dicts, masks per (scene, view), feature buffer)


def cpu_context(args, scene_data):
    """Everything the CPU legs need, prepared OUTSIDE the timed region (and before the workers fork)."""
    global _CPU_CTX
    import torch
    from xmask3d_b200 import synthetic as syn
    sets = [torch.from_numpy(syn.make_partition_masks(9000 + i, args.k)).float() for i in range(8)]   # 15 MB each
    masks = {}
    for s, d in enumerate(scene_data):
        for v in range(d["w2c"].shape[0]):
            masks[(s, v)] = sets[(s * 5 + v) % len(sets)]
    fbuf = torch.randn(65_000, args.c, generator=torch.Generator().manual_seed(7))
    _CPU_CTX = (args, scene_data, masks, fbuf)


def _cpu_worker_init():
    import torch
    torch.set_num_threads(1)             # one view per process, like the reference's DataLoader workers
    try:
        from threadpoolctl import threadpool_limits
        threadpool_limits(1)
    except Exception:                    # noqa: BLE001
        pass


def cpu_one_view(job):
    """One (scene, view) unit of the path with the reference's functions (oracle/refarm.py).  Returns N."""
    from oracle import refarm
    s, v = job
    args, scene_data, masks, fbuf = _CPU_CTX
    kind, vox_cls, get_mapping = refarm.load()
    d = scene_data[s]
    mapper = get_mapping()
    refarm.one_view(mapper, vox_cls, dict(voxel_size=args.voxel, **LOADER_VOX), 5557 + 1000 * (d["seed"] - 1000) + v,
                    d["xyz"], d["colors"], d["labels"], d["poses"][v], d["depth"][v] / 1000, masks[(s, v)], fbuf)
    return d["xyz"].shape[0]


def cpu_kind():
    from oracle import refarm
    return refarm.load()[0]


def make_pool(procs: int):
    import multiprocessing as mp
    return mp.get_context("fork").Pool(procs, initializer=_cpu_worker_init)


def host_procs(args):
    n = os.cpu_count() or 1
    try:
        n = len(os.sched_getaffinity(0))
    except Exception:                    # noqa: BLE001
        pass
    return max(1, min(n, args.cpu_procs if args.cpu_procs > 0 else n))


def cpu_baseline_legs(args, pool, procs, jobs, budget_s=25.0):
    """Reference path on the host cores over `jobs` ((scene, view) units), three ways; the best is the baseline:
    one process / one thread, one process / library default threads, a pool of `procs` single-thread processes."""
    import torch
    res = {}
    pv_of = lambda js: sum(_CPU_CTX[1][s]["xyz"].shape[0] for s, _ in js)      # noqa: E731
    cpu_one_view(jobs[0])                                          # warm-up (imports, first-touch)
    sample = jobs[:max(2, min(len(jobs), 6))]
    t0 = time.perf_counter()
    for j in sample:
        cpu_one_view(j)
    dt = time.perf_counter() - t0
    res["default_threads"] = {"value": pv_of(sample) / dt, "threads": torch.get_num_threads()}
    nt = torch.get_num_threads()
    try:
        from threadpoolctl import threadpool_limits
        torch.set_num_threads(1)
        with threadpool_limits(1):
            t0 = time.perf_counter()
            for j in sample:
                cpu_one_view(j)
            dt = time.perf_counter() - t0
        res["one_thread"] = {"value": pv_of(sample) / dt, "threads": 1}
    finally:
        torch.set_num_threads(nt)
    if pool is not None:
        per_view = dt / len(sample)
        n_jobs = int(max(procs, budget_s * procs / max(per_view, 1e-3)))      # ~budget_s seconds of work for the pool
        js = (jobs * (n_jobs // len(jobs) + 1))[:n_jobs]
        pool.map(cpu_one_view, js[:procs])                         # warm the workers
        t0 = time.perf_counter()
        pool.map(cpu_one_view, js, chunksize=1)
        dt = time.perf_counter() - t0
        res["process_pool"] = {"value": pv_of(js) / dt, "threads": procs, "views": len(js), "seconds": dt}
    best = max(res, key=lambda k: res[k]["value"])
    return best, res


def workload_name(args):
    return (f"configs[1]: {args.scenes} scenes x {args.views} views per GPU, {args.points} pts/scene, "
            f"{int(args.voxel * 100)} cm voxels, C={args.c}, K={args.k} {args.masks} masks/view")


def run_reference(args, rank: int, world: int):
    """`--impl reference`: the reference's own CPU implementation of the path (oracle/_ref when vendored, else the
    port) on the host cores, on the SAME config as the native arm: every step is the full 8 scenes x 20 views of
    one GPU's batch, its (scene, view) units spread over one single-thread process per host cpu (the reference
    parallelises this stage with DataLoader workers).  Wall clock per step."""
    if rank != 0:
        return
    procs = host_procs(args)
    gen = make_pool(min(procs, args.scenes))
    _, data = build_batch(args, 0, 1, gen)
    gen.close()
    cpu_context(args, data)
    jobs = [(s, v) for s in range(len(data)) for v in range(data[s]["w2c"].shape[0])]
    pool = make_pool(procs)
    pv_step = sum(data[s]["xyz"].shape[0] for s, _ in jobs)
    for _ in range(max(args.warmup, 1)):
        pool.map(cpu_one_view, jobs[:max(procs, len(jobs) // 4)], chunksize=1)
    tot_t = 0.0
    for _ in range(args.steps):
        t0 = time.perf_counter()
        pool.map(cpu_one_view, jobs, chunksize=1)
        tot_t += time.perf_counter() - t0
    val = pv_step * args.steps / tot_t
    best, legs = cpu_baseline_legs(args, None, procs, jobs[:20])
    pool.close()
    line = {"metric": METRIC, "value": val, "unit": UNIT, "impl": "reference", "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": 1e3 * tot_t / max(args.steps, 1), "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f64/u64/f32", "data": "synthetic",
            "config": {"workload": workload_name(args), "sample": "every step = the full batch of one GPU "
                       f"({len(data)} scenes x {args.views} views = {len(jobs)} (scene, view) units)"},
            "cpu_baseline": {"value": val, "unit": UNIT, "cores": procs, "kind": cpu_kind(),
                             "sample": f"{args.steps} x {len(jobs)} (scene, view) units over {procs} single-thread processes "
                                       f"({os.cpu_count()} host cpus): compute_mapping + compaction + Voxelizer.voxelize + "
                                       "mask gather + per-mask feature[mask].mean(0)",
                             "single_process": legs},
            "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)


# ----------------------------------------------------------------------------- native arm helpers
class StdoutToStderr:
    """NCCL prints its version banner on stdout while the communicator is created; stdout carries exactly one JSON line."""

    def __enter__(self):
        sys.stdout.flush()
        self.saved = os.dup(1)
        os.dup2(2, 1)

    def __exit__(self, *a):
        sys.stdout.flush()
        os.dup2(self.saved, 1)
        os.close(self.saved)


def init_nccl(dev):
    import torch
    import torch.distributed as dist
    with StdoutToStderr():
        dist.init_process_group("nccl", device_id=dev)
        dist.barrier()                                  # creates the communicator (and prints the banner) now
        torch.cuda.synchronize()


class Timer:
    """CUDA-event timing on the current stream with the max over ranks."""

    def __init__(self, dev, world):
        import torch
        self.torch, self.dev, self.world = torch, dev, world
        self.e0, self.e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)

    def barrier(self):
        self.torch.cuda.synchronize()
        if self.world > 1:
            import torch.distributed as dist
            dist.barrier()
            self.torch.cuda.synchronize()

    def time(self, fn, n, after=None):
        """ms for n calls of fn (barrier + synchronize on both sides), max over ranks."""
        self.barrier()
        self.e0.record()
        for i in range(n):
            fn(i)
        if after is not None:
            after()
        self.e1.record()
        self.barrier()
        return self.allmax(self.e0.elapsed_time(self.e1))

    def allmax(self, x):
        t = self.torch.tensor([float(x)], dtype=self.torch.float64, device=self.dev)
        if self.world > 1:
            import torch.distributed as dist
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def allsum(self, x):
        t = self.torch.tensor([float(x)], dtype=self.torch.float64, device=self.dev)
        if self.world > 1:
            import torch.distributed as dist
            dist.all_reduce(t, op=dist.ReduceOp.SUM)
        return float(t.item())


def prepare_pipeline(batch, k, c, dev, mask_kind, seed, overlap=True, pool_path="auto", vox_mode=0):
    """Pipeline + masks + features of a batch; one untimed projection fixes the visible counts (deterministic),
    which size the feature tensor."""
    import torch
    from xmask3d_b200 import ops
    from xmask3d_b200.pipeline import CorrespondencePipeline
    pipe = CorrespondencePipeline(batch, k, c, dev, overlap=overlap, pool_path=pool_path, vox_mode=vox_mode)
    xyz_h = torch.from_numpy(batch.xyz).pin_memory()
    depth_h = torch.from_numpy(batch.depth_mm.view(np.int16)).pin_memory()
    pipe.upload(xyz_h, depth_h)
    pr = pipe.project()
    n_vis = pr.n_vis.cpu().numpy().astype(np.int64)
    total_vis = int(n_vis.sum())
    pipe.set_cap(total_vis)
    masks, mode, mask_bytes = make_masks(k, mask_kind, batch.n_views, dev, seed)
    member0, _ = ops.gather_masks(masks, pr.rowcol, pr.vis_off, mode=mode, cap=total_vis)
    total_pairs = int(ops._popcount32(member0[:total_vis]).sum().item())
    del member0
    pipe.pairs_per_point = total_pairs / max(total_vis, 1)
    pipe._size_pool_ws()
    feat = torch.empty((max(total_vis, 1), c), dtype=torch.float32, device=dev)
    g = torch.Generator(device=dev).manual_seed(seed + 1)
    for a in range(0, total_vis, 1 << 20):
        feat[a:a + (1 << 20)].normal_(generator=g)
    return pipe, masks, mode, mask_bytes, feat, xyz_h, depth_h, n_vis, total_vis, total_pairs


def peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        pk = json.load(open(path))
        return float(pk["hbm_gbs"]), float(pk.get("bf16_tflops", 1678.5)), "MEASURED_PEAKS.json hbm_gbs (of measured)"
    return 6650.0, 1590.0, "B200_PROFILING.md fallback (of fallback)"


# ----------------------------------------------------------------------------- native arm
def run_native(args, rank: int, world: int, local_rank: int):
    import ctypes

    import torch
    import torch.distributed as dist
    procs = host_procs(args)
    extras = set() if args.extras in ("", "none") else set(args.extras.split(","))
    t_setup = time.perf_counter()
    # ---- host-side data first: the worker processes fork BEFORE this process touches CUDA
    gen = make_pool(max(1, min(procs // max(1, min(world, 8)) if world > 1 else procs, 16)))
    batch, scene_data = build_batch(args, rank, world, gen)
    cpu_pool = None
    if rank == 0 and not args.no_cpu:
        cpu_context(args, scene_data[:1])
        cpu_pool = make_pool(procs)                    # inherits the CPU context; used after the GPU legs

    from xmask3d_b200 import _lib as L
    from xmask3d_b200 import ops
    from xmask3d_b200.pipeline import StageTimes, algorithmic_bytes
    assert torch.cuda.is_available(), "the native arm needs a CUDA device (no CPU fallback)"
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        init_nccl(dev)
    T = Timer(dev, world)
    pipe, masks, mode, mask_bytes, feat, xyz_h, depth_h, n_vis, total_vis, total_pairs = prepare_pipeline(
        batch, args.k, args.c, dev, args.masks, 4242 + rank, overlap=not args.no_overlap, pool_path=args.pool_path,
        vox_mode=2)       # 150 k-point scenes: every view's visible points fit the shared-memory units (status checked)
    torch.cuda.synchronize()
    t_setup = time.perf_counter() - t_setup

    if args.profile_steps:
        for _ in range(args.profile_steps):
            pipe.run(masks, feat, mode)
        torch.cuda.synchronize()
        gen.close()
        return

    out = None
    for _ in range(max(args.warmup, 3)):
        out = pipe.run(masks, feat, mode)
    assert int(out["proj"].status.item()) == 0 and int(out["vox"].status.item()) == 0
    assert int(out["pool_status"].item()) == 0
    m_vox = out["vox"].m.cpu().numpy().astype(np.int64)
    total_vox = int(m_vox.sum())

    # the whole step is captured once and replayed as one CUDA graph
    use_graph = not args.no_graph
    l0 = L.lib().xm3d_launch_count()
    if use_graph:
        pipe.capture(masks, feat, mode)
    else:
        pipe.run(masks, feat, mode)
    launches_per_step = L.lib().xm3d_launch_count() - l0

    def do_step(_i=0):
        return pipe.replay() if use_graph else pipe.run(masks, feat, mode)

    # ---- device-resident throughput (`value`): K steps between barriers, CUDA events, max over ranks
    with ClockSampler(local_rank) as clk:
        ms = T.time(do_step, args.steps)
        # per-stage durations: a second pass with events between the stages (one stream, no overlap)
        stage_acc = {}
        ev_a, ev_b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        ev_a.record(); ev_b.record()                       # materialise the cudaEvent_t handles
        torch.cuda.synchronize()
        L.lib().xm3d_set_pool_events(ctypes.c_void_p(ev_a.cuda_event), ctypes.c_void_p(ev_b.cuda_event))
        for _ in range(min(args.steps, 20)):
            st = StageTimes()
            pipe.run(masks, feat, mode, times=st)
            torch.cuda.synchronize()
            for kk, vv in st.result().items():
                stage_acc.setdefault(kk, []).append(vv)
            stage_acc.setdefault("pool_main_kernel", []).append(ev_a.elapsed_time(ev_b))
        L.lib().xm3d_set_pool_events(None, None)
    pv_rank = batch.point_views
    pv_job = T.allsum(pv_rank)
    launches = T.allsum(launches_per_step * args.steps)
    value = pv_job * args.steps / (ms * 1e-3)
    step_ms = ms / args.steps
    stage_ms = {kk: float(np.mean(vv)) for kk, vv in stage_acc.items()}

    # ---- end to end through the host-facing API -----------------------------------------------------------
    # host -> device every step: scene xyz, depth PNG arrays, view records (the loader-side numpy inputs).
    # device -> host every step ("full" variant): x/y labels (int16), inds_reconstruct, first-occurrence indices and
    # voxel coordinates (int16) of the M voxels, offsets, pooled means + counts.  M is data dependent: the copy
    # covers the voxel count of the previous pass + 2 % (checked against uniq_off every step).
    V = batch.n_views
    m_bound = min(total_vis, int(total_vox * 1.02) + 64)
    shapes = {"rowcol16": ((total_vis, 2), torch.int16), "inverse": ((total_vis,), torch.int32),
              "first": ((m_bound,), torch.int32), "voxel16": ((m_bound, 3), torch.int16),
              "vis_off": ((V + 1,), torch.int64), "uniq_off": ((V + 1,), torch.int64),
              "mean": ((V, args.k, args.c), torch.float32), "cnt": ((V, args.k), torch.int32),
              "status": ((4,), torch.int32)}
    small = ("vis_off", "uniq_off", "cnt", "status")          # what a device-side consumer still reads on the host
    NSLOT = max(2, args.e2e_slots)
    s_in, s_out = torch.cuda.Stream(device=dev), torch.cuda.Stream(device=dev)
    comp = torch.cuda.current_stream()
    xyz_st = [torch.empty_like(pipe.xyz) for _ in range(NSLOT)]
    dep_st = [torch.empty_like(pipe.depth) for _ in range(NSLOT)]
    out_st = [{k2: torch.empty(sh, dtype=dt, device=dev) for k2, (sh, dt) in shapes.items()} for _ in range(NSLOT)]
    res_h = [{k2: torch.empty(sh, dtype=dt).pin_memory() for k2, (sh, dt) in shapes.items()} for _ in range(NSLOT)]
    ev = {nm: [torch.cuda.Event() for _ in range(NSLOT)] for nm in ("in_ready", "consumed", "out_ready", "out_free", "done")}
    for b in range(NSLOT):
        for nm in ("consumed", "out_free", "done"):
            ev[nm][b].record(comp)
    h2d = xyz_h.numel() * 4 + depth_h.numel() * 2 + 192 * V
    d2h_full = sum(int(np.prod(sh)) * torch.empty(0, dtype=dt).element_size() for sh, dt in shapes.values())
    d2h_small = sum(int(np.prod(shapes[k2][0])) * torch.empty(0, dtype=shapes[k2][1]).element_size() for k2 in small)

    def stage_outputs(o, ost, keys):
        ost["status"].zero_()
        if "rowcol16" in keys:
            ops.pack_i16(o["proj"].rowcol[:total_vis], o["proj"].vis_off[-1:], out=ost["rowcol16"], status=ost["status"][3:])
            ops.pack_i16(o["vox"].voxel_xyz[:m_bound], o["vox"].uniq_off[-1:], out=ost["voxel16"], status=ost["status"][3:])
            ost["inverse"].copy_(o["vox"].inverse[:total_vis], non_blocking=True)
            ost["first"].copy_(o["vox"].first[:m_bound], non_blocking=True)
            ost["mean"].copy_(o["mean"], non_blocking=True)
        ost["vis_off"].copy_(o["proj"].vis_off, non_blocking=True)
        ost["uniq_off"].copy_(o["vox"].uniq_off, non_blocking=True)
        ost["cnt"].copy_(o["cnt"], non_blocking=True)
        ost["status"][0:1].copy_(o["proj"].status, non_blocking=True)
        ost["status"][1:2].copy_(o["vox"].status, non_blocking=True)
        ost["status"][2:3].copy_(o["pool_status"], non_blocking=True)

    # One CUDA graph PER SLOT: it reads the slot's input buffers (the copy-in stream writes them) and leaves every output,
    # the int16-packed maps included, in the slot's own tensors (the copy-out stream reads them) — no staging copies on the
    # compute stream.  Without graphs (--no-graph) the step runs eagerly and stages through out_st.
    slot_graphs = None
    if use_graph:
        slot_graphs = []
        keep_xyz, keep_depth = pipe.xyz, pipe.depth
        for b in range(NSLOT):
            xyz_st[b].copy_(keep_xyz)
            dep_st[b].copy_(keep_depth)
            pipe.xyz, pipe.depth = xyz_st[b], dep_st[b]
            torch.cuda.synchronize()
            g_b = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g_b):
                o = pipe.run(masks, feat, mode)
                st = torch.zeros(4, dtype=torch.int32, device=dev)
                r16 = ops.pack_i16(o["proj"].rowcol[:total_vis], o["proj"].vis_off[-1:], status=st[3:])
                v16 = ops.pack_i16(o["vox"].voxel_xyz[:m_bound], o["vox"].uniq_off[-1:], status=st[3:])
                st[0:1].copy_(o["proj"].status)
                st[1:2].copy_(o["vox"].status)
                st[2:3].copy_(o["pool_status"])
            slot_graphs.append((g_b, {"rowcol16": r16, "voxel16": v16, "inverse": o["vox"].inverse[:total_vis],
                                      "first": o["vox"].first[:m_bound], "mean": o["mean"], "cnt": o["cnt"],
                                      "vis_off": o["proj"].vis_off, "uniq_off": o["vox"].uniq_off, "status": st}))
        pipe.xyz, pipe.depth = keep_xyz, keep_depth
        for b in range(NSLOT):
            for k2, (sh, dt) in shapes.items():
                t_ = slot_graphs[b][1][k2]
                assert tuple(t_.shape) == tuple(sh) and t_.dtype == dt, (k2, t_.shape, t_.dtype)

    def e2e_pipelined(i, keys, compute=True, copy_in=True):
        """Step i of the pipelined loop: copy-in stream -> the slot's input buffers -> compute stream (the slot's graph:
        the whole step + int16 packing) -> the slot's outputs -> copy-out stream -> pinned host; the host waits for the
        oldest step in flight."""
        b = i % NSLOT
        if copy_in:
            with torch.cuda.stream(s_in):
                s_in.wait_event(ev["consumed"][b])
                xyz_st[b].copy_(xyz_h, non_blocking=True)
                dep_st[b].copy_(depth_h, non_blocking=True)
                ev["in_ready"][b].record(s_in)
            comp.wait_event(ev["in_ready"][b])
        src = out_st[b]
        if slot_graphs is not None:
            src = slot_graphs[b][1]
            comp.wait_event(ev["out_free"][b])
            if compute:
                slot_graphs[b][0].replay()
            ev["consumed"][b].record(comp)
        else:
            if compute:
                pipe.xyz.copy_(xyz_st[b], non_blocking=True)
                pipe.depth.copy_(dep_st[b], non_blocking=True)
            ev["consumed"][b].record(comp)
            comp.wait_event(ev["out_free"][b])
            if compute:
                stage_outputs(do_step(), out_st[b], keys)
        ev["out_ready"][b].record(comp)
        with torch.cuda.stream(s_out):
            s_out.wait_event(ev["out_ready"][b])
            for k2 in keys:
                res_h[b][k2].copy_(src[k2], non_blocking=True)
            ev["out_free"][b].record(s_out)
            ev["done"][b].record(s_out)
        if i >= NSLOT - 1:
            ev["done"][(i - NSLOT + 1) % NSLOT].synchronize()      # the caller consumes the oldest step in flight

    def drain():
        comp.wait_stream(s_out)
        comp.wait_stream(s_in)

    n_e2e = max(4, min(args.steps, 50))
    full_keys = tuple(shapes)

    def run_e2e(keys, compute=True, copy_in=True, copy_out=True):
        ks = keys if copy_out else ()
        for i in range(2 * NSLOT):
            e2e_pipelined(i, ks, compute, copy_in)
        t = T.time(lambda i: e2e_pipelined(i, ks, compute, copy_in), n_e2e, after=drain)
        return t / n_e2e

    e2e_ms = run_e2e(full_keys)
    last = res_h[(n_e2e - 1) % NSLOT]
    assert torch.equal(last["cnt"], out["cnt"].cpu()), "the host did not receive what the device computed"
    assert int(last["uniq_off"][-1]) <= m_bound and not bool(last["status"].any())
    dc_ms = run_e2e(small)                                         # device-side consumer: maps stay in HBM
    ceil_ms = run_e2e(full_keys, compute=False)                    # the same copies, no kernels: the PCIe ceiling
    ceil_in_ms = run_e2e((), compute=False, copy_out=False)        # host -> device alone
    ceil_out_ms = run_e2e(full_keys, compute=False, copy_in=False)  # device -> host alone

    def e2e_serial(_i=0):
        pipe.upload(xyz_h, depth_h)
        stage_outputs(do_step(), out_st[0], full_keys)
        for k2 in full_keys:
            res_h[0][k2].copy_(out_st[0][k2], non_blocking=True)
        torch.cuda.current_stream().synchronize()
    e2e_serial()
    serial_ms = T.time(e2e_serial, 10) / 10
    rate = lambda ms_: pv_job / (ms_ * 1e-3)                                   # noqa: E731
    gbs = lambda nbytes, ms_: world * nbytes / (ms_ * 1e-3) / 1e9              # noqa: E731
    e2e = {"value": rate(e2e_ms), "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h_full,
           "ms_per_step": e2e_ms, "unpipelined_value": rate(serial_ms), "frac_of_pcie_ceiling": ceil_ms / e2e_ms,
           "note": f"{NSLOT}-slot pipeline: copy-in / compute / copy-out streams overlap across steps, one CUDA graph per slot "
                   "(it reads the slot's input buffers and writes the slot's outputs).  host->device per step "
                   "and rank: scene xyz, depth images, view records (the loader's numpy inputs); device->host: x/y labels "
                   "(int16), inds_reconstruct (int32), first-occurrence indices (int32) and voxel coordinates (int16) of the "
                   "M voxels, offsets, pooled means (float32) and counts.  Per-point features and 2D masks are consumed on "
                   "the device, where the reference's API produces them (3D UNet / mask head outputs)"}
    e2e_dc = {"value": rate(dc_ms), "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h_small,
              "ms_per_step": dc_ms,
              "note": "what the real train / infer loop needs (run/train.py:483-485): x_label / y_label / inds_reconstruct / "
                      "coords / pooled features never leave the GPU; the host reads back offsets, counts and status flags"}
    pcie = {"ms_per_step_both": ceil_ms, "value_both": rate(ceil_ms), "ms_h2d_alone": ceil_in_ms,
            "ms_d2h_alone": ceil_out_ms, "aggregate_h2d_gbs_alone": gbs(h2d, ceil_in_ms),
            "aggregate_d2h_gbs_alone": gbs(d2h_full, ceil_out_ms),
            "aggregate_gbs_both": gbs(h2d + d2h_full, ceil_ms), "unit": UNIT,
            "note": "the e2e loop's pinned-memory copies of the same byte counts on all ranks at once, without the kernels: "
                    "`value_both` is the end-to-end rate no kernel speed-up can exceed on this host"}
    del xyz_st, dep_st, out_st, res_h

    # ---- extra BASELINE configs (all ranks take part: some hold collectives) ---------------------------------
    cfgs = {}
    peak, peak_tf, peak_src = peaks()
    if "c2" in extras:
        cfgs["configs[2]"] = extra_k100(args, batch, dev, T, rank, peak, pv_job, feat_shared=feat)
    if "ov" in extras and args.masks == "partition":
        del masks
        masks = None
        torch.cuda.empty_cache()
        cfgs["overlapping_masks"] = extra_overlap(args, batch, dev, T, rank, peak, pv_job, feat_shared=feat)
    del masks
    torch.cuda.empty_cache()
    if "c3" in extras:
        del feat, pipe
        torch.cuda.empty_cache()
        cfgs["configs[3]"] = split_scene_leg(args, rank, world, dev, gen, T, steps=max(3, min(args.steps, 10)))
        feat = pipe = None
    if "c4" in extras:
        if feat is not None:
            del feat, pipe
            feat = pipe = None
            torch.cuda.empty_cache()
        cfgs["configs[4]"] = sweep_leg(args, rank, world, dev, gen, T, peak)
    gen.close()

    if rank != 0:
        if world > 1:
            dist.barrier()
            dist.destroy_process_group()
        return

    if "c0" in extras:
        try:
            cfgs["configs[0]"] = single_view_leg(args, scene_data[0], dev)
        except Exception as e:                      # noqa: BLE001
            cfgs["configs[0]"] = {"error": f"{type(e).__name__}: {e}"[:300]}
    logits_info = logits_leg(args, dev, total_vis, peak, peak_tf)
    mask_prep_info = mask_prep_leg(args, V, dev)

    # ---- roofline of the dominant kernel (pool) and of the whole step
    n_pts_view = np.diff(batch.scene_off)[batch.view_scene]
    alg = algorithmic_bytes(n_pts_view, n_vis, m_vox, args.k, args.c, mask_bytes)
    # algorithmic bytes of the dominant kernel: every visible point's C-float row read once + its
    # 4-byte row index (pair lists) / membership words, plus the [K,C] sums it produces per view
    overlap_masks = total_pairs > total_vis + 1
    pool_bytes = int((4 * args.c + 4) * total_vis + V * 4 * args.k * args.c)
    pool_ms = stage_ms.get("pool_main_kernel", float("nan"))
    achieved = pool_bytes / (pool_ms * 1e-3) / 1e9
    kernel = ("pool_mma2_kernel (tcgen05: 0/1 membership operand in TMEM, tf32 hi + bf16 lo feature tiles via TMA)" if overlap_masks and args.pool_path in ("auto", "mma")
              else "pool_rows_kernel (point-major)" if overlap_masks and args.pool_path == "rows"
              else "pool_sum_kernel<4> (pair lists)")
    # dram__bytes_read+write of this kernel from THIS ROUND's `ncu --set full` capture of the same default workload
    traffic, traffic_src = None, None
    tname = "r02_pool_mma2_kernel_ncu_full.json" if overlap_masks else "r02_pool_sum_kernel_ncu_full.json"
    default_wl = (args.scenes, args.views, args.points, args.k, args.c, world) == (8, 20, 150_000, 50, 768, 1)
    tpath = os.path.join(ROOT, "profiles", tname)
    if default_wl and os.path.exists(tpath):
        traffic = json.load(open(tpath)).get("_summary", {}).get("traffic_bytes_per_launch")
        traffic_src = "profiles/" + tname
    roof = {"bound": "hbm", "kernel": kernel + "; CUDA events recorded around this launch alone, inside the step",
            "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": traffic,
            "traffic_source": traffic_src, "bytes_per_launch": pool_bytes, "ms_per_launch": pool_ms, "peak_source": peak_src}
    pipe_gbs = alg["total"] / (step_ms * 1e-3) / 1e9
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
        "ms_per_step": step_ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64/u64/f32",
        "data": "synthetic",
        "config": {"workload": workload_name(args), "point_views_per_step_all_gpus": int(pv_job),
                   "point_views_rank0": pv_rank, "visible_pairs_rank0": total_vis, "mask_memberships_rank0": total_pairs,
                   "voxels_rank0": total_vox, "scene_ids_rank0": rank_scene_ids(args, 0, world),
                   "cache": "inputs larger than L2 (features %.1f GB per GPU)" % (total_vis * args.c * 4 / 1e9),
                   "parallelism": f"scenes sharded over {world} rank(s), no data-path collective; " +
                                  ("every rank processes a copy of scenes 0..7" if args.replicas else
                                   f"{world * args.scenes} distinct scenes balanced by visible pairs (LPT)"),
                   "launch": ("one CUDA graph replay per step" if use_graph else "eager launches") +
                             ("; voxelize overlapped with gather+pool on a second stream" if not args.no_overlap else ""),
                   "kernels_per_step": int(launches_per_step), "setup_s": round(t_setup, 1),
                   "parity_tolerance": POOLED_TOL},
        "roofline": roof,
        "pipeline_roofline": {"algorithmic_bytes_per_step": alg, "achieved": pipe_gbs, "peak": peak, "unit": "GB/s",
                              "frac": pipe_gbs / peak},
        "stage_ms": stage_ms,
        "e2e": e2e, "e2e_device_consumer": e2e_dc, "pcie_ceiling": pcie,
        "configs": cfgs,
        "logits": logits_info,
        "mask_prep": mask_prep_info,
        "gpu_launches": int(launches),
        "clocks": clk.summary(),
    }
    if cpu_pool is not None:
        jobs = [(0, v) for v in range(scene_data[0]["w2c"].shape[0])]
        best, legs = cpu_baseline_legs(args, cpu_pool, procs, jobs, budget_s=15.0)
        cpu_pool.close()
        line["cpu_baseline"] = {"value": legs[best]["value"], "unit": UNIT, "cores": legs[best]["threads"], "kind": cpu_kind(),
                                "best_of": best, "legs": legs,
                                "sample": f"scene {rank_scene_ids(args, 0, world)[0]} x {len(jobs)} views (the full scene), repeated "
                                          f"over the pool for ~15 s: the reference's compute_mapping + compaction + "
                                          f"Voxelizer.voxelize + mask gather + per-mask feature[mask].mean(0) "
                                          f"({os.cpu_count()} host cpus)"}
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


# ----------------------------------------------------------------------------- extra legs
def extra_k100(args, batch, dev, T, rank, peak, pv_job, feat_shared):
    """configs[2]: ScanNet200 mask count — the same batch with K = 100 masks per view."""
    import torch
    from xmask3d_b200.pipeline import algorithmic_bytes
    pipe, masks, mode, mask_bytes, feat, _, _, n_vis, total_vis, total_pairs = prepare_pipeline_shared(
        batch, 100, args.c, dev, 5252 + rank, feat_shared)
    for _ in range(3):
        out = pipe.run(masks, feat, mode)
    assert int(out["proj"].status.item()) == 0 and int(out["vox"].status.item()) == 0 and int(out["pool_status"].item()) == 0
    pipe.capture(masks, feat, mode)
    n = max(5, min(args.steps, 20))
    ms = T.time(lambda i: pipe.replay(), n) / n
    m_vox = out["vox"].m.cpu().numpy().astype(np.int64)
    n_pts_view = np.diff(batch.scene_off)[batch.view_scene]
    alg = algorithmic_bytes(n_pts_view, n_vis, m_vox, 100, args.c, mask_bytes)
    res = {"workload": "configs[2]: the headline batch with K=100 partition masks/view (ScanNet200 B170N30); the 201-class "
                       "logits contraction is under `logits`",
           "ms_per_step": ms, "value": pv_job / (ms * 1e-3), "unit": UNIT,
           "pipeline_roofline_frac": alg["total"] / (ms * 1e-3) / 1e9 / peak}
    del pipe, masks
    torch.cuda.empty_cache()
    return res


def extra_overlap(args, batch, dev, T, rank, peak, pv_job, feat_shared):
    """The criterion path's masks (models/utils/criterion.py:83-85): raw thresholded float32 predictions that OVERLAP
    (about 7.7 memberships per visible point) — pooled by the tensor-core kernel, every feature row read once."""
    import ctypes
    import torch
    from xmask3d_b200 import _lib as L, ops
    from xmask3d_b200.pipeline import CorrespondencePipeline, StageTimes, algorithmic_bytes
    pipe = CorrespondencePipeline(batch, args.k, args.c, dev, vox_mode=2)
    pipe.upload(torch.from_numpy(batch.xyz), torch.from_numpy(batch.depth_mm.view(np.int16)))
    pr = pipe.project()
    n_vis = pr.n_vis.cpu().numpy().astype(np.int64)
    total_vis = int(n_vis.sum())
    pipe.set_cap(total_vis)
    masks, mode, mask_bytes = make_masks(args.k, "overlap", batch.n_views, dev, 7171 + rank)
    member0, _ = ops.gather_masks(masks, pr.rowcol, pr.vis_off, mode=mode, cap=total_vis)
    total_pairs = int(ops._popcount32(member0[:total_vis]).sum().item())
    del member0
    pipe.pairs_per_point = total_pairs / max(total_vis, 1)
    pipe._size_pool_ws()
    feat = feat_shared
    for _ in range(3):
        out = pipe.run(masks, feat, mode)
    assert int(out["proj"].status.item()) == 0 and int(out["vox"].status.item()) == 0 and int(out["pool_status"].item()) == 0
    ev_a, ev_b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev_a.record(); ev_b.record()
    torch.cuda.synchronize()
    L.lib().xm3d_set_pool_events(ctypes.c_void_p(ev_a.cuda_event), ctypes.c_void_p(ev_b.cuda_event))
    kms = []
    for _ in range(5):
        pipe.run(masks, feat, mode, times=StageTimes())
        torch.cuda.synchronize()
        kms.append(ev_a.elapsed_time(ev_b))
    L.lib().xm3d_set_pool_events(None, None)
    pipe.capture(masks, feat, mode)
    n = max(5, min(args.steps, 20))
    ms = T.time(lambda i: pipe.replay(), n) / n
    m_vox = out["vox"].m.cpu().numpy().astype(np.int64)
    n_pts_view = np.diff(batch.scene_off)[batch.view_scene]
    alg = algorithmic_bytes(n_pts_view, n_vis, m_vox, args.k, args.c, mask_bytes)
    feat_bytes = 4 * args.c * total_vis
    k_ms = float(np.mean(kms))
    res = {"workload": f"the headline batch with OVERLAPPING float32 mask logits (sigmoid >= 0.5; {total_pairs / max(total_vis, 1):.1f} "
                       "memberships per visible point, the criterion path): pooling on the tensor cores (pool_mma2_kernel)",
           "ms_per_step": ms, "value": pv_job / (ms * 1e-3), "unit": UNIT,
           "pipeline_roofline_frac": alg["total"] / (ms * 1e-3) / 1e9 / peak,
           "pool_kernel_ms": k_ms, "pool_kernel_frac_of_hbm_peak": feat_bytes / (k_ms * 1e-3) / 1e9 / peak,
           "memberships": total_pairs}
    del pipe, masks
    torch.cuda.empty_cache()
    return res


def prepare_pipeline_shared(batch, k, c, dev, seed, feat):
    """prepare_pipeline on a batch whose feature tensor already exists."""
    import torch
    from xmask3d_b200 import ops
    from xmask3d_b200.pipeline import CorrespondencePipeline
    pipe = CorrespondencePipeline(batch, k, c, dev, vox_mode=2)
    pipe.upload(torch.from_numpy(batch.xyz), torch.from_numpy(batch.depth_mm.view(np.int16)))
    pr = pipe.project()
    n_vis = pr.n_vis.cpu().numpy().astype(np.int64)
    total_vis = int(n_vis.sum())
    pipe.set_cap(total_vis)
    masks, mode, mask_bytes = make_masks(k, "partition", batch.n_views, dev, seed)
    return pipe, masks, mode, mask_bytes, feat, None, None, n_vis, total_vis, total_vis


def split_scene_leg(args, rank, world, dev, gen, T, steps, n_pts=1_000_000, n_views_total=100, voxel=0.01):
    """configs[3]: one dense scene (1M points, 1 cm voxels), 100 views block-partitioned over the ranks; every rank
    pools its views under a scene-level set of K masks and ONE all-reduce(SUM) of the packed [K, C+1] float64
    sums / counts yields the scene-level mask features on every rank.  Strong scaling."""
    import torch
    import torch.distributed as dist
    from xmask3d_b200 import dist as xd
    from xmask3d_b200.pipeline import CorrespondencePipeline
    mine = list(xd.shard_views(n_views_total, world, rank))
    # the views are generated in parallel chunks (each worker rebuilds the scene: cheaper than shipping it)
    chunks = [mine[i::4] for i in range(4) if mine[i::4]]
    parts = gen.map(gen_scene, [(2000, n_pts, 0, voxel, (12.0, 10.0, 3.0), ch) for ch in chunks]) if chunks else []
    order = np.argsort(np.concatenate([np.array(ch) for ch in chunks])) if chunks else np.zeros(0, int)
    cat = lambda key: np.concatenate([p[key] for p in parts])[order]            # noqa: E731
    data = [{"xyz": parts[0]["xyz"], "w2c": cat("w2c"), "depth": cat("depth"), "rt": cat("rt")}]
    batch = make_batch(data)
    pipe = CorrespondencePipeline(batch, args.k, args.c, dev)
    pipe.upload(torch.from_numpy(batch.xyz).pin_memory(), torch.from_numpy(batch.depth_mm.view(np.int16)).pin_memory())
    pr = pipe.project()
    total_vis = int(pr.n_vis.sum().item())
    pipe.set_cap(total_vis)
    masks, mode, _ = make_masks(args.k, "partition", len(mine), dev, 777 + rank)
    g = torch.Generator(device=dev).manual_seed(11)          # the scene's per-point features: same on every rank
    feat = torch.empty((n_pts, args.c), dtype=torch.float32, device=dev)
    for a in range(0, n_pts, 1 << 18):
        feat[a:a + (1 << 18)].normal_(generator=g)

    def step(_i=0):
        o = pipe.run(masks, feat, mode, feat_per_point=True)
        tot, cnt = xd.allreduce_mask_sums(o["sum"], o["cnt"])
        return xd.finalize_mean(tot, cnt), cnt

    for _ in range(3):
        mean, cnt = step()
    ms = T.time(step, steps) / steps
    mean, cnt = step()
    chk = torch.stack([mean.double().sum(), cnt.double().sum()])
    lo, hi = chk.clone(), chk.clone()
    if world > 1:
        dist.all_reduce(lo, op=dist.ReduceOp.MIN)
        dist.all_reduce(hi, op=dist.ReduceOp.MAX)
    res = {"workload": f"configs[3]: 1 scene x {n_pts} pts, 1 cm voxels, {n_views_total} views split over {world} rank(s), "
                       f"K={args.k} scene-level masks, C={args.c}, one NCCL all-reduce of [K,C+1] float64 per step",
           "ms_per_step": ms, "value": n_pts * n_views_total / (ms * 1e-3), "unit": UNIT, "scaling": "strong",
           "visible_pairs_rank0": total_vis, "pooled_pairs_all_ranks": int(cnt.sum().item()),
           "ranks_agree": bool(torch.equal(lo, hi))}
    del pipe, feat, masks
    torch.cuda.empty_cache()
    return res


def sweep_leg(args, rank, world, dev, gen, T, peak, views=20, batch_scenes=8):
    """configs[4]: throughput sweep over 312 synthetic ScanNet-val-shaped scenes (lognormal sizes 30k..500k points,
    20 views each) sharded over the ranks by visible pairs; every rank walks its scenes in batches of 8, each batch
    one pipeline pass timed on the device (best of 3 after a warm-up); the sweep time is the sum over batches, max
    over ranks.  Host generation of the next batch overlaps the GPU work on the current one."""
    import torch
    from xmask3d_b200 import _lib as L
    from xmask3d_b200 import dist as xd
    from xmask3d_b200.pipeline import algorithmic_bytes
    sizes = sweep_sizes(args.sweep_scenes)
    costs = [scene_cost(1000 + s, int(n), views) for s, n in enumerate(sizes)]
    mine = xd.shard_scenes(costs, world)[rank]
    mine = sorted(mine, key=lambda s: int(sizes[s]))                  # homogeneous batches
    batches = [mine[i:i + batch_scenes] for i in range(0, len(mine), batch_scenes)]
    jobs = lambda b: [(1000 + s, int(sizes[s]), views, args.voxel, None, None) for s in b]     # noqa: E731
    pending = gen.map_async(gen_scene, jobs(batches[0])) if batches else None
    tot_ms, tot_pv, tot_bytes, t_wall = 0.0, 0, 0, time.perf_counter()
    n_fallback = 0
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    for bi, b in enumerate(batches):
        data = pending.get()
        pending = gen.map_async(gen_scene, jobs(batches[bi + 1])) if bi + 1 < len(batches) else None
        batch = make_batch(data)
        pipe, masks, mode, mask_bytes, feat, _, _, n_vis, total_vis, _ = prepare_pipeline(batch, args.k, args.c, dev,
                                                                                         "partition", 31 + bi, vox_mode=2)
        out = pipe.run(masks, feat, mode)
        if int(out["vox"].status.item()) & L.FLAG_VOX_FALLBACK:    # this batch needs the multi-kernel voxel path
            pipe.vox_mode = 0
            n_fallback += 1
            out = pipe.run(masks, feat, mode)
        if not args.no_graph:
            pipe.capture(masks, feat, mode)                 # one CUDA graph per batch, like the headline step
        best = None
        for _ in range(3):
            torch.cuda.synchronize()
            e0.record()
            out = pipe.run(masks, feat, mode) if args.no_graph else pipe.replay()
            e1.record()
            torch.cuda.synchronize()
            t = e0.elapsed_time(e1)
            best = t if best is None else min(best, t)
        assert int(out["proj"].status.item()) == 0 and int(out["vox"].status.item()) == 0 and int(out["pool_status"].item()) == 0
        m_vox = out["vox"].m.cpu().numpy().astype(np.int64)
        alg = algorithmic_bytes(np.diff(batch.scene_off)[batch.view_scene], n_vis, m_vox, args.k, args.c, mask_bytes)
        tot_ms += best
        tot_pv += batch.point_views
        tot_bytes += alg["total"]
        del pipe, masks, feat, out
    torch.cuda.empty_cache()
    t_wall = time.perf_counter() - t_wall
    ms = T.allmax(tot_ms)
    pv = T.allsum(tot_pv)
    nbytes = T.allsum(tot_bytes)
    return {"workload": f"configs[4]: {args.sweep_scenes} scenes (N ~ lognormal(150k, 0.5) in [30k, 500k]) x {views} views, "
                        f"K={args.k}, C={args.c}, sharded over {world} rank(s) by visible pairs (LPT), batches of {batch_scenes} scenes",
            "value": pv / (ms * 1e-3), "unit": UNIT, "device_ms_total_max_rank": ms, "point_views": int(pv),
            "scenes_rank0": len(mine), "batches_on_multi_kernel_voxel_path_rank0": n_fallback, "pipeline_roofline_frac": nbytes / max(world, 1) / (ms * 1e-3) / 1e9 / peak,
            "wall_s_rank0_incl_host_generation": round(t_wall, 1),
            "note": "device time of the pipeline passes (one CUDA graph per batch, one pass per batch, best of 3); scene / feature / mask "
                    "generation is outside the timed region"}


def single_view_leg(args, d, dev):
    """configs[0]: ONE posed view of one 150k-point scene through the per-call drop-in shims (numpy in, numpy out,
    like the reference's loader calls them) — latency, with the reference's CPU functions beside it."""
    import torch
    from oracle import refarm
    from xmask3d_b200 import synthetic as syn
    from xmask3d_b200.logits import cal_pred_logits
    from xmask3d_b200.mapping_util import getMapping
    from xmask3d_b200.pooling import masked_mean_pool
    from xmask3d_b200.voxelizer import Voxelizer
    kind, ref_vox, ref_get_mapping = refarm.load()
    from oracle import ref_port as P
    xyz, pose, depth_m = d["xyz"], d["poses"][0], d["depth"][0] / 1000
    mask = torch.from_numpy(syn.make_partition_masks(9000, args.k)).float()

    def med(fn, n=5):
        fn(); fn()
        ts = []
        for _ in range(n):
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            r = fn()
            torch.cuda.synchronize()
            ts.append(time.perf_counter() - t0)
        return float(np.median(ts)) * 1e3, r

    mapper, ref_mapper = getMapping(), ref_get_mapping()
    t_proj, mapping = med(lambda: mapper.compute_mapping(pose, xyz, depth_m))
    tr_proj, ref_mapping = med(lambda: ref_mapper.compute_mapping(pose, xyz, depth_m), 3)
    assert np.array_equal(mapping, ref_mapping)
    vis, x_label, y_label = P.compact_mapping(mapping)
    locs, col, lab = xyz[vis], d["colors"][vis], d["labels"][vis]

    def vox(cls):
        np.random.seed(5557)
        return cls(voxel_size=args.voxel, **LOADER_VOX).voxelize(locs, col, lab)
    t_vox, r = med(lambda: vox(Voxelizer))
    tr_vox, rr = med(lambda: vox(ref_vox), 3)
    assert all(np.array_equal(a, b) for a, b in zip(r, rr))
    feat = torch.randn(locs.shape[0], args.c, generator=torch.Generator().manual_seed(1))
    feat_d, mask_d = feat.to(dev), mask.to(dev)[None]
    xl, yl = torch.from_numpy(x_label), torch.from_numpy(y_label)
    t_pool, _ = med(lambda: masked_mean_pool([feat_d], [xl], [yl], mask_d, mode="ge0.5"))
    tr_pool, _ = med(lambda: P.masked_mean_pool(feat, P.gather_masks(mask, xl, yl, "ge0.5")), 3)
    me, te, ne = syn.make_embeddings(1, 1, args.k, 20)
    o = {"mask_embed": torch.from_numpy(me).to(dev), "text_embed": torch.from_numpy(te).to(dev),
         "null_embed": torch.from_numpy(ne).to(dev), "labels": [[str(i)] for i in range(19)], "logit_scale": 1 / 0.07}
    t_log, _ = med(lambda: cal_pred_logits(o))
    oc = {k2: (v2.cpu() if torch.is_tensor(v2) else v2) for k2, v2 in o.items()}
    oc["logit_scale"] = torch.tensor(1 / 0.07)
    tr_log, _ = med(lambda: P.cal_pred_logits(oc), 3)
    tot, rtot = t_proj + t_vox + t_pool + t_log, tr_proj + tr_vox + tr_pool + tr_log
    n = xyz.shape[0]
    return {"workload": f"configs[0]: 1 scene x {n} pts, 1 view, 2 cm voxels, K={args.k}, 20 classes (B15N4) through the "
                        "per-call drop-in shims (numpy in / numpy out, host<->device copies inside)",
            "latency_ms": {"compute_mapping": t_proj, "Voxelizer.voxelize": t_vox, "masked_mean_pool": t_pool,
                           "cal_pred_logits": t_log, "total": tot},
            "reference_cpu_ms": {"compute_mapping": tr_proj, "Voxelizer.voxelize": tr_vox, "masked_mean_pool": tr_pool,
                                 "cal_pred_logits": tr_log, "total": rtot, "kind": kind},
            "value": n / (tot * 1e-3), "reference_value": n / (rtot * 1e-3), "unit": UNIT,
            "outputs_identical": True, "visible_points": int(vis.sum())}


def logits_leg(args, dev, total_vis, peak_hbm, peak_tf):
    """Stage 4 (text logits, the only dense contraction): timed apart, it is not part of the metric."""
    import torch
    info = {}
    try:
        from xmask3d_b200 import ops
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        gl = torch.Generator(device=dev).manual_seed(1)
        for name, rows, t in (("configs[1] B15N4 160x50 masks x 20 classes", 160 * 50, 20),
                              ("configs[2] B170N30 160x100 masks x 201 classes", 160 * 100, 201)):
            me = torch.randn(rows, args.c, device=dev, generator=gl)
            te = torch.randn(t - 1, args.c, device=dev, generator=gl)
            ne = torch.randn(1, args.c, device=dev, generator=gl)
            for _ in range(3):
                ops.logits(me, te, ne, [1] * (t - 1), 1 / 0.07)
            torch.cuda.synchronize()
            e0.record()
            for _ in range(20):
                ops.logits(me, te, ne, [1] * (t - 1), 1 / 0.07)
            e1.record()
            torch.cuda.synchronize()
            ms_l = e0.elapsed_time(e1) / 20
            fl = 2.0 * rows * args.c * t
            info[name] = {"ms": ms_l, "useful_tflops": fl / (ms_l * 1e-3) / 1e12,
                          "mma_tflops_3xtf32": 3 * fl / (ms_l * 1e-3) / 1e12,
                          "frac_of_bf16_peak": 3 * fl / (ms_l * 1e-3) / 1e12 / peak_tf,
                          "note": "prep + tcgen05 GEMM + epilogue; <= 5 GFLOP, launch/latency bound by construction"}
        n = min(total_vis, 2_400_000)
        feat = torch.randn(n, args.c, device=dev, generator=gl)
        for name, t, ens in (("per-point argmax, 19 classes", 19, False), ("per-point argmax, 200 classes", 200, False),
                             ("per-point fused-stream ensemble + argmax, 19 classes", 19, True)):
            te = torch.randn(t, args.c, device=dev, generator=gl)
            kw = {}
            if ens:
                kf = 40
                kw = dict(binary=(torch.rand(n, device=dev, generator=gl) > 0.5).float(),
                          is_base=torch.arange(t, device=dev) < 15,
                          mask_label=torch.randint(-1, kf, (n,), device=dev, generator=gl, dtype=torch.int32),
                          mask_probs=torch.rand(kf, t, device=dev, generator=gl).softmax(-1), base_ratio=0.65, novel_ratio=0.35)
            for _ in range(2):
                ops.point_logits(feat, te, 1 / 0.07, want_logits=False, **kw)
            torch.cuda.synchronize()
            e0.record()
            for _ in range(5):
                ops.point_logits(feat, te, 1 / 0.07, want_logits=False, **kw)
            e1.record()
            torch.cuda.synchronize()
            ms_l = e0.elapsed_time(e1) / 5
            fl = 2.0 * n * args.c * t
            info[name] = {"ms": ms_l, "rows": n, "feature_gbs": 4.0 * args.c * n / (ms_l * 1e-3) / 1e9,
                          "frac_of_hbm_peak": 4.0 * args.c * n / (ms_l * 1e-3) / 1e9 / peak_hbm,
                          "useful_tflops": fl / (ms_l * 1e-3) / 1e12}
    except Exception as e:                      # noqa: BLE001
        info["error"] = f"{type(e).__name__}: {e}"[:200]
    return info


def mask_prep_leg(args, n_views, dev):
    """Mask preparation: low-resolution mask-head logits -> membership words / partition labels
    of every view of the batch, one fused pass (timed apart, not part of the metric)."""
    import torch
    try:
        import torch.nn.functional as F
        from xmask3d_b200 import ops
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        gm = torch.Generator(device=dev).manual_seed(2)
        lg = torch.randn(n_views, args.k, 16, 16, device=dev, generator=gm)
        lg = F.interpolate(lg, size=(128, 128), mode="bicubic", align_corners=False).contiguous() * 3.0
        scm = torch.rand(n_views, args.k, device=dev, generator=gm) + 0.05

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
            up = F.interpolate(lg, size=(240, 320), mode="bilinear", align_corners=False)
            sg = up.sigmoid()
            return sg > 0.5, (scm.view(n_views, args.k, 1, 1) * sg).argmax(1)
        return {"shape": f"{n_views} views x {args.k} masks, 128x128 -> 240x320",
                "bits_ms": _time(lambda: ops.mask_prep(lg, (240, 320), mode="sigmoid_gt0.5", want_bits=True), 10),
                "partition_ms": _time(lambda: ops.mask_prep(lg, (240, 320), scores=scm, want_bits=False, want_partition=True), 10),
                "torch_cuda_ops_ms": _time(_torch_ops, 3),
                "note": "fused bilinear upsample + sigmoid threshold (+ score-weighted argmax partition and areas); "
                        "torch_cuda_ops = interpolate + sigmoid + mul + argmax + compare on the same GPU"}
    except Exception as e:                      # noqa: BLE001
        return {"error": f"{type(e).__name__}: {e}"[:200]}


def run_split_scene(args, rank: int, world: int, local_rank: int):
    """`--workload split_scene`: configs[3] alone as the JSON line (strong scaling)."""
    import torch
    import torch.distributed as dist
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    gen = make_pool(4)
    if world > 1:
        init_nccl(dev)
    T = Timer(dev, world)
    r = split_scene_leg(args, rank, world, dev, gen, T, steps=args.steps)
    gen.close()
    if rank == 0:
        line = {"metric": METRIC, "value": r["value"], "unit": UNIT, "n_gpus": world, "steps": args.steps,
                "warmup": 3, "ms_per_step": r["ms_per_step"], "higher_is_better": True, "scaling": "strong",
                "vs_baseline": None, "dtype": "f64/u64/f32", "data": "synthetic", "config": r}
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
