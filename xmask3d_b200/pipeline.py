"""Batched correspondence pipeline: project -> voxelize -> masks-at-points -> pool (+ logits)
for a batch of scenes x views, stream-ordered on the GPU with no host round trip between the
stages (visible counts / voxel counts stay in device memory and drive the next stage's
segment offsets).

This is the path the reference runs per (scene, view) across its loaders and model:
dataset/data_loader_infer.py:161-270 (project, compact, voxelize), models/xmask3d.py:356-367 and
models/utils/criterion.py:83-94, 148-157 (masks at points, pooling), models/xmask3d.py:129-143
(logits).
"""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import Dict, List, Optional, Sequence

import numpy as np
import torch

from . import _lib as L
from . import ops

IMG_W, IMG_H = 320, 240


@dataclass
class Batch:
    """Host description of a batch: scenes (concatenated float32 xyz) and their posed views."""
    xyz: np.ndarray                 # [sum N, 3] float32
    scene_off: np.ndarray           # [S+1] int64 point offsets
    view_scene: np.ndarray          # [V] scene index of every view
    w2c: np.ndarray                 # [V,4,4] float64 world -> camera
    depth_mm: np.ndarray            # [V,240,320] uint16
    rt: np.ndarray                  # [V,3,4] float64 voxelizer matrices (rows 0..2)
    intr: Sequence[float]

    @property
    def n_views(self) -> int:
        return int(self.view_scene.shape[0])

    @property
    def point_views(self) -> int:
        n = np.diff(self.scene_off)
        return int(n[self.view_scene].sum())


@dataclass
class StageTimes:
    names: List[str] = field(default_factory=list)
    events: List = field(default_factory=list)

    def mark(self, name: str):
        ev = torch.cuda.Event(enable_timing=True)
        ev.record()
        self.names.append(name)
        self.events.append(ev)

    def result(self) -> Dict[str, float]:
        out = {}
        for i in range(1, len(self.events)):
            out[self.names[i]] = out.get(self.names[i], 0.0) + self.events[i - 1].elapsed_time(self.events[i])
        return out


class CorrespondencePipeline:
    """Owns the device-resident inputs of a batch and the workspaces; `run` enqueues one pass."""

    def __init__(self, batch: Batch, k: int, c: int, device, cap_vis: Optional[int] = None,
                 cut_bound: int = 10, vis_thres: float = 0.25, depth_scale: float = 1000.0,
                 pairs_per_point: float = 1.0, overlap: bool = True, pool_path: str = "auto", vox_mode: int = 0):
        ops._require_cuda()
        self.batch, self.k, self.c, self.dev = batch, int(k), int(c), device
        self.cut_bound, self.vis_thres, self.depth_scale = cut_bound, vis_thres, depth_scale
        n_pts = np.diff(batch.scene_off)[batch.view_scene]
        self.views, self.out_off = ops.make_views(batch.w2c, batch.intr, batch.scene_off[batch.view_scene], n_pts,
                                                  batch.depth_mm.shape[1:])
        self.n_views = batch.n_views
        self.views_dev = ops.views_to_device(self.views, device)      # keeps project() graph-capturable
        self.graph = None
        self._graph_out = None
        self.overlap = bool(overlap)                       # voxelize || (gather + pool) on two streams
        self.pool_path = pool_path                         # "auto" | "pair_lists" | "rows" | "mma" (ops.pool)
        # vox_mode 2 = XM3D_VOX_FAST_ONLY: when every segment (one view's visible points) is known to stay below 224 k
        # points the shared-memory units always apply and the eight fallback kernels (43 us of launches that return at
        # once) need not be launched.  A batch that needed them raises FLAG_VOX_FALLBACK in vox.status: rerun with 0.
        self.vox_mode = int(vox_mode)
        self._side = torch.cuda.Stream(device=device)
        self._side_status = torch.zeros(1, dtype=torch.int32, device=device)
        self.total_pv = int(self.out_off[-1])
        self.xyz = torch.empty((batch.xyz.shape[0], 3), dtype=torch.float32, device=device)
        self.depth = torch.empty(batch.depth_mm.shape, dtype=torch.int16, device=device)
        self.rt = torch.from_numpy(np.ascontiguousarray(batch.rt)).to(device)
        self.cap_vis = self.total_pv if cap_vis is None else int(cap_vis)
        lib = L.lib()
        self.ws_proj = ops._ws(lib.xm3d_project_ws_bytes(self.n_views, self.total_pv, int(n_pts.max())), device)
        self.ws_vox = ops._ws(lib.xm3d_voxelize_ws_bytes(self.n_views, self.cap_vis), device)
        self.ws_gather = ops._ws(lib.xm3d_gather_ws_bytes(self.n_views, self.k, IMG_H, IMG_W), device)
        self.pairs_per_point = float(pairs_per_point)     # bound on masks per visible point (1 = partition)
        self._size_pool_ws()

    def _size_pool_ws(self):
        self.cap_pairs = int(self.cap_vis * self.pairs_per_point) + 1
        self.ws_pool = ops._ws(L.lib().xm3d_pool_ws_bytes(self.n_views, self.k, self.c, self.cap_vis, self.cap_pairs), self.dev)

    def upload(self, xyz_host: torch.Tensor, depth_host: torch.Tensor):
        """H2D of the loader-side inputs (pinned host tensors -> device), on the current stream."""
        self.xyz.copy_(xyz_host, non_blocking=True)
        self.depth.copy_(depth_host, non_blocking=True)

    def set_cap(self, cap_vis: int):
        self.cap_vis = int(cap_vis)
        self.ws_vox = ops._ws(L.lib().xm3d_voxelize_ws_bytes(self.n_views, self.cap_vis), self.dev)
        self._size_pool_ws()

    def project(self):
        return ops.project_batch(self.xyz, self.views, self.out_off, self.depth, depth_scale=self.depth_scale,
                                 image_dim=(IMG_W, IMG_H), cut_bound=self.cut_bound, vis_thres=self.vis_thres,
                                 cap_vis=self.cap_vis, ws=self.ws_proj, views_dev=self.views_dev)

    def run(self, masks: torch.Tensor, feat: torch.Tensor, mode: str = "ge0.5",
            times: Optional[StageTimes] = None, feat_per_point: bool = False):
        """masks [V,k,240,320] bool/uint8/float32.  feat: [cap_vis, c] float32 with row j = j-th
        visible (view, point) pair in view-major, point-ascending order — or, with
        feat_per_point=True, the per-point features [N, c] of a single-scene batch, gathered through
        the visible-point indices (the pred_3d[inds_reconstruct] pattern of models/xmask3d.py:152).
        Returns a dict of device tensors."""
        overlap = self.overlap and times is None
        self._side_status.zero_()                    # the pair-overflow flag describes THIS pass, not an earlier one
        if times is not None:
            times.mark("start")
        if overlap:
            # fork 1: the masks -> per-pixel membership words pass (HBM bound, independent of the
            # projection) runs on the side stream NEXT TO the projection (instruction bound)
            main = torch.cuda.current_stream()
            side = self._side
            side.wait_stream(main)
            with torch.cuda.stream(side):
                prep = ops.pixel_bits(masks, mode=mode, ws=self.ws_gather)
        pr = self.project()
        if times is not None:
            times.mark("project")
        if overlap:
            # fork 2: masks-at-points + pooling (HBM-bandwidth bound) stay on the side stream while the
            # voxelization (shared-memory units, little bandwidth) runs on the caller's stream
            side.wait_stream(main)
            with torch.cuda.stream(side):
                member, _ = ops.point_bits(prep, pr.rowcol, pr.vis_off, cap=self.cap_vis)
                s, cnt, mean = ops.pool(feat, pr.vis_off, self.k, member=member, cap=self.cap_vis,
                                        cap_pairs=self.cap_pairs,
                                        row_index=(pr.vis_idx if feat_per_point else None), ws=self.ws_pool,
                                        status=self._side_status, path=self.pool_path)
            vox = ops.voxelize_batch(pr.xyz_vis, pr.vis_off, self.rt, cap=self.cap_vis, collate=True, ws=self.ws_vox,
                                     mode=self.vox_mode)
            main.wait_stream(side)                      # join
            for t in (member, s, cnt, mean):
                t.record_stream(main)
            return {"proj": pr, "vox": vox, "member": member, "sum": s, "cnt": cnt, "mean": mean,
                    "pool_status": self._side_status}
        vox = ops.voxelize_batch(pr.xyz_vis, pr.vis_off, self.rt, cap=self.cap_vis, collate=True, ws=self.ws_vox,
                                 mode=self.vox_mode)
        if times is not None:
            times.mark("voxelize")
        member, _ = ops.gather_masks(masks, pr.rowcol, pr.vis_off, mode=mode, cap=self.cap_vis, ws=self.ws_gather)
        if times is not None:
            times.mark("gather")
        s, cnt, mean = ops.pool(feat, pr.vis_off, self.k, member=member, cap=self.cap_vis, cap_pairs=self.cap_pairs,
                                row_index=(pr.vis_idx if feat_per_point else None), ws=self.ws_pool,
                                status=self._side_status, path=self.pool_path)
        if times is not None:
            times.mark("pool")
        return {"proj": pr, "vox": vox, "member": member, "sum": s, "cnt": cnt, "mean": mean,
                "pool_status": self._side_status}

    # -- CUDA graph replay ------------------------------------------------------------------
    def capture(self, masks: torch.Tensor, feat: torch.Tensor, mode: str = "ge0.5", feat_per_point: bool = False):
        """Capture one pass (≈ 30 launches) into a CUDA graph.  Inputs are read from the tensors
        given here (update them in place); the returned dict holds the graph's static outputs.
        Run at least one eager pass first (kernel attributes are set lazily)."""
        torch.cuda.synchronize()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            out = self.run(masks, feat, mode, feat_per_point=feat_per_point)
        self.graph, self._graph_out = g, out
        return out

    def replay(self):
        self.graph.replay()
        return self._graph_out


def algorithmic_bytes(n_pts_per_view: np.ndarray, n_vis: np.ndarray, m_vox: np.ndarray, k: int, c: int,
                      mask_elem_bytes: int = 1) -> Dict[str, int]:
    """Compulsory I/O of every stage (SURVEY.md §8d / BASELINE.md §2), summed over views."""
    N, n, M = n_pts_per_view.astype(np.int64), n_vis.astype(np.int64), m_vox.astype(np.int64)
    hw = IMG_H * IMG_W
    v = len(N)
    project = int((12 * N + N + 8 * n).sum() + v * (2 * hw + 128))
    voxel = int((12 * n + 4 * n + 4 * M + 12 * M).sum() + v * 96)
    pool = int((4 * c * n + 8 * n).sum() + v * (mask_elem_bytes * k * hw + 4 * k * c + 4 * k))
    return {"project": project, "voxelize": voxel, "pool": pool, "total": project + voxel + pool}
