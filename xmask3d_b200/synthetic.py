"""Deterministic synthetic inputs for the cross-modal correspondence path.

Nothing here is on the product path: it only manufactures scenes, posed RGB-D
views, 2D masks and embeddings of the shapes SURVEY.md §8(d) prescribes, so that
tests, `bench.py` and `__graft_entry__.smoke()` all draw identical inputs from a
seed (there is no dataset and no network on the build or GPU boxes).

Conventions follow the reference's data loader:
  * scene coordinates are float32 `[N,3]` metres, z up (ScanNet `.pth`,
    /root/reference/dataset/data_loader.py:92);
  * a pose is a camera-to-world 4x4 float64 with +z forward, +x right, +y down
    (what `np.loadtxt(pose.txt)` yields, dataset/data_loader_infer.py:166);
  * depth is a uint16 millimetre image 240x320 (the PNG), which the loader turns
    into float64 metres by `/ 1000` (dataset/data_loader_infer.py:168-171);
  * the camera is the fixed ScanNet one of models/utils/mapping_util.py:10-39.
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field
from typing import List, Optional, Tuple

import numpy as np

IMG_W, IMG_H = 320, 240


def scannet_intrinsics() -> Tuple[float, float, float, float]:
    """(fx, fy, cx, cy) after the 640x480 -> 320x240 rescale.

    Same arithmetic as `adjust_intrinsic` (models/utils/fusion_util.py:18-33)
    applied to the constants in `getMapping` (models/utils/mapping_util.py:16-22),
    evaluated in the same order so the doubles are bit-identical.
    """
    fx = fy = 577.870605
    mx, my = 319.5, 239.5
    resize_width = int(math.floor(IMG_H * float(640) / float(480)))
    fx *= float(resize_width) / float(640)
    fy *= float(IMG_H) / float(480)
    mx *= float(IMG_W - 1) / float(640 - 1)
    my *= float(IMG_H - 1) / float(480 - 1)
    return fx, fy, mx, my


@dataclass
class Scene:
    seed: int
    xyz: np.ndarray            # [N,3] float32
    colors: np.ndarray         # [N,3] float32 in (-1,1)
    labels: np.ndarray         # [N]   float64 class ids
    room: Tuple[float, float, float]
    boxes: np.ndarray          # [nb,6] float64  (x0,y0,z0,x1,y1,z1) furniture


@dataclass
class View:
    pose: np.ndarray           # [4,4] float64 camera-to-world
    depth_mm: np.ndarray       # [240,320] uint16

    @property
    def depth_m(self) -> np.ndarray:
        return self.depth_mm / 1000


def _sample_rect(rng, n, origin, e1, e2):
    u = rng.random((n, 1))
    v = rng.random((n, 1))
    return origin[None, :] + u * e1[None, :] + v * e2[None, :]


def make_scene(seed: int, n_points: int = 150_000, num_classes: int = 19,
               room: Optional[Tuple[float, float, float]] = None) -> Scene:
    """A box room with furniture; points live on surfaces (floor 30 %, walls 40 %,
    furniture 30 %) so voxel collisions and occlusions look like ScanNet's."""
    rng = np.random.default_rng(seed)
    if room is None:
        lx, ly, lz = rng.uniform(4, 9), rng.uniform(3, 7), rng.uniform(2.4, 3.0)
    else:
        lx, ly, lz = room
    nb = int(rng.integers(8, 16))
    boxes = np.zeros((nb, 6))
    for b in range(nb):
        sx, sy, sz = rng.uniform(0.4, 1.6), rng.uniform(0.4, 1.6), rng.uniform(0.4, 1.5)
        x0, y0 = rng.uniform(0.1, lx - sx - 0.1), rng.uniform(0.1, ly - sy - 0.1)
        boxes[b] = (x0, y0, 0.0, x0 + sx, y0 + sy, sz)

    n_floor = int(0.3 * n_points)
    n_wall = int(0.4 * n_points)
    n_furn = n_points - n_floor - n_wall
    parts = [_sample_rect(rng, n_floor, np.zeros(3), np.array([lx, 0, 0.]), np.array([0, ly, 0.]))]
    walls = [
        (np.array([0, 0, 0.]), np.array([lx, 0, 0.]), np.array([0, 0, lz])),
        (np.array([0, ly, 0.]), np.array([lx, 0, 0.]), np.array([0, 0, lz])),
        (np.array([0, 0, 0.]), np.array([0, ly, 0.]), np.array([0, 0, lz])),
        (np.array([lx, 0, 0.]), np.array([0, ly, 0.]), np.array([0, 0, lz])),
    ]
    wlen = np.array([lx, lx, ly, ly])
    wcnt = np.floor(n_wall * wlen / wlen.sum()).astype(int)
    wcnt[0] += n_wall - wcnt.sum()
    for (o, e1, e2), c in zip(walls, wcnt):
        parts.append(_sample_rect(rng, int(c), o, e1, e2))
    # furniture: top + four sides of every box, by area
    faces = []
    for (x0, y0, z0, x1, y1, z1) in boxes:
        faces.append((np.array([x0, y0, z1]), np.array([x1 - x0, 0, 0.]), np.array([0, y1 - y0, 0.])))
        faces.append((np.array([x0, y0, z0]), np.array([x1 - x0, 0, 0.]), np.array([0, 0, z1 - z0])))
        faces.append((np.array([x0, y1, z0]), np.array([x1 - x0, 0, 0.]), np.array([0, 0, z1 - z0])))
        faces.append((np.array([x0, y0, z0]), np.array([0, y1 - y0, 0.]), np.array([0, 0, z1 - z0])))
        faces.append((np.array([x1, y0, z0]), np.array([0, y1 - y0, 0.]), np.array([0, 0, z1 - z0])))
    area = np.array([np.linalg.norm(np.cross(e1, e2)) for _, e1, e2 in faces])
    fcnt = np.floor(n_furn * area / area.sum()).astype(int)
    fcnt[0] += n_furn - fcnt.sum()
    for (o, e1, e2), c in zip(faces, fcnt):
        parts.append(_sample_rect(rng, int(c), o, e1, e2))
    xyz = np.concatenate(parts, 0)
    xyz = xyz[rng.permutation(xyz.shape[0])]
    xyz = xyz + rng.normal(0.0, 0.002, xyz.shape)
    colors = rng.uniform(-1, 1, (n_points, 3)).astype(np.float32)
    labels = rng.integers(0, num_classes, n_points).astype(np.float64)
    return Scene(seed, xyz.astype(np.float32), colors, labels, (lx, ly, lz), boxes)


def make_pose(rng, room) -> np.ndarray:
    lx, ly, lz = room
    # ScanNet trajectories hug the walls and look across the room: stand in the outer
    # band of the floor plan and face the centre +-35 degrees.
    ang = rng.uniform(0, 2 * np.pi)
    rad = rng.uniform(0.55, 0.9)
    cam = np.array([0.5 * lx * (1 + rad * np.cos(ang)), 0.5 * ly * (1 + rad * np.sin(ang)),
                    rng.uniform(1.2, 1.8)])
    yaw = np.arctan2(0.5 * ly - cam[1], 0.5 * lx - cam[0]) + rng.uniform(-0.6, 0.6)
    pitch = np.deg2rad(rng.uniform(-30.0, 0.0))
    f = np.array([np.cos(yaw) * np.cos(pitch), np.sin(yaw) * np.cos(pitch), np.sin(pitch)])
    r = np.cross(f, np.array([0, 0, 1.0]))
    r /= np.linalg.norm(r)
    d = np.cross(f, r)
    pose = np.eye(4)
    pose[:3, 0], pose[:3, 1], pose[:3, 2], pose[:3, 3] = r, d, f, cam
    return pose


def raycast_depth_mm(scene: Scene, pose: np.ndarray) -> np.ndarray:
    """Exact z-depth of the room + furniture seen from `pose`, quantised like a
    ScanNet depth PNG (uint16 millimetres, 0 = nothing hit)."""
    fx, fy, cx, cy = scannet_intrinsics()
    u, v = np.meshgrid(np.arange(IMG_W, dtype=np.float64), np.arange(IMG_H, dtype=np.float64))
    dcam = np.stack([(u - cx) / fx, (v - cy) / fy, np.ones_like(u)], -1).reshape(-1, 3)
    dw = dcam @ pose[:3, :3].T           # un-normalised: parameter t == camera-space z
    o = pose[:3, 3]
    lx, ly, lz = scene.room
    with np.errstate(divide="ignore", invalid="ignore"):
        inv = 1.0 / dw
        # inside the room the exit distance of the room box is the wall hit
        t_hi = np.where(dw > 0, (np.array([lx, ly, lz]) - o) * inv, (np.zeros(3) - o) * inv)
        t_hi = np.where(dw == 0, np.inf, t_hi)
        t = t_hi.min(1)
        for (x0, y0, z0, x1, y1, z1) in scene.boxes:
            ta = (np.array([x0, y0, z0]) - o) * inv
            tb = (np.array([x1, y1, z1]) - o) * inv
            tn = np.nanmax(np.minimum(ta, tb), 1)
            tf = np.nanmin(np.maximum(ta, tb), 1)
            hit = (tn <= tf) & (tn > 1e-6)
            t = np.where(hit & (tn < t), tn, t)
    t = np.where(np.isfinite(t) & (t > 0), t, 0.0)
    mm = np.clip(np.rint(t * 1000.0), 0, 65535).astype(np.uint16)
    return mm.reshape(IMG_H, IMG_W)


def make_view(scene: Scene, view_idx: int) -> View:
    rng = np.random.default_rng([scene.seed, view_idx])
    pose = make_pose(rng, scene.room)
    return View(pose, raycast_depth_mm(scene, pose))


def make_partition_masks(seed, k: int, h: int = IMG_H, w: int = IMG_W) -> np.ndarray:
    """K bool masks `[K,h,w]` that tile the image (nearest-seed Voronoi) — the shape of
    the argmax partition built at models/xmask3d.py:418-435."""
    rng = np.random.default_rng(seed)
    sy, sx = rng.uniform(0, h, k), rng.uniform(0, w, k)
    yy, xx = np.mgrid[0:h, 0:w]
    d2 = (yy[None] - sy[:, None, None]) ** 2 + (xx[None] - sx[:, None, None]) ** 2
    lab = d2.argmin(0)
    return lab[None] == np.arange(k)[:, None, None]


def make_overlap_logits(seed, k: int, h: int = IMG_H, w: int = IMG_W, lo: int = 128) -> np.ndarray:
    """K float32 mask logits `[K,h,w]`: smooth random fields at `lo`x`lo`, bilinearly
    up-sampled like `F.interpolate(..., align_corners=False)` (models/xmask3d.py:326-331).
    Masks overlap; each covers roughly 2-30 % of the image. No exact zeros."""
    import torch
    import torch.nn.functional as F
    g = torch.Generator().manual_seed(int(seed))
    z = torch.randn(k, 1, lo, lo, generator=g)
    ker = torch.arange(-12, 13, dtype=torch.float32)
    ker = torch.exp(-0.5 * (ker / 5.0) ** 2)
    ker = (ker / ker.sum())
    z = F.conv2d(z, ker.view(1, 1, -1, 1), padding=(12, 0))
    z = F.conv2d(z, ker.view(1, 1, 1, -1), padding=(0, 12))
    z = z / z.flatten(1).std(1).view(k, 1, 1, 1)
    cover = torch.rand(k, generator=g) * 0.28 + 0.02
    flat = z.flatten(1)
    thr = torch.stack([torch.quantile(flat[i], 1.0 - cover[i]) for i in range(k)])
    z = (z - thr.view(k, 1, 1, 1)) * 4.0
    up = F.interpolate(z.transpose(0, 1), size=(h, w), mode="bilinear", align_corners=False)[0]
    up = torch.where(up == 0, torch.full_like(up, 1e-3), up)
    return up.numpy().astype(np.float32)


def make_embeddings(seed, b: int, k: int, t: int, c: int = 768):
    """(mask_embed [b,k,c], text_embed [t-1,c], null_embed [1,c]) float32 ~ N(0,1)."""
    rng = np.random.default_rng(seed)
    return (rng.standard_normal((b, k, c), dtype=np.float32),
            rng.standard_normal((t - 1, c), dtype=np.float32),
            rng.standard_normal((1, c), dtype=np.float32))


@dataclass
class Workload:
    scenes: List[Scene] = field(default_factory=list)
    views: List[List[View]] = field(default_factory=list)   # per scene

    @property
    def point_views(self) -> int:
        return sum(s.xyz.shape[0] * len(v) for s, v in zip(self.scenes, self.views))


def make_workload(n_scenes: int, n_views: int, n_points: int = 150_000, seed0: int = 1000,
                  room=None) -> Workload:
    wl = Workload()
    for s in range(n_scenes):
        sc = make_scene(seed0 + s, n_points, room=room)
        wl.scenes.append(sc)
        wl.views.append([make_view(sc, v) for v in range(n_views)])
    return wl
