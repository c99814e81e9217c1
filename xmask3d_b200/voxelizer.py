"""Drop-in for the reference's `dataset/voxelizer.py` (numpy in, numpy out) on libxm3d.

Reference: Voxelizer ctor dataset/voxelizer.py:12-30, get_transformation_matrix :32-58, clip
:60-79, voxelize :81-132.  The rigid transformation (a 4x4 float64, a handful of np.random
draws and three scipy expm calls) stays on the host and consumes np.random in the reference's
order — 3x uniform, 1x shuffle, 1x uniform — so a seeded run draws the same matrices; the
per-point work (transform, floor, min, FNV key, unique / inverse) runs on the GPU.

`voxelize_views` is the batched entry point used by the data loaders' per-view loop
(dataset/data_loader_infer.py:161-270): one call for all views of a scene / batch.
"""
from __future__ import annotations

from collections.abc import Iterable

import numpy as np
import torch
from scipy.linalg import expm, norm

from . import ops


def M(axis, theta):
    """Rotation about `axis` by `theta` (reference :7-8)."""
    return expm(np.cross(np.eye(3), axis / norm(axis) * theta))


class Voxelizer:
    def __init__(self, voxel_size=1, clip_bound=None, use_augmentation=False,
                 scale_augmentation_bound=None, rotation_augmentation_bound=None,
                 translation_augmentation_ratio_bound=None, ignore_label=255):
        self.voxel_size = voxel_size
        self.clip_bound = clip_bound
        self.ignore_label = ignore_label
        self.use_augmentation = use_augmentation
        self.scale_augmentation_bound = scale_augmentation_bound
        self.rotation_augmentation_bound = rotation_augmentation_bound
        self.translation_augmentation_ratio_bound = translation_augmentation_ratio_bound

    # -- host side -------------------------------------------------------------------------
    def get_transformation_matrix(self):
        """(scale matrix, rotation matrix), both 4x4 float64.  np.random is consumed in the
        reference's order (:32-58): one uniform per bounded axis, one shuffle, one uniform."""
        scale_m, rot_m = np.eye(4), np.eye(4)
        bounds = self.rotation_augmentation_bound
        if self.use_augmentation and bounds is not None:
            if not isinstance(bounds, Iterable):
                raise ValueError()
            per_axis = [M(np.eye(3)[a], np.random.uniform(*b) if b is not None else 0)
                        for a, b in enumerate(bounds)]
            np.random.shuffle(per_axis)
            rot_m[:3, :3] = per_axis[0] @ per_axis[1] @ per_axis[2]
        s = 1 / self.voxel_size
        if self.use_augmentation and self.scale_augmentation_bound is not None:
            s *= np.random.uniform(*self.scale_augmentation_bound)
        scale_m[[0, 1, 2], [0, 1, 2]] = s
        return scale_m, rot_m

    def clip(self, coords, center=None, trans_aug_ratio=None):
        """Axis-aligned crop predicate around the (optionally shifted) centre (:60-79)."""
        lo, hi = coords.min(0).astype(float), coords.max(0).astype(float)
        extent = hi - lo
        if center is None:
            center = lo + extent * 0.5
        if trans_aug_ratio is not None:
            center += np.multiply(trans_aug_ratio, extent)
        lim = np.asarray(self.clip_bound, dtype=float)               # [3, 2]: (low, high) per axis
        inside = (coords >= lim[:, 0] + center) & (coords < lim[:, 1] + center)
        return inside.all(axis=1)

    def draw_rigid_transformation(self):
        """(rigid_transformation 4x4, M_r 4x4) exactly as voxelize draws them (:104-108)."""
        M_v, M_r = self.get_transformation_matrix()
        rigid_transformation = M_v
        if self.use_augmentation:
            rigid_transformation = M_r @ rigid_transformation
        return rigid_transformation, M_r

    # -- reference signature ---------------------------------------------------------------
    def voxelize(self, coords, feats, labels, center=None, link=None, return_ind=False):
        assert coords.shape[1] == 3 and coords.shape[0] == feats.shape[0] and coords.shape[0]
        if self.clip_bound is not None:
            trans_aug_ratio = np.zeros(3)
            if self.use_augmentation and self.translation_augmentation_ratio_bound is not None:
                for axis_ind, trans_ratio_bound in enumerate(self.translation_augmentation_ratio_bound):
                    trans_aug_ratio[axis_ind] = np.random.uniform(*trans_ratio_bound)
            clip_inds = self.clip(coords, center, trans_aug_ratio)
            if clip_inds.sum():
                coords, feats = coords[clip_inds], feats[clip_inds]
                if labels is not None:
                    labels = labels[clip_inds]
        rigid_transformation, M_r = self.draw_rigid_transformation()
        res = voxelize_views([coords], [rigid_transformation])
        coords_aug, inds, inds_reconstruct = res[0]
        feats, labels = feats[inds], labels[inds]
        if feats.shape[1] > 6:
            feats[:, 3:6] = feats[:, 3:6] @ (M_r[:3, :3].T)
        if return_ind:
            return coords_aug, feats, labels, np.array(inds_reconstruct), inds
        if link is not None:
            return coords_aug, feats, labels, np.array(inds_reconstruct), link[inds]
        return coords_aug, feats, labels, np.array(inds_reconstruct)


def voxelize_views(coords_list, rigid_transformations, device=None):
    """Batched core of Voxelizer.voxelize: for every (points [n_i,3] float32, 4x4 float64 matrix)
    (float64 coordinates are accepted as well) returns (coords_aug float64 [M_i,3], inds int64 [M_i], inds_reconstruct int64 [n_i]) —
    the values the reference computes at dataset/voxelizer.py:110-122."""
    ops._require_cuda()
    dev = torch.device("cuda", torch.cuda.current_device()) if device is None else device
    n = [int(c.shape[0]) for c in coords_list]
    off = np.concatenate([[0], np.cumsum(n)]).astype(np.int64)
    # float32 (ScanNet .pth) stays float32 — the kernel widens it exactly like numpy's matmul promotion; anything else
    # (float64 after ElasticDistortion, dataset/augmentation.py:171; integer coordinates) goes through float64, the
    # dtype numpy computes `homo_coords @ rigid_transformation.T` in for those inputs (voxelizer.py:110-113)
    if not all(c.dtype == np.float32 for c in coords_list):
        coords_list = [np.asarray(c, np.float64) for c in coords_list]
    xyz = torch.from_numpy(np.ascontiguousarray(np.concatenate(coords_list, 0))).to(dev)
    rt = torch.from_numpy(np.ascontiguousarray(np.stack([np.asarray(r, np.float64)[:3, :4]
                                                         for r in rigid_transformations]))).to(dev)
    u = ops.voxelize_batch(xyz, torch.from_numpy(off).to(dev), rt)
    # ONE host synchronisation: all results are copied asynchronously into cached pinned buffers first
    total = int(off[-1])
    host = _to_host({"status": u.status, "m": u.m, "uoff": u.uniq_off, "first": u.first[:total], "inverse": u.inverse[:total],
                     "voxel": u.voxel_xyz[:total]})
    status = int(host["status"][0])
    if status:
        raise RuntimeError(f"voxelize: device status flags {status:#x} (grid range / sentinel / capacity)")
    m, uoff, first, inverse, voxel = host["m"], host["uoff"], host["first"], host["inverse"], host["voxel"]
    out = []
    for i in range(len(n)):
        a, b = int(uoff[i]), int(uoff[i]) + int(m[i])
        out.append((voxel[a:b].astype(np.float64), first[a:b].astype(np.int64),
                    inverse[off[i]:off[i + 1]].astype(np.int64)))
    return out


_PINNED = {}


def _to_host(tensors):
    """Device tensors -> numpy arrays with one stream synchronisation (pinned staging buffers are cached and grow
    only; the returned arrays are views of them, valid until the next call)."""
    out = {}
    for name, t in tensors.items():
        t = t.contiguous()
        nbytes = t.numel() * t.element_size()
        buf = _PINNED.get(name)
        if buf is None or buf.numel() < nbytes:
            buf = torch.empty(max(nbytes, 1 << 16), dtype=torch.uint8).pin_memory()
            _PINNED[name] = buf
        dst = buf[:nbytes].view(t.dtype).view(t.shape)
        dst.copy_(t, non_blocking=True)
        out[name] = dst
    torch.cuda.current_stream().synchronize()
    return {k: v.numpy() for k, v in out.items()}
