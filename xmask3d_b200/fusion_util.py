"""Drop-in for the reference's `models/utils/fusion_util.py` on libxm3d.

Reference: make_intrinsic :7-15, adjust_intrinsic :18-33, PointCloudToImageMapper :36-142.
`compute_mapping` keeps the reference signature (numpy in, int64 [N,3] numpy out);
`compute_mapping_batch` projects all views of a scene in one launch sequence — the loop of
dataset/data_loader_infer.py:161-174.
"""
from __future__ import annotations

import math

import numpy as np
import torch

from . import ops


def make_intrinsic(fx, fy, mx, my):
    intrinsic = np.eye(4)
    intrinsic[0][0] = fx
    intrinsic[1][1] = fy
    intrinsic[0][2] = mx
    intrinsic[1][2] = my
    return intrinsic


def adjust_intrinsic(intrinsic, intrinsic_image_dim, image_dim):
    if intrinsic_image_dim == image_dim:
        return intrinsic
    resize_width = int(math.floor(image_dim[1] * float(intrinsic_image_dim[0]) / float(intrinsic_image_dim[1])))
    intrinsic[0, 0] *= float(resize_width) / float(intrinsic_image_dim[0])
    intrinsic[1, 1] *= float(image_dim[1]) / float(intrinsic_image_dim[1])
    intrinsic[0, 2] *= float(image_dim[0] - 1) / float(intrinsic_image_dim[0] - 1)
    intrinsic[1, 2] *= float(image_dim[1] - 1) / float(intrinsic_image_dim[1] - 1)
    return intrinsic


class PointCloudToImageMapper(object):
    def __init__(self, image_dim, visibility_threshold=0.25, cut_bound=0, intrinsics=None):
        self.image_dim = image_dim
        self.vis_thres = visibility_threshold
        self.cut_bound = cut_bound
        self.intrinsics = intrinsics

    def _intr(self, intrinsic):
        if self.intrinsics is not None:
            intrinsic = self.intrinsics
        return (intrinsic[0][0], intrinsic[1][1], intrinsic[0][2], intrinsic[1][2])

    def compute_mapping(self, camera_to_world, coords, depth=None, intrinsic=None):
        """:46-142 — int64 [N,3] rows (pixel row, pixel col, visible)."""
        depths = None if depth is None else [depth]
        return self.compute_mapping_batch([camera_to_world], coords, depths, intrinsic)[0]

    def compute_mapping_batch(self, poses, coords, depths=None, intrinsic=None, device=None):
        """All views of one scene: poses [V] 4x4 camera-to-world, coords [N,3], depths [V] images
        (uint16 raw millimetres, or float64 metres as the loader's `imread(png)/1000`) or None.
        Returns a list of V int64 [N,3] arrays."""
        ops._require_cuda()
        dev = torch.device("cuda", torch.cuda.current_device()) if device is None else device
        coords = np.asarray(coords)
        assert coords.ndim == 2 and coords.shape[1] + 1 == 4, "[!] Shape error"      # :67
        n, v = coords.shape[0], len(poses)
        # float32 coordinates are widened exactly; anything else is narrowed like the loaders' data
        xyz = torch.from_numpy(np.ascontiguousarray(coords, dtype=np.float32)).to(dev)
        if coords.dtype != np.float32 and not np.array_equal(coords, xyz.cpu().numpy().astype(coords.dtype)):
            raise TypeError("compute_mapping: coordinates are not exactly representable in float32")
        w2c = np.stack([np.linalg.inv(np.asarray(p, np.float64)) for p in poses])      # :70
        dshape, dten = None, None
        if depths is not None:
            d0 = np.asarray(depths[0])
            dshape = d0.shape
            if d0.dtype == np.uint16:
                dten = torch.from_numpy(np.ascontiguousarray(np.stack(depths)).view(np.int16)).to(dev)
            else:
                dten = torch.from_numpy(np.ascontiguousarray(np.stack(depths), dtype=np.float64)).to(dev)
        views, out_off = ops.make_views(w2c, self._intr(intrinsic), [0] * v, [n] * v, dshape)
        pr = ops.project_batch(xyz, views, out_off, dten, depth_scale=1000.0, image_dim=self.image_dim,
                               cut_bound=self.cut_bound, vis_thres=self.vis_thres, want_mapping=True,
                               want_compact=False)
        mp = pr.mapping.cpu().numpy()
        return [mp[out_off[i]:out_off[i + 1]] for i in range(v)]
