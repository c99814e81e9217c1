"""On-disk formats that feed the path, and the host side that keeps the GPU fed.

What the reference's loaders read per sample (dataset/data_loader.py:92, 160-171;
dataset/data_loader_infer.py:150-171):

  * scene     `torch.load(<scene>.pth)` -> `(locs float32 [N,3], feats float32 [N,3], labels [N])`
  * pose      `np.loadtxt(<frame>.txt)` -> camera-to-world 4x4 float64
  * depth     `imageio.imread(<frame>.png) / 1000` -> uint16 millimetres -> float64 metres, 240x320

`read_*` return exactly those arrays (the PNG is decoded with OpenCV: imageio is not in the image;
both hand back the stored uint16 samples).  `write_*` produce the same formats (used by the tests
and by anyone who wants to export synthetic scenes).  `FrameBatchLoader` assembles the host `Batch`
of `pipeline.py` from files with a thread pool (np.loadtxt / cv2 / torch.load release the GIL for
the bulk of their work) and `StagingRing` moves it to the device through pinned buffers on a copy
stream, double-buffered, so file IO, H2D copies and the kernels of consecutive batches overlap.
`SceneCache` is the loader's `/dev/shm` scene cache (dataset/point_loader.py:123-181).

The depth image stays uint16 on the way to the GPU: the projection kernel divides by the scale
itself (`depth_scale`, the `/ 1000` above) with the same IEEE division, see csrc/project.cu.
"""
from __future__ import annotations

import os
from concurrent.futures import ThreadPoolExecutor
from dataclasses import dataclass
from typing import List, Optional, Sequence, Tuple

import numpy as np
import torch


# ----------------------------------------------------------------------------- shared-memory scene cache
def preprocess_scene(locs, feats, labels):
    """What the point loader applies to a scene before it is cached or used (dataset/point_loader.py:140-147
    and :182-188): labels -100 -> 255 and uint8; a scalar-zero colour array -> zeros; colours from
    [-1, 1] to [0, 255] (float, `(feats + 1) * 127.5`)."""
    labels = np.array(labels, copy=True)
    labels[labels == -100] = 255
    labels = labels.astype(np.uint8)
    if np.ndim(feats) == 0 and feats == 0:         # scalar (or 0-d array) 0: the scene has no colours
        feats = np.zeros_like(locs)
    feats = (feats + 1.0) * 127.5
    return locs, feats, labels


class SceneCache:
    """The loader's `/dev/shm` scene cache (dataset/point_loader.py:123-181: `sa_create("shm://<dataset>_
    <split>_<identifier>_{locs,feats,labels}_<index>")` once, `SA.attach(...).copy()` per sample), so that the
    DataLoader workers of every rank share ONE decoded copy of each scene.  SharedArray is not in the image;
    the arrays are stored as `.npy` files under the same keys in `root` (default `/dev/shm`) and attached as
    read-only memory maps — the same lifetime and sharing semantics (page cache of a tmpfs), the same key
    scheme, the same preprocessing.  `init` is idempotent (the reference skips it when scene 0 exists)."""

    def __init__(self, dataset_name: str, split: str, identifier: int = 0, root: str = "/dev/shm"):
        self.dataset_name, self.split, self.identifier, self.root = dataset_name, split, int(identifier), root

    def key(self, what: str, index: int) -> str:
        return "%s_%s_%06d_%s_%08d" % (self.dataset_name, self.split, self.identifier, what, index)

    def _path(self, what: str, index: int) -> str:
        return os.path.join(self.root, self.key(what, index) + ".npy")

    def ready(self) -> bool:
        return os.path.exists(self._path("locs", 0))

    def init(self, scene_paths: Sequence[str], workers: Optional[int] = None) -> int:
        """Decode every scene file once (thread pool) and publish it; returns the number of scenes written
        (0 if the cache was already there)."""
        if self.ready():
            return 0

        def one(item):
            i, path = item
            locs, feats, labels = preprocess_scene(*read_scene_pth(path))
            for what, arr in (("locs", locs), ("feats", feats), ("labels", labels)):
                tmp = self._path(what, i) + ".tmp"
                with open(tmp, "wb") as f:
                    np.save(f, np.ascontiguousarray(arr))
                os.replace(tmp, self._path(what, i))            # readers never see a half-written file
            return 1

        order = list(enumerate(scene_paths))
        with ThreadPoolExecutor(max_workers=workers or min(16, os.cpu_count() or 1)) as ex:
            done = sum(ex.map(one, order[1:]))
        return done + one(order[0]) if order else 0             # scene 0 last: `ready()` means complete

    def attach(self, index: int, copy: bool = True):
        """(locs, feats, labels) of scene `index`; `copy=True` is the reference's `SA.attach(...).copy()`."""
        out = []
        for what in ("locs", "feats", "labels"):
            a = np.load(self._path(what, index), mmap_mode="r")
            out.append(np.array(a) if copy else a)
        return tuple(out)

    def delete(self, n_scenes: int) -> None:
        for i in range(n_scenes):
            for what in ("locs", "feats", "labels"):
                try:
                    os.remove(self._path(what, i))
                except FileNotFoundError:
                    pass


# ----------------------------------------------------------------------------- single files
def read_scene_pth(path: str) -> Tuple[np.ndarray, np.ndarray, np.ndarray]:
    """(locs_in, feats_in, labels_in) of dataset/data_loader.py:92, as numpy arrays."""
    locs, feats, labels = torch.load(path, weights_only=False)
    # a scalar 0 colour entry ("no colours", dataset/point_loader.py:143-145) must stay a Python / numpy scalar so
    # that preprocess_scene's np.isscalar test still fires; np.asarray would turn it into a 0-d array
    if np.ndim(feats) == 0:
        feats = feats.item() if hasattr(feats, "item") else feats
    else:
        feats = np.asarray(feats)
    return np.asarray(locs), feats, np.asarray(labels)


def write_scene_pth(path: str, locs: np.ndarray, feats: np.ndarray, labels: np.ndarray) -> None:
    torch.save((np.asarray(locs, np.float32), np.asarray(feats, np.float32), np.asarray(labels)), path)


def read_pose_txt(path: str) -> np.ndarray:
    """np.loadtxt(posepath) (dataset/data_loader.py:162-163): 4x4 float64 camera-to-world."""
    pose = np.loadtxt(path)
    assert pose.shape == (4, 4), f"{path}: expected a 4x4 matrix, got {pose.shape}"
    return pose


def write_pose_txt(path: str, pose: np.ndarray) -> None:
    np.savetxt(path, np.asarray(pose, np.float64), fmt="%.17g")      # round-trips float64 exactly


def read_depth_png(path: str) -> np.ndarray:
    """The raw uint16 millimetre samples of the depth PNG (imageio.imread(...) before the `/ 1000`,
    dataset/data_loader.py:165-168)."""
    import cv2
    img = cv2.imread(path, cv2.IMREAD_UNCHANGED)
    if img is None:
        raise FileNotFoundError(path)
    if img.dtype != np.uint16 or img.ndim != 2:
        raise ValueError(f"{path}: expected a single-channel 16-bit PNG, got {img.dtype} {img.shape}")
    return img


def write_depth_png(path: str, depth_mm: np.ndarray) -> None:
    import cv2
    assert depth_mm.dtype == np.uint16 and depth_mm.ndim == 2
    if not cv2.imwrite(path, depth_mm):
        raise IOError(f"could not write {path}")


def depth_metres(depth_mm: np.ndarray) -> np.ndarray:
    """`imread(png) / 1000`: the float64 image the reference hands to compute_mapping."""
    return depth_mm / 1000


# ----------------------------------------------------------------------------- batches from files
@dataclass
class SceneFrames:
    """One scene file and the frames (pose + depth) projected into it."""
    scene_path: str
    pose_paths: Sequence[str]
    depth_paths: Sequence[str]


class FrameBatchLoader:
    """Builds the host-side `pipeline.Batch` of a list of scenes from their files.

    rt_for(scene_index, frame_index) supplies the voxelizer matrix rows (dataset/voxelizer.py:104-108)
    — drawn by the caller so that the numpy RNG is consumed in the reference's order."""

    def __init__(self, intr: Sequence[float], rt_for, workers: Optional[int] = None):
        self.intr = tuple(float(x) for x in intr)
        self.rt_for = rt_for
        self.pool = ThreadPoolExecutor(max_workers=workers or min(32, (os.cpu_count() or 4)))

    def load(self, scenes: Sequence[SceneFrames]):
        from .pipeline import Batch
        f_scene = [self.pool.submit(read_scene_pth, s.scene_path) for s in scenes]
        f_pose = [[self.pool.submit(read_pose_txt, p) for p in s.pose_paths] for s in scenes]
        f_depth = [[self.pool.submit(read_depth_png, p) for p in s.depth_paths] for s in scenes]
        xyz, off, view_scene, w2c, depth, rts = [], [0], [], [], [], []
        extras = []
        for si, s in enumerate(scenes):
            locs, feats, labels = f_scene[si].result()
            assert locs.ndim == 2 and locs.shape[1] == 3
            xyz.append(np.ascontiguousarray(locs, np.float32))
            extras.append((feats, labels))
            off.append(off[-1] + locs.shape[0])
            assert len(s.pose_paths) == len(s.depth_paths)
            for fi in range(len(s.pose_paths)):
                pose = f_pose[si][fi].result()
                view_scene.append(si)
                w2c.append(np.linalg.inv(pose))                 # models/utils/fusion_util.py:70
                depth.append(f_depth[si][fi].result())
                rts.append(np.asarray(self.rt_for(si, fi), np.float64)[:3, :4])
        batch = Batch(np.concatenate(xyz), np.asarray(off, np.int64), np.asarray(view_scene, np.int64),
                      np.stack(w2c), np.stack(depth), np.stack(rts), self.intr)
        return batch, extras

    def close(self):
        self.pool.shutdown(wait=True)


# ----------------------------------------------------------------------------- host -> device staging
class StagingRing:
    """Pinned host buffers + a copy stream: `stage(batch)` copies the loader-side arrays of a batch
    (scene xyz, uint16 depth images) into the next pinned slot and enqueues the H2D copies; the
    returned event tells the compute stream when the device tensors are complete.  With two or
    more slots the copy of batch i+1 overlaps the kernels of batch i."""

    def __init__(self, max_points: int, max_views: int, depth_hw=(240, 320), device=None, slots: int = 2):
        assert torch.cuda.is_available(), "StagingRing needs a CUDA device"
        self.device = device if device is not None else torch.device("cuda", torch.cuda.current_device())
        self.slots = int(slots)
        self.stream = torch.cuda.Stream(device=self.device)
        h, w = depth_hw
        self.xyz_h = [torch.empty((max_points, 3), dtype=torch.float32).pin_memory() for _ in range(self.slots)]
        self.dep_h = [torch.empty((max_views, h, w), dtype=torch.int16).pin_memory() for _ in range(self.slots)]
        self.xyz_d = [torch.empty((max_points, 3), dtype=torch.float32, device=self.device) for _ in range(self.slots)]
        self.dep_d = [torch.empty((max_views, h, w), dtype=torch.int16, device=self.device) for _ in range(self.slots)]
        self.ready = [torch.cuda.Event() for _ in range(self.slots)]
        self.free = [torch.cuda.Event() for _ in range(self.slots)]
        for e in self.free:
            e.record(torch.cuda.current_stream(self.device))
        self._next = 0

    def stage(self, xyz: np.ndarray, depth_mm: np.ndarray):
        """Returns (xyz_dev [N,3] float32, depth_dev [V,h,w] int16 view of the uint16 bits, ready event, slot)."""
        b = self._next
        self._next = (b + 1) % self.slots
        n, v = xyz.shape[0], depth_mm.shape[0]
        assert n <= self.xyz_h[b].shape[0] and v <= self.dep_h[b].shape[0], "batch larger than the ring's slots"
        self.free[b].synchronize()                       # the kernels that read this slot have finished
        self.xyz_h[b][:n].copy_(torch.from_numpy(np.ascontiguousarray(xyz, np.float32)))
        self.dep_h[b][:v].copy_(torch.from_numpy(np.ascontiguousarray(depth_mm).view(np.int16)))
        with torch.cuda.stream(self.stream):
            self.xyz_d[b][:n].copy_(self.xyz_h[b][:n], non_blocking=True)
            self.dep_d[b][:v].copy_(self.dep_h[b][:v], non_blocking=True)
            self.ready[b].record(self.stream)
        return self.xyz_d[b][:n], self.dep_d[b][:v], self.ready[b], b

    def release(self, slot: int, stream: Optional[torch.cuda.Stream] = None):
        """Call after enqueuing the kernels that consume `slot` (on `stream`, default current)."""
        self.free[slot].record(stream if stream is not None else torch.cuda.current_stream(self.device))
