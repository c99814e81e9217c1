"""Drop-in integration: the reference's OWN caller code, unmodified, running on xmask3d_b200's modules.

`ScannetLoaderFull.__getitem__` (dataset/data_loader_infer.py:113-310 — the per-view loop around
`compute_mapping` and `Voxelizer.voxelize`, :161-270) is executed twice on a small synthetic ScanNet-style dataset
written to disk with `scene_io.write_*`: once with the reference's own `dataset.voxelizer` /
`models.utils.mapping_util`, once with `xmask3d_b200.voxelizer` / `xmask3d_b200.mapping_util` injected under those
names through `sys.modules` (the import swap INTEGRATION.md describes).  Every returned tensor must be identical.

The reference sources are the unmodified copies `oracle/vendor_ref.py` places under `oracle/_ref/` (they travel to
the GPU box; nothing is committed).  Packages the loader imports but the path does not use (imageio, SharedArray)
are served by minimal stand-ins; `torch.load` gets `weights_only=False` (the reference predates that default).
"""
import importlib
import json
import os
import sys
import types

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = os.path.join(ROOT, "oracle", "_ref")
HAVE = os.path.exists(os.path.join(REF, "dataset", "data_loader_infer.py"))
pytestmark = pytest.mark.skipif(not HAVE, reason="oracle/_ref (vendored reference callers) not present: run __graft_entry__.build()")


class Split(dict):
    """category_split: the reference indexes it both as a dict and through attributes (omegaconf)."""
    __getattr__ = dict.__getitem__


def build_dataset(root, n_views=3, n_points=30_000):
    import cv2
    from xmask3d_b200 import scene_io, synthetic as syn
    sc = syn.make_scene(1234, n_points)
    name = "scene0000_00"
    d3 = os.path.join(root, "scannet_3d", "val")
    d2 = os.path.join(root, "scannet_2d", name)
    for sub in ("color", "pose", "depth", "label"):
        os.makedirs(os.path.join(d2, sub), exist_ok=True)
    os.makedirs(d3, exist_ok=True)
    scene_io.write_scene_pth(os.path.join(d3, name + "_vh_clean_2.pth"), sc.xyz, sc.colors, sc.labels)
    rng = np.random.default_rng(0)
    caps = {name: {}}
    for v in range(n_views):
        vw = syn.make_view(sc, v)
        frame = str(v * 20)
        cv2.imwrite(os.path.join(d2, "color", frame + ".jpg"), rng.integers(0, 255, (240, 320, 3), dtype=np.uint8))
        scene_io.write_pose_txt(os.path.join(d2, "pose", frame + ".txt"), vw.pose)
        scene_io.write_depth_png(os.path.join(d2, "depth", frame + ".png"), vw.depth_mm)
        cv2.imwrite(os.path.join(d2, "label", frame + ".png"), rng.integers(0, 25, (240, 320), dtype=np.uint8))
        caps[name][frame] = f"a synthetic room, view {v}"
    cap_path = os.path.join(root, "captions.json")
    json.dump(caps, open(cap_path, "w"))
    return os.path.join(root, "scannet_3d"), os.path.join(root, "scannet_2d"), cap_path


def _stand_ins():
    import cv2
    imageio = types.ModuleType("imageio")
    v2 = types.ModuleType("imageio.v2")

    def imread(path):
        a = cv2.imread(path, cv2.IMREAD_UNCHANGED)
        return a[..., ::-1].copy() if a.ndim == 3 else a
    v2.imread = imread
    imageio.v2 = v2
    return {"imageio": imageio, "imageio.v2": v2, "SharedArray": types.ModuleType("SharedArray")}


def run_loader(dataset, use_dropins: bool):
    """Import the reference loader freshly (with or without the drop-ins under the reference's module names) and
    return what __getitem__(0) yields."""
    import collections
    import collections.abc
    collections.Sequence, collections.Iterable = collections.abc.Sequence, collections.abc.Iterable
    for m in [k for k in sys.modules if k == "dataset" or k.startswith("dataset.") or k == "models" or k.startswith("models.")]:
        del sys.modules[m]
    saved = {k: sys.modules.get(k) for k in ("imageio", "imageio.v2", "SharedArray")}
    sys.modules.update(_stand_ins())
    sys.path.insert(0, REF)
    real_load = torch.load
    torch.load = lambda *a, **k: real_load(*a, **{**k, "weights_only": False})
    try:
        if use_dropins:
            importlib.import_module("dataset")                      # the package shells of the reference
            importlib.import_module("models.utils")
            import xmask3d_b200.mapping_util as mu
            import xmask3d_b200.voxelizer as vx
            sys.modules["dataset.voxelizer"] = vx                   # the swap: same names, same signatures
            sys.modules["models.utils.mapping_util"] = mu
        mod = importlib.import_module("dataset.data_loader_infer")
        if use_dropins:
            assert mod.getMapping.__module__ == "xmask3d_b200.mapping_util"
            assert sys.modules["dataset.point_loader"].Voxelizer.__module__ == "xmask3d_b200.voxelizer"
        d3, d2, cap = dataset
        split = Split(base_category=list(range(15)), novel_category=list(range(15, 19)), ignore_category=[19, 20],
                      all_category=list(range(19)))
        loader = mod.ScannetLoaderFull(d3, d2, label_2d=list(range(1, 20)), category_split=split, caption_path=cap,
                                       voxel_size=0.02, split="val", aug=False, input_color=True)
        np.random.seed(77)
        return loader[0]
    finally:
        torch.load = real_load
        sys.path.remove(REF)
        for k, v in saved.items():
            if v is None:
                sys.modules.pop(k, None)
            else:
                sys.modules[k] = v
        for m in [k for k in sys.modules if k == "dataset" or k.startswith("dataset.") or k == "models" or k.startswith("models.")]:
            del sys.modules[m]


@pytest.fixture(scope="module")
def dataset(tmp_path_factory):
    return build_dataset(str(tmp_path_factory.mktemp("scannet")))


def test_reference_loader_runs_on_written_files(dataset):
    """The files written by scene_io are what the reference's loader reads: its own __getitem__ (reference modules
    only, CPU) yields views with hundreds of visible points."""
    out = run_loader(dataset, use_dropins=False)
    assert out[0] == "scene0000_00" and len(out[3]) >= 2
    assert all(x.shape[0] > 400 for x in out[3])


@pytest.mark.gpu
def test_reference_loader_with_dropins_is_identical(dataset):
    assert torch.cuda.is_available(), "GPU tests need a CUDA device"
    ref = run_loader(dataset, use_dropins=False)
    got = run_loader(dataset, use_dropins=True)
    assert len(ref) == len(got) == 15
    n_cmp = 0
    for a, b in zip(ref, got):
        if isinstance(a, (list, tuple)):
            assert len(a) == len(b)
            for u, v in zip(a, b):
                if torch.is_tensor(u):
                    assert u.dtype == v.dtype and torch.equal(u, v)
                    n_cmp += 1
                else:
                    assert u == v
        elif torch.is_tensor(a):
            assert a.dtype == b.dtype and torch.equal(a, b)
            n_cmp += 1
        else:
            assert a == b
    assert n_cmp > 20
