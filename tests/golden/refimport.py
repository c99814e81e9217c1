"""Import the UNMODIFIED reference hot-path functions from /root/reference (build container
only).  Used by make_golden.py and by the optional live cross-checks in tests/ (skipped when
/root/reference is absent, e.g. on the GPU box).

The reference targets Python 3.9 and a stack of heavy dependencies that are not installed
here; two shims make its own source importable without touching it:
  * `collections.Sequence/Iterable` aliases removed in 3.10;
  * a meta-path finder that serves inert stub modules for the missing packages.
"""
import collections
import collections.abc
import importlib.machinery
import os
import sys
import types

REF = "/root/reference"


def available() -> bool:
    return os.path.isdir(os.path.join(REF, "dataset"))


class _Meta(type):
    def __getattr__(cls, n):
        if n.startswith("__"):
            raise AttributeError(n)
        return _Dummy


class _Dummy(metaclass=_Meta):
    def __init__(self, *a, **k):
        pass

    def __call__(self, *a, **k):
        return _Dummy()

    def __getattr__(self, n):
        if n.startswith("__"):
            raise AttributeError(n)
        return _Dummy()


class _Stub(types.ModuleType):
    __path__ = []

    def __getattr__(self, n):
        if n.startswith("__"):
            raise AttributeError(n)
        return _Dummy


class _Finder:
    ROOTS = {"timm", "detectron2", "mask2former", "MinkowskiEngine", "ldm", "open_clip", "diffdist",
             "omegaconf", "fvcore", "iopath", "SharedArray", "imageio", "tensorboardX", "clip"}

    def find_spec(self, name, path=None, target=None):
        if name.split(".")[0] in self.ROOTS:
            return importlib.machinery.ModuleSpec(name, self, is_package=True)

    def create_module(self, spec):
        return _Stub(spec.name)

    def exec_module(self, m):
        pass


_done = False


def load():
    """Returns a namespace with the reference callables."""
    global _done
    if not available():
        raise RuntimeError("reference tree not present")
    if not _done:
        collections.Sequence = collections.abc.Sequence
        collections.Iterable = collections.abc.Iterable
        sys.meta_path.insert(0, _Finder())
        sys.path.insert(0, REF)
        _done = True
    ns = types.SimpleNamespace()
    from dataset.voxelizer import Voxelizer
    from dataset.voxelization_utils import sparse_quantize, fnv_hash_vec, ravel_hash_vec
    from models.utils.mapping_util import getMapping
    from models.utils.fusion_util import PointCloudToImageMapper
    from models.utils.fuser import mask_mapper
    from models.modeling.meta_arch.helper import ensemble_logits_with_labels
    from models.xmask3d import XMASK3d
    from dataset.point_loader import Point3DLoader
    from models.utils.criterion import Criterion
    ns.Criterion = Criterion
    ns.Voxelizer, ns.sparse_quantize, ns.fnv_hash_vec, ns.ravel_hash_vec = \
        Voxelizer, sparse_quantize, fnv_hash_vec, ravel_hash_vec
    ns.getMapping, ns.PointCloudToImageMapper, ns.mask_mapper = getMapping, PointCloudToImageMapper, mask_mapper
    ns.ensemble_logits_with_labels, ns.XMASK3d = ensemble_logits_with_labels, XMASK3d
    ns.Point3DLoader = Point3DLoader
    return ns
