"""ORACLE build step — test infrastructure, NOT product code.

Vendors the reference's own source files for the hot path from /root/reference into
`oracle/_ref/` (git-ignored, NOT gpurun-ignored: it travels to the GPU box like a built .so), so that
`bench.py --impl reference` and the `cpu_baseline` leg can run THE REFERENCE ITSELF on the GPU box's host
cores (`cpu_baseline.kind = "reference"`).  The files are copied unmodified; the two Python-3.9-isms they
contain (`collections.Sequence` / `collections.Iterable`) are bridged at import time by `oracle/refarm.py`.
Nothing under `oracle/_ref/` is ever committed and nothing in `xmask3d_b200/` imports it.

    python oracle/vendor_ref.py            # run by __graft_entry__.build() when /root/reference is present
"""
import hashlib
import json
import os
import shutil
import sys

REF = "/root/reference"
HERE = os.path.dirname(os.path.abspath(__file__))
DST = os.path.join(HERE, "_ref")
FILES = [
    "dataset/voxelizer.py",               # V0-V2  Voxelizer.voxelize, get_transformation_matrix, clip
    "dataset/voxelization_utils.py",      # V3-V4  fnv_hash_vec, ravel_hash_vec, sparse_quantize
    "models/utils/fusion_util.py",        # P0-P1  PointCloudToImageMapper.compute_mapping
    "models/utils/mapping_util.py",       # P0     getMapping
    "models/utils/fuser.py",              # S1     mask_mapper
    "models/modeling/meta_arch/helper.py",  # L2   ensemble_logits_with_labels
]
# The reference's own CALLERS of the path (test-only): tests/test_dropin_integration.py runs ScannetLoaderFull.__getitem__
# unmodified, once with the reference's modules and once with xmask3d_b200's drop-ins injected through sys.modules.
FILES += [
    "dataset/data_loader_infer.py",       # ScannetLoaderFull.__getitem__: the per-view loop around compute_mapping / voxelize
    "dataset/point_loader.py",            # Point3DLoader (base class: builds the Voxelizer)
    "dataset/augmentation.py",            # imported by point_loader
]


def vendor() -> bool:
    if not os.path.isdir(os.path.join(REF, "dataset")):
        return False
    manifest = {}
    for rel in FILES:
        src, dst = os.path.join(REF, rel), os.path.join(DST, rel)
        os.makedirs(os.path.dirname(dst), exist_ok=True)
        shutil.copyfile(src, dst)
        manifest[rel] = hashlib.sha256(open(src, "rb").read()).hexdigest()
        d = os.path.dirname(dst)
        while os.path.abspath(d) != os.path.abspath(DST):     # package markers (the reference's own are empty too)
            open(os.path.join(d, "__init__.py"), "a").close()
            d = os.path.dirname(d)
    with open(os.path.join(DST, "MANIFEST.json"), "w") as f:
        json.dump({"source": REF, "sha256": manifest}, f, indent=1)
    return True


if __name__ == "__main__":
    ok = vendor()
    print("vendored" if ok else "reference tree not present: nothing vendored", file=sys.stderr)
