"""Golden vectors for Voxelizer.clip and for voxelize() with a clip_bound (dataset/voxelizer.py:60-102):
the crop predicate, the extra np.random draws of the translation augmentation and the voxelization of
the cropped cloud, produced by the reference's own Voxelizer (imported unmodified through refimport).

    python tests/golden/make_golden_clip.py
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))

import refimport  # noqa: E402
from make_golden_params import vox_kwargs  # noqa: E402
from xmask3d_b200 import synthetic as syn  # noqa: E402

R = refimport.load()
CLIP = ((-1.6, 1.4), (-1.2, 1.9), (-0.4, 2.0))


def main():
    sc = syn.make_scene(51, 6000)
    rng = np.random.default_rng(4)
    colors = rng.uniform(-1, 1, (6000, 3)).astype(np.float32)
    labels = rng.integers(0, 20, 6000).astype(np.float64)
    kw = vox_kwargs(0.05)
    kw["clip_bound"] = CLIP
    vox = R.Voxelizer(**kw)
    inside_default = vox.clip(sc.xyz)
    centre = np.array([2.0, 1.5, 1.0])
    inside_centre = vox.clip(sc.xyz, center=centre.copy(), trans_aug_ratio=np.array([0.1, -0.05, 0.0]))
    np.random.seed(991)
    grid, feats, lab, inv, inds = vox.voxelize(sc.xyz, colors.copy(), labels.copy(), return_ind=True)
    path = os.path.join(HERE, "clip.npz")
    np.savez_compressed(path, xyz=sc.xyz, colors=colors, labels=labels, clip=np.array(CLIP),
                        inside_default=inside_default, inside_centre=inside_centre, centre=centre,
                        grid=grid, feats=feats, lab=lab, inv=inv, inds=inds)
    print(f"clip: {os.path.getsize(path) / 1024:.0f} KiB, {int(inside_default.sum())} / 6000 inside, {len(inds)} voxels")


if __name__ == "__main__":
    main()
