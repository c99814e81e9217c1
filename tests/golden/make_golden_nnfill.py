"""Golden vectors for the step after the path at inference: nearest-seen-neighbour fill and the
per-scene feature maximum.  The reference has no function for either — the statements are inline in
run/infer.py:651-656, 684-694 and models/xmask3d.py:154-159 — so this script EXECUTES THE SAME
STATEMENTS (sklearn KDTree + torch, the reference's own dependencies) on seeded synthetic inputs.

    python tests/golden/make_golden_nnfill.py
"""
import os
import sys

import numpy as np
import torch
from sklearn.neighbors import KDTree

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
from xmask3d_b200 import synthetic as syn  # noqa: E402


def reference_fill(scene_coords, counter, scene_pred):
    """run/infer.py:651-656 and 684-688, verbatim apart from the variable holding the argmax."""
    scene_true = scene_coords[counter != 0]
    scene_false = scene_coords[counter == 0]
    flase_idx = torch.where(counter == 0)[0]
    true_idx = torch.where(counter != 0)[0]
    kdtree = KDTree(scene_true)
    distances, indices = kdtree.query(scene_false, k=1)
    match = true_idx[indices.flatten()]
    scene_pred = scene_pred.clone()
    scene_pred[flase_idx] = scene_pred[match]
    return scene_pred, flase_idx, match, distances.flatten()


def main():
    rng = np.random.default_rng(77)
    sc = syn.make_scene(41, 24_000)
    views = [syn.make_view(sc, v) for v in range(3)]
    # "seen" = visible in one of three synthetic views (as the vote counter would record it)
    from oracle import ref_port
    counter = np.zeros(sc.xyz.shape[0], np.int64)
    mapper = ref_port.getMapping()
    for v in views:
        m = mapper.compute_mapping(v.pose, sc.xyz, v.depth_mm / 1000.0)
        counter += m[:, 2]
    pred = rng.integers(0, 20, sc.xyz.shape[0])
    coords = torch.from_numpy(sc.xyz)
    filled, false_idx, match, dist = reference_fill(coords, torch.from_numpy(counter), torch.from_numpy(pred))
    # per-scene maximum, models/xmask3d.py:154-159
    feat = torch.from_numpy(rng.standard_normal((1200, 40)).astype(np.float32))
    idx = torch.from_numpy(np.sort(rng.integers(0, 4, 1200)))
    rows = []
    for scene_idx in torch.unique(idx):
        single = feat[idx == scene_idx]
        rows.append(torch.max(single, dim=0, keepdim=False)[0])
    segmax = torch.stack(rows)
    path = os.path.join(HERE, "nnfill.npz")
    np.savez_compressed(path, xyz=sc.xyz, counter=counter.astype(np.int32), pred=pred, filled=filled.numpy(),
                        false_idx=false_idx.numpy(), match=match.numpy(), dist=dist,
                        feat=feat.numpy(), idx=idx.numpy(), segmax=segmax.numpy())
    print(f"nnfill: {os.path.getsize(path) / 1024:.0f} KiB, {int((counter == 0).sum())} unseen of {len(counter)}")


if __name__ == "__main__":
    main()
