"""CPU tests of host-side logic that is plain torch on the kernels' outputs (no GPU needed)."""
import numpy as np
import torch

from xmask3d_b200 import ops


def test_accept_views_matches_the_loaders_rule():
    """dataset/data_loader.py:190-199: keep a frame iff 400 < sum(mask) < 65000 and more than 10 of its
    visible points have a non-ignored label — batched over views on the projection's compaction outputs."""
    rng = np.random.default_rng(0)
    n_scene = [5000, 3000]
    scene_off = np.array([0, 5000, 8000])
    view_scene = np.array([0, 0, 1, 1, 1, 0])
    labels = rng.integers(0, 21, 8000)
    labels[labels == 20] = 255
    vis = [rng.random(n_scene[view_scene[v]]) < p for v, p in enumerate([0.2, 0.05, 0.5, 0.0, 0.3, 0.9])]
    labels[5000:8000][vis[4]] = 255                         # view 4: every visible point is ignored
    n_vis = np.array([m.sum() for m in vis])
    vis_off = np.concatenate([[0], np.cumsum(n_vis)])
    idx = np.concatenate([np.nonzero(m)[0] for m in vis]).astype(np.int32)
    vis_idx = np.zeros(len(idx) + 50, np.int32)             # capacity larger than the visible count
    vis_idx[:len(idx)] = idx
    pr = ops.Projection(None, torch.from_numpy(n_vis.astype(np.int32)), torch.from_numpy(vis_off),
                        torch.from_numpy(vis_idx), None, None, None, None, None)
    got = ops.accept_views(pr, torch.from_numpy(labels), torch.from_numpy(scene_off[view_scene]),
                           max_points=4000)
    ref = []
    for v in range(len(vis)):
        lab = labels[scene_off[view_scene[v]]:scene_off[view_scene[v] + 1]][vis[v]]
        ref.append(bool(vis[v].sum() > 400 and (lab != 255).sum() > 10 and vis[v].sum() < 4000))
    assert got.tolist() == ref
    assert ref == [True, False, True, False, False, False]   # ok, too few, ok, none, all ignored, too many
