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


def test_ragged_mask_lists_pad_to_never_hitting_planes():
    """mask_mapper's caller passes per-scene lists with different numbers of masks and mixed bool / float dtypes
    (models/utils/criterion.py:262-340); the padding must never hit under the call site's threshold."""
    import torch
    from xmask3d_b200.fuser import _stack_ragged_embeds, _stack_ragged_masks
    g = torch.Generator().manual_seed(0)
    a = torch.rand(3, 4, 5, generator=g) > 0.5                   # bool, K' = 3
    b = torch.zeros(5, 4, 5)                                     # float zeros, K = 5
    c = torch.rand(2, 4, 5, generator=g)                         # float probabilities, K' = 2
    m, ks = _stack_ragged_masks([a, b, c], "ge0.5")
    assert ks == [3, 5, 2] and m.shape == (3, 5, 4, 5) and m.dtype == torch.uint8
    assert torch.equal(m[0, :3].bool(), a) and not m[0, 3:].any()
    assert not m[1].any()
    assert torch.equal(m[2, :2].bool(), c >= 0.5) and not m[2, 2:].any()
    lg = [torch.randn(2, 4, 5, generator=g), torch.randn(4, 4, 5, generator=g)]
    m, ks = _stack_ragged_masks(lg, "sigmoid_ge0.5")
    assert m.dtype == torch.float32 and torch.equal(m[0, :2], lg[0]) and bool((m[0, 2:].sigmoid() < 0.5).all())
    same = [torch.rand(3, 4, 5, generator=g) for _ in range(2)]
    m, ks = _stack_ragged_masks(same, "ge0.5")
    assert torch.equal(m, torch.stack(same)) and ks == [3, 3]
    e = _stack_ragged_embeds([torch.ones(3, 6), torch.ones(5, 6), torch.ones(2, 6)], 5)
    assert e.shape == (3, 5, 6) and float(e.sum()) == (3 + 5 + 2) * 6


def test_bench_scene_sharding_is_a_balanced_partition():
    """bench.py's rank_scene_ids: the 8 x N distinct scenes of a job are assigned to ranks exactly once, every rank
    gets work, and the visible-pair load (xmask3d_b200/scene_costs.json) is balanced to a few per cent; the 312-scene
    sweep (configs[4]) likewise over 1 / 2 / 4 / 8 ranks."""
    import argparse

    import numpy as np

    import bench
    from xmask3d_b200 import dist as xd
    for world in (1, 2, 4, 8):
        args = argparse.Namespace(scenes=8, views=20, points=150_000, replicas=False)
        shards = [bench.rank_scene_ids(args, r, world) for r in range(world)]
        assert sorted(sum(shards, [])) == list(range(8 * world))
        loads = [sum(bench.scene_cost(1000 + g, 150_000, 20) for g in s) for s in shards]
        assert max(loads) / np.mean(loads) < 1.03
        sizes = bench.sweep_sizes(312)
        assert sizes.min() >= 30_000 and sizes.max() <= 500_000
        costs = [bench.scene_cost(1000 + s, int(n), 20) for s, n in enumerate(sizes)]
        sh = xd.shard_scenes(costs, world)
        assert sorted(sum(sh, [])) == list(range(312))
        ld = [sum(costs[i] for i in s) for s in sh]
        assert max(ld) / np.mean(ld) < 1.01
    assert bench.rank_scene_ids(argparse.Namespace(scenes=8, views=20, points=150_000, replicas=True), 3, 8) == list(range(8))
