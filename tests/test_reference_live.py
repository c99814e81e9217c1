"""Live cross-check of the oracle against the UNMODIFIED reference functions on random inputs.

Runs only where the reference tree is mounted (the build container); it is skipped on the GPU box, where
the committed golden fixtures (produced by the same functions) take over.  Nothing here touches the GPU."""
import os
import sys

import numpy as np
import pytest
from hypothesis import given, settings, strategies as st

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "golden"))
import refimport  # noqa: E402

pytestmark = pytest.mark.skipif(not refimport.available(), reason="reference tree not mounted")

from oracle import ref_port  # noqa: E402


@pytest.fixture(scope="module")
def R():
    return refimport.load()


@settings(max_examples=20, deadline=None, derandomize=True)
@given(seed=st.integers(0, 2 ** 31 - 1), n=st.integers(1, 3000), with_depth=st.booleans())
def test_compute_mapping_live(R, seed, n, with_depth):
    """models/utils/fusion_util.py:46-142 through models/utils/mapping_util.py:getMapping."""
    rng = np.random.default_rng(seed)
    xyz = rng.uniform(-4, 4, (n, 3)).astype(np.float32)
    yaw = rng.uniform(0, 2 * np.pi)
    pose = np.eye(4)
    pose[:3, :3] = np.array([[np.cos(yaw), 0, np.sin(yaw)], [0, 1, 0], [-np.sin(yaw), 0, np.cos(yaw)]])
    pose[:3, 3] = rng.uniform(-1, 1, 3)
    depth = None
    if with_depth:
        depth = rng.integers(0, 6000, (240, 320)).astype(np.uint16) / 1000.0
    state = np.random.get_state()                          # getMapping reseeds numpy / torch as a side effect
    ref = R.getMapping().compute_mapping(pose, xyz, depth)
    np.random.set_state(state)
    got = ref_port.getMapping().compute_mapping(pose, xyz, depth)
    assert got.dtype == ref.dtype and np.array_equal(got, ref)


@settings(max_examples=20, deadline=None, derandomize=True)
@given(seed=st.integers(0, 2 ** 31 - 1), n=st.integers(1, 4000), voxel=st.sampled_from([0.02, 0.05, 0.25]))
def test_voxelize_and_quantize_live(R, seed, n, voxel):
    """dataset/voxelizer.py:81-132 (same np.random state) and dataset/voxelization_utils.py:6-102."""
    from tests.golden.make_golden_params import vox_kwargs
    rng = np.random.default_rng(seed)
    xyz = rng.uniform(-3, 3, (n, 3)).astype(np.float32)
    colors = rng.uniform(-1, 1, (n, 3)).astype(np.float32)
    labels = rng.integers(0, 20, n).astype(np.float64)
    outs = []
    for V in (R.Voxelizer, ref_port.Voxelizer):
        np.random.seed(seed % (2 ** 31))
        outs.append(V(**vox_kwargs(voxel)).voxelize(xyz, colors.copy(), labels.copy(), return_ind=True))
    for a, b in zip(*outs):
        assert np.asarray(a).dtype == np.asarray(b).dtype and np.array_equal(a, b)
    # float64 coordinates (what ElasticDistortion hands the voxelizer on the augmented path, augmentation.py:171)
    xyz64 = xyz.astype(np.float64) + rng.normal(0, 0.01, xyz.shape)
    outs = []
    for V in (R.Voxelizer, ref_port.Voxelizer):
        np.random.seed(seed % (2 ** 31))
        outs.append(V(**vox_kwargs(voxel)).voxelize(xyz64.copy(), colors.copy(), labels.copy(), return_ind=True))
    for a, b in zip(*outs):
        assert np.asarray(a).dtype == np.asarray(b).dtype and np.array_equal(a, b)
    grid = np.floor(rng.uniform(0, 30, (n, 3)))
    assert np.array_equal(ref_port.fnv_hash_vec(grid), R.fnv_hash_vec(grid))
    assert np.array_equal(ref_port.ravel_hash_vec(grid.copy() - 7.0), R.ravel_hash_vec(grid.copy() - 7.0))
    lab = rng.integers(0, 5, n)
    for kw in (dict(return_index=True), dict(labels=lab, return_index=True, set_ignore_label_when_collision=True),
               dict(return_index=True, hash_type="ravel"), dict(feats=colors), dict(quantization_size=0.5)):
        a = R.sparse_quantize(grid.copy(), **{k: (v.copy() if isinstance(v, np.ndarray) else v) for k, v in kw.items()})
        b = ref_port.sparse_quantize(grid.copy(), **{k: (v.copy() if isinstance(v, np.ndarray) else v) for k, v in kw.items()})
        a, b = (a if isinstance(a, tuple) else (a,)), (b if isinstance(b, tuple) else (b,))
        assert len(a) == len(b) and all(np.array_equal(x, y) for x, y in zip(a, b)), kw


@settings(max_examples=15, deadline=None, derandomize=True)
@given(seed=st.integers(0, 2 ** 31 - 1), b=st.integers(1, 3), k=st.integers(1, 30), t=st.integers(2, 40))
def test_cal_pred_logits_live(R, seed, b, k, t):
    """models/xmask3d.py:129-143 — the reference's own method (self is unused)."""
    import torch
    g = torch.Generator().manual_seed(seed)
    outputs = {"mask_embed": torch.randn(b, k, 48, generator=g), "text_embed": torch.randn(t - 1, 48, generator=g),
               "null_embed": torch.randn(1, 48, generator=g), "labels": [[str(i)] for i in range(t - 1)],
               "logit_scale": torch.tensor(1 / 0.07)}
    ref = R.XMASK3d.cal_pred_logits(None, dict(outputs))
    got = ref_port.cal_pred_logits(dict(outputs))
    assert torch.equal(got, ref)


@settings(max_examples=15, deadline=None, derandomize=True)
@given(seed=st.integers(0, 2 ** 31 - 1), n=st.integers(1, 400), k=st.integers(1, 12), density=st.sampled_from([0.0, 0.1, 0.6]))
def test_mask_mapper_live(R, seed, n, k, density):
    """models/utils/fuser.py:6-53 run on CPU tensors with identity fc layers and an additive fuser: its
    mask -> point scatter-mean (output_2d, since fc2 is the identity) equals the port's bit for bit, and the
    fused output follows from it."""
    import types

    import torch
    g = torch.Generator().manual_seed(seed)
    h, w, c = 24, 32, 8
    masks = (torch.rand(k, h, w, generator=g) < density).float()
    x = torch.randint(0, h, (n,), generator=g)
    y = torch.randint(0, w, (n,), generator=g)
    emb = torch.randn(k, c, generator=g)
    pred = torch.randn(n, c, generator=g)
    ident = torch.nn.Identity()
    fuse = lambda a, b: a + b                                # noqa: E731
    cfg = types.SimpleNamespace(caption_contra_2d_pre=True)
    out, out2d, out3d, pre = R.mask_mapper([x], [y], [masks], [emb], [pred], fuse, ident, ident, cfg)
    member = ref_port.gather_masks(masks, x, y, "ge0.5")
    feat2d, counter = ref_port.scatter_mask_embed(member, emb, torch.zeros(n, c))
    assert torch.equal(out2d[0], feat2d) and torch.equal(out3d[0], pred)
    covered = counter[:, 0] >= 1
    want = pred.clone()
    want[covered] = feat2d[covered] + pred[covered]
    assert torch.equal(out[0], want) and torch.equal(pre[0], feat2d[covered])


@settings(max_examples=10, deadline=None, derandomize=True)
@given(seed=st.integers(0, 2 ** 31 - 1), ks=st.lists(st.integers(1, 9), min_size=2, max_size=4))
def test_mask_mapper_ragged_lists_live(R, seed, ks):
    """The caller's real input (criterion.py:262-340): per-scene lists with DIFFERENT numbers of masks, bool
    partition masks for some scenes and float32 zero masks for others — reference vs the port's restatement."""
    import types

    import torch
    g = torch.Generator().manual_seed(seed)
    h, w, c = 24, 32, 8
    xs, ys, masks, embs, preds = [], [], [], [], []
    for i, k in enumerate(ks):
        n = int(torch.randint(1, 300, (1,), generator=g))
        xs.append(torch.randint(0, h, (n,), generator=g))
        ys.append(torch.randint(0, w, (n,), generator=g))
        if i % 3 == 2:
            masks.append(torch.zeros(k, h, w))                                   # "nothing kept": float zeros
        else:
            lab = torch.randint(0, k + 1, (h, w), generator=g)                   # partition, label k = background
            masks.append(lab.unsqueeze(0) == torch.arange(k).view(k, 1, 1))      # bool
        embs.append(torch.randn(k, c, generator=g))
        preds.append(torch.randn(n, c, generator=g))
    ident = torch.nn.Identity()
    fuse = lambda a, b: a * 0.5 + b                          # noqa: E731
    cfg = types.SimpleNamespace(caption_contra_2d_pre=True)
    ref = R.mask_mapper(xs, ys, masks, embs, preds, fuse, ident, ident, cfg)
    got = ref_port.mask_mapper_ref(xs, ys, masks, embs, preds, fuse, ident, ident, True)
    for a, b in zip(ref, got):
        assert len(a) == len(b) and all(torch.equal(u, v) for u, v in zip(a, b))


def _contra_inputs(seed, b, k, h, w, c, n_lo=60, n_hi=900, frac_ignored=0.1):
    """Synthetic loss_contra inputs: low-resolution mask logits whose up-sampled masks cover 5-60 % of the image,
    binary_gt spatially coherent (left part novel, right part base) so that both candidate classes occur."""
    import torch
    g = torch.Generator().manual_seed(seed)
    pm = torch.randn(b, k, h // 2, w // 2, generator=g)
    pm = torch.nn.functional.avg_pool2d(pm, 5, stride=1, padding=2) * 6 - torch.rand(b, k, 1, 1, generator=g) * 2
    xs, ys, gts, ff, f3 = [], [], [], [], []
    for _ in range(b):
        n = int(torch.randint(n_lo, n_hi, (1,), generator=g))
        x, y = torch.randint(0, h, (n,), generator=g), torch.randint(0, w, (n,), generator=g)
        split = int(torch.randint(w // 4, 3 * w // 4, (1,), generator=g))
        gt = (y >= split).float()
        gt[torch.rand(n, generator=g) < frac_ignored] = 2.0
        xs.append(x); ys.append(y); gts.append(gt)
        ff.append(torch.randn(n, c, generator=g)); f3.append(torch.randn(n, c, generator=g))
    return pm, xs, ys, gts, ff, f3, torch.randn(b, k, c, generator=g), torch.randn(b, k, c, generator=g)


@settings(max_examples=12, deadline=None, derandomize=True)
@given(seed=st.integers(0, 2 ** 31 - 1), b=st.integers(1, 3), k=st.integers(1, 14), sparse=st.booleans())
def test_loss_contra_live(R, seed, b, k, sparse):
    """models/utils/criterion.py:39-182 — the reference's own `Criterion.loss_contra` (self only supplies
    cfg.mask_shape) against the port: same loss, same selected mask planes in the same order."""
    import types

    import torch
    h, w, c = 24, 32, 8
    pm, xs, ys, gts, ff, f3, me, ce = _contra_inputs(seed, b, k, h, w, c, n_lo=(3 if sparse else 200), n_hi=(40 if sparse else 900))
    self = types.SimpleNamespace(cfg=types.SimpleNamespace(mask_shape=[h, w]))
    outputs = {"pred_masks": pm, "mask_embed": me, "mask_embed_clip": ce, "fused_pred_feature": ff, "pure3d_pred_feature": f3}
    ref_loss, ref_masks = R.Criterion.loss_contra(self, xs, ys, None, None, gts, outputs)
    loss, fm, _, _ = ref_port.loss_contra_ref(xs, ys, pm, me, ce, ff, f3, gts, [h, w])
    assert torch.equal(ref_loss["loss_3d_contra"], loss)
    assert len(ref_masks) == len(fm)
    for (rb, rm), (pb, pmk) in zip(ref_masks, fm):
        assert rb == pb and torch.equal(rm, pmk)
