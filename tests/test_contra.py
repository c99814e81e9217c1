"""loss_contra's mask selection + pooling (models/utils/criterion.py:39-182) and the per-scene means
(models/xmask3d.py:239-258): oracle vs the golden produced by the reference's own `Criterion.loss_contra`
(CPU), and the device path vs both (`-m gpu`)."""
import types

import numpy as np
import pytest
import torch

H, W = 48, 64


def _split(g, tag):
    off = g[f"{tag}_off"]
    b = len(off) - 1
    xs = [torch.from_numpy(g[f"{tag}_x"][off[i]:off[i + 1]]) for i in range(b)]
    ys = [torch.from_numpy(g[f"{tag}_y"][off[i]:off[i + 1]]) for i in range(b)]
    gts = [torch.from_numpy(g[f"{tag}_gt"][off[i]:off[i + 1]]) for i in range(b)]
    f3 = [torch.from_numpy(g[f"{tag}_f3"][off[i]:off[i + 1]]) for i in range(b)]
    return xs, ys, gts, f3, torch.from_numpy(g[f"{tag}_pred_masks"]), torch.from_numpy(g[f"{tag}_clip"]), \
        torch.from_numpy(g[f"{tag}_me"]), g[f"{tag}_sel"], float(g[f"{tag}_loss"])


@pytest.mark.parametrize("tag", ["a", "b", "c"])
def test_oracle_loss_contra_golden(golden, tag):
    """The port reproduces what the reference's own method returned: the same pooled masks in the same order,
    the same float32 loss."""
    from oracle import ref_port
    xs, ys, gts, f3, pm, clip, me, sel, loss = _split(golden("contra"), tag)
    got_loss, fm, _, _ = ref_port.loss_contra_ref(xs, ys, pm, me, clip, f3, f3, gts, [H, W])
    assert float(got_loss) == loss
    up = torch.nn.functional.interpolate(pm, size=(H, W), mode="bilinear", align_corners=False)
    got_sel = -np.ones_like(sel)
    for s, planes in fm:
        for j, pl in enumerate(planes):
            got_sel[s, j] = [m for m in range(pm.shape[1]) if torch.equal(up[s, m], pl)][0]
    assert np.array_equal(got_sel, sel)


@pytest.fixture(scope="module")
def dev():
    assert torch.cuda.is_available(), "GPU tests need a CUDA device"
    return torch.device("cuda", 0)


@pytest.mark.gpu
@pytest.mark.parametrize("tag", ["a", "b", "c"])
def test_loss_contra_golden_gpu(golden, dev, tag):
    """Device drop-in vs the reference's own result: identical selection (mask planes, order), loss within float32
    rounding of the pooled means (1e-5)."""
    from xmask3d_b200.criterion import loss_contra
    xs, ys, gts, f3, pm, clip, me, sel, loss = _split(golden("contra"), tag)
    self = types.SimpleNamespace(cfg=types.SimpleNamespace(mask_shape=[H, W]))
    outputs = {"pred_masks": pm.to(dev), "mask_embed": me.to(dev), "mask_embed_clip": clip.to(dev),
               "fused_pred_feature": [f.to(dev) for f in f3], "pure3d_pred_feature": [f.to(dev) for f in f3]}
    got_loss, fm = loss_contra(self, xs, ys, None, None, gts, outputs)
    assert abs(float(got_loss["loss_3d_contra"]) - loss) < 1e-5 * max(1.0, abs(loss))
    up = torch.nn.functional.interpolate(pm.to(dev), size=(H, W), mode="bilinear", align_corners=False)
    want = [(s, sel[s][sel[s] >= 0]) for s in range(sel.shape[0]) if (sel[s] >= 0).any()]
    assert [s for s, _ in fm] == [s for s, _ in want]
    for (s, planes), (_, idx) in zip(fm, want):
        assert torch.equal(planes, up[s, torch.from_numpy(idx).to(dev)])


@pytest.mark.gpu
@pytest.mark.parametrize("k,n_lo,n_hi", [(50, 3000, 9000), (100, 5, 60), (33, 900, 1000)])
def test_contra_select_vs_oracle(dev, k, n_lo, n_hi):
    """Full-size masks (240 x 320, K = 50 / 100): counts after the guard, kinds, scores, the ranking and the pooled
    means of the selected masks against the restatement (oracle/ref_port.py: contra_select_ref, loss_contra_ref)."""
    from oracle import ref_port
    from xmask3d_b200 import ops
    from xmask3d_b200.criterion import contra_select
    g = torch.Generator().manual_seed(k)
    b, h, w, c = 5, 240, 320, 128
    pm = torch.randn(b, k, 30, 40, generator=g)
    pm = torch.nn.functional.avg_pool2d(pm, 5, stride=1, padding=2) * 6 - torch.rand(b, k, 1, 1, generator=g) * 2
    up = torch.nn.functional.interpolate(pm, size=(h, w), mode="bilinear", align_corners=False)
    up = torch.where(up.abs() < 1e-4, torch.full_like(up, 1e-3), up)          # keep clear of sigmoid(x) == 0.5 ties
    xs, ys, gts, f3 = [], [], [], []
    for s in range(b):
        n = int(torch.randint(n_lo, n_hi, (1,), generator=g))
        x, y = torch.randint(0, h, (n,), generator=g), torch.randint(0, w, (n,), generator=g)
        gt = (y >= int(torch.randint(w // 4, 3 * w // 4, (1,), generator=g))).float()
        gt[torch.rand(n, generator=g) < 0.1] = 2.0
        xs.append(x); ys.append(y); gts.append(gt); f3.append(torch.randn(n, c, generator=g))
    csel, member, seg_d, seg = contra_select(xs, ys, up.to(dev), gts)
    _, cnt5, mean5 = ops.pool(torch.cat(f3).to(dev), seg_d, 5, member=csel.sel_member, cap_pairs=5 * int(seg[-1]) + 2)
    counts, kind, score = csel.counts.cpu().numpy(), csel.kind.cpu().numpy(), csel.score.cpu().numpy()
    sel, n_sel = csel.sel.cpu().numpy(), csel.n_sel.cpu().numpy()
    n_selected = 0
    for s in range(b):
        r = ref_port.contra_select_ref(xs[s], ys[s], up[s], gts[s])
        mb = r["member"]
        assert np.array_equal(counts[s, :, 0], mb.sum(1).numpy())
        assert np.array_equal(counts[s, :, 1], (mb & (gts[s] == 0)).sum(1).numpy())
        assert np.array_equal(counts[s, :, 2], (mb & (gts[s] == 1)).sum(1).numpy())
        assert np.array_equal(kind[s], r["kind"].numpy())
        cand = r["kind"].numpy() > 0
        np.testing.assert_allclose(score[s][cand], r["score"].numpy()[cand], rtol=1e-5)
        assert np.all(np.isnan(score[s][~cand]))
        assert n_sel[s] == len(r["sel"]) and sel[s, :n_sel[s]].tolist() == r["sel"] and np.all(sel[s, n_sel[s]:] == -1)
        for j, m in enumerate(r["sel"]):
            ref = f3[s][mb[m]].double().mean(0)
            err = (mean5[s, j].cpu().double() - ref).abs().max() / ref.abs().max()
            assert float(err) < 1e-5 and int(cnt5[s, j]) == int(mb[m].sum())
            n_selected += 1
    if n_lo >= 900:
        assert n_selected >= 3                     # the case really selects something


@pytest.mark.gpu
def test_scene_mean_and_pool_mean_grad(dev):
    """models/xmask3d.py:239-258 (`feature.mean(0)` per scene) and the gradient of the masked mean pooling against
    torch autograd of the reference formula `feature[mask].mean(0)`."""
    from xmask3d_b200.autograd import pool_mean
    from xmask3d_b200.criterion import scene_mean
    g = torch.Generator().manual_seed(3)
    feats = [torch.randn(n, 256, generator=g).to(dev) for n in (700, 1, 3000)]
    got = scene_mean(feats)
    ref = torch.stack([f.double().mean(0) for f in feats])
    assert float(((got.double() - ref).abs().amax(1) / ref.abs().amax(1)).max()) < 1e-5
    # gradient: two segments, 5 overlapping masks
    n, k, c = [400, 250], 5, 64
    seg = torch.tensor([0, 400, 650], dtype=torch.int64, device=dev)
    mb = [torch.rand(k, ni, generator=g) < 0.3 for ni in n]
    mb[1][4] = False                                             # an empty mask: mean 0, no gradient
    words = torch.cat([(m.long() << torch.arange(k).view(k, 1)).sum(0) for m in mb]).to(torch.int32).view(-1, 1).to(dev)
    feat = torch.randn(650, c, generator=g).to(dev).requires_grad_(True)
    wgt = torch.randn(2, k, c, generator=g).to(dev)
    mean, cnt = pool_mean(feat, seg, words, k, cap_pairs=650 * k)
    (mean * wgt).sum().backward()
    f2 = feat.detach().clone().requires_grad_(True)
    tot = 0
    for s in range(2):
        fs = f2[int(seg[s]):int(seg[s + 1])]
        for m in range(k):
            if mb[s][m].any():
                tot = tot + (fs[mb[s][m].to(dev)].mean(0) * wgt[s, m]).sum()
    tot.backward()
    assert float((feat.grad - f2.grad).abs().max()) < 1e-5 * float(f2.grad.abs().max())
    got2 = scene_mean([f.clone().requires_grad_(True) for f in feats])
    assert float(((got2.detach().double() - ref).abs().amax(1) / ref.abs().amax(1)).max()) < 1e-5
