"""Fused mask preparation against the torch-CPU restatement of the reference's
inline sequence (oracle/ref_port.py:mask_prep_ref; models/xmask3d.py:326-331, 356-358, 391-435).

Bars: the bilinear upsample is bit-exact with torch's CPU F.interpolate; thresholded memberships and
partition labels are identical wherever the decision is not within float32 rounding of a tie
(|logit| < 1e-5, or top-2 gap of score * sigmoid below 1e-6 — sigmoid goes through a different exp
implementation on the GPU); areas agree up to the number of such ambiguous pixels.
"""
import numpy as np
import pytest
import torch

from oracle import ref_port


def _smooth_logits(g, k, hs, ws):
    z = torch.randn(k, 1, hs // 8, ws // 8, generator=g)
    z = torch.nn.functional.interpolate(z, size=(hs, ws), mode="bicubic", align_corners=False)[:, 0]
    return (z * 3.0 + torch.randn(k, 1, 1, generator=g) * 0.5).contiguous()


def test_oracle_bilinear_formula_matches_torch():
    """The op sequence the CUDA kernel uses (fma source index, fma(w0, a, w1*b) lerps) restated in
    numpy reproduces torch's CPU bilinear kernel bit for bit — pins the arithmetic without a GPU."""
    g = torch.Generator().manual_seed(5)
    x = torch.randn(3, 128, 128, generator=g)
    ref = torch.nn.functional.interpolate(x[None], size=(240, 320), mode="bilinear", align_corners=False)[0].numpy()
    f = np.float32
    xs = x.numpy()

    def fma(a, b, c):
        return (a.astype(np.float64) * b.astype(np.float64) + np.asarray(c, np.float64)).astype(f)

    def src(n_in, n_out):
        d = np.arange(n_out, dtype=f) + f(0.5)
        s = np.maximum(fma(np.full_like(d, f(n_in) / f(n_out)), d, -0.5), f(0))
        i0 = np.minimum(s.astype(np.int64), n_in - 1)
        i1 = i0 + (i0 < n_in - 1)
        l1 = np.clip((s - i0.astype(f)).astype(f), f(0), f(1))
        return i0, i1, (f(1) - l1).astype(f), l1
    y0, y1, ly0, ly1 = src(128, 240)
    x0, x1, lx0, lx1 = src(128, 320)
    a, b = xs[:, y0][:, :, x0], xs[:, y0][:, :, x1]
    c, d = xs[:, y1][:, :, x0], xs[:, y1][:, :, x1]
    LX0, LX1 = np.broadcast_to(lx0[None, None, :], a.shape), np.broadcast_to(lx1[None, None, :], a.shape)
    LY0, LY1 = np.broadcast_to(ly0[None, :, None], a.shape), np.broadcast_to(ly1[None, :, None], a.shape)
    r0 = fma(LX0, a, (LX1 * b).astype(f))
    r1 = fma(LX0, c, (LX1 * d).astype(f))
    out = fma(LY0, r0, (LY1 * r1).astype(f))
    assert np.array_equal(out, ref)


@pytest.mark.gpu
@pytest.mark.parametrize("k,hs,ws,h,w", [(50, 128, 128, 240, 320), (7, 64, 48, 100, 130), (100, 128, 128, 240, 320),
                                          (33, 32, 32, 32, 32)])
def test_mask_prep_vs_torch(k, hs, ws, h, w):
    from xmask3d_b200 import ops
    dev = torch.device("cuda", 0)
    g = torch.Generator().manual_seed(100 + k)
    n_seg = 3
    lg = torch.stack([_smooth_logits(g, k, hs, ws) for _ in range(n_seg)])
    scores = torch.rand(n_seg, k, generator=g)
    scores[:, ::7] = 0.0                                       # dropped by keep = scores > 0
    keep = scores > 0
    pm = ops.mask_prep(lg.to(dev), (h, w), scores=scores.to(dev), keep=keep.to(dev), mode="sigmoid_gt0.5",
                       want_bits=True, want_partition=True, want_upsampled=True)
    torch.cuda.synchronize()
    up = pm.upsampled.cpu()
    label = pm.label.cpu().to(torch.int64)
    areas = pm.areas.cpu().to(torch.int64)
    bits = pm.pixbits.cpu().numpy().view(np.uint32).reshape(n_seg, -1, h, w)
    for s in range(n_seg):
        ref = ref_port.mask_prep_ref(lg[s], scores[s], (h, w), 0.0)
        assert torch.equal(up[s], ref["up"]), "bilinear upsample must be bit-exact with torch CPU"
        # (a) thresholded membership: sigmoid(x) > 0.5
        got = np.stack([(bits[s, m >> 5] >> (m & 31)) & 1 for m in range(k)]).astype(bool)
        want = (ref["sig"] > 0.5).numpy()
        sure = (ref["up"].abs() > 1e-5).numpy()
        assert np.array_equal(got[sure], want[sure])
        # (b) partition labels / areas
        prob = ref["prob"]
        top2 = prob.topk(2, dim=0).values
        amb = ((top2[0] - top2[1]) < 1e-6) | (ref["up"].abs().min(0).values < 1e-5)
        in_final = torch.zeros(k, dtype=torch.bool)
        in_final[ref["final_keep"]] = True
        want_label = torch.full((h, w), -1, dtype=torch.int64)
        for j, m in enumerate(ref["final_keep"]):
            want_label[ref["final_mask"][j]] = m
        assert torch.equal(label[s][~amb], want_label[~amb])
        n_amb = int(amb.sum())
        assert int((areas[s] - ref["areas"]).abs().max()) <= n_amb
        if n_amb == 0:
            assert torch.equal(areas[s][:, 2] > 0, in_final)


@pytest.mark.gpu
def test_mask_prep_points_and_pool():
    """The label image drives pooling exactly like the reference's stacked bool masks gathered at the
    points: `mask[:, x_label, y_label] >= 0.5` then feature[mask].mean(0)."""
    from xmask3d_b200 import ops
    from xmask3d_b200.mask_prep import partition_labels, partition_masks, masks_at_points_lowres
    dev = torch.device("cuda", 0)
    g = torch.Generator().manual_seed(9)
    n_seg, k, hs, ws, h, w, c = 2, 20, 128, 128, 240, 320, 64
    lg = torch.stack([_smooth_logits(g, k, hs, ws) for _ in range(n_seg)])
    scores = torch.rand(n_seg, k, generator=g) + 0.01
    n_pts = [5000, 3777]
    rowcol = torch.cat([torch.stack([torch.randint(10, 230, (n,), generator=g), torch.randint(10, 310, (n,), generator=g)], 1)
                        for n in n_pts]).to(torch.int32)
    seg_off = torch.tensor([0, n_pts[0], sum(n_pts)], dtype=torch.int64)
    feat = torch.randn(sum(n_pts), c, generator=g)
    label, final_keep, areas = partition_labels(lg.to(dev), scores.to(dev), (h, w))
    pl = ops.gather_labels(label, rowcol.to(dev), seg_off.to(dev))
    s_, cnt, mean = ops.pool(feat.to(dev), seg_off.to(dev), k, label=pl)
    torch.cuda.synchronize()
    pl_c, mean_c, cnt_c = pl.cpu().to(torch.int64), mean.cpu(), cnt.cpu()
    for s in range(n_seg):
        ref = ref_port.mask_prep_ref(lg[s], scores[s], (h, w), 0.0)
        rc = rowcol[seg_off[s]:seg_off[s + 1]].to(torch.int64)
        f = feat[seg_off[s]:seg_off[s + 1]]
        top2 = ref["prob"].topk(2, dim=0).values
        amb = ((top2[0] - top2[1]) < 1e-6) | (ref["up"].abs().min(0).values < 1e-5)
        assert int(amb[rc[:, 0], rc[:, 1]].sum()) == 0, "test points must avoid ambiguous pixels"
        for j, m in enumerate(ref["final_keep"]):
            m3d = ref["final_mask"][j][rc[:, 0], rc[:, 1]]
            assert torch.equal(pl_c[seg_off[s]:seg_off[s + 1]] == m, m3d)
            if int(m3d.sum()):
                want = f[m3d].mean(0)
                assert int(cnt_c[s, m]) == int(m3d.sum())
                assert float((mean_c[s, m] - want).abs().max()) <= 1e-5 * float(want.abs().max()) + 1e-6
        # reference-shaped single-view call
        idx, masks = partition_masks(lg[s].to(dev), scores[s].to(dev), (h, w))
        assert idx.cpu().tolist() == ref["final_keep"]
        assert torch.equal(masks.cpu()[:, ~amb], torch.stack(ref["final_mask"])[:, ~amb])
    # overlapping masks at the points from the low-resolution logits
    member = masks_at_points_lowres(lg.to(dev), (h, w), rowcol.to(dev), seg_off.to(dev))
    mb = member.cpu().numpy().view(np.uint32)
    for s in range(n_seg):
        up = torch.nn.functional.interpolate(lg[s][None], size=(h, w), mode="bilinear", align_corners=False)[0]
        rc = rowcol[seg_off[s]:seg_off[s + 1]].to(torch.int64)
        vals = up[:, rc[:, 0], rc[:, 1]]
        want = (vals.sigmoid() > 0.5).numpy()
        got = np.stack([(mb[seg_off[s]:seg_off[s + 1], m >> 5] >> (m & 31)) & 1 for m in range(k)]).astype(bool)
        sure = (vals.abs() > 1e-5).numpy()
        assert np.array_equal(got[sure], want[sure])
