"""The oracle (numpy port + C restatement) replayed against fixtures produced by the
reference's own functions (tests/golden/make_golden.py).  CPU only."""
import numpy as np
import pytest
import torch

from oracle import ref_port as P

def _intr(g):
    k = g["intrinsics"]
    return (k[0, 0], k[1, 1], k[0, 2], k[1, 2])


def test_fnv_known_answers(golden, cport):
    g = golden("hash")
    # hand-checkable values recorded in SURVEY.md §8a-V3
    assert int(g["kat_fnv"][0]) == 15658191375538532279
    assert int(g["kat_fnv"][1]) == 15657232601398921515
    for arr, ref in ((g["kat"], g["kat_fnv"]), (g["rnd"], g["rnd_fnv"])):
        assert np.array_equal(P.fnv_hash_vec(arr), ref)
        assert np.array_equal(cport.fnv(arr), ref)
    assert np.array_equal(P.ravel_hash_vec(g["rnd"] - 100.0), g["rnd_ravel"])


def test_sparse_quantize_paths(golden, cport):
    g = golden("hash")
    inds, inv = P.sparse_quantize(g["ex"], return_index=True)
    assert inds.tolist() == [3, 0, 1] and inv.tolist() == [1, 2, 1, 0, 2]
    assert np.array_equal(inds, g["ex_inds"]) and np.array_equal(inv, g["ex_inv"])
    li, ll = P.sparse_quantize(g["dup"], labels=g["dup_labels"].copy(), return_index=True,
                               set_ignore_label_when_collision=True)
    assert np.array_equal(li, g["dup_lab_inds"]) and np.array_equal(ll, g["dup_lab_out"])
    ri, rinv = P.sparse_quantize(g["dup"] - 5.0, return_index=True, hash_type="ravel")
    assert np.array_equal(ri, g["dup_ravel_inds"]) and np.array_equal(rinv, g["dup_ravel_inv"])
    first, inverse, counts = cport.unique_u64(P.fnv_hash_vec(g["dup"]))
    assert np.array_equal(first, li)
    expect = g["dup_labels"][first].copy()
    expect[counts > 1] = 255
    assert np.array_equal(expect, ll)
    q = golden("quant")
    qi, qinv = P.sparse_quantize(q["coords"], return_index=True, quantization_size=0.05)
    assert np.array_equal(qi, q["inds"]) and np.array_equal(qinv, q["inv"])
    with pytest.raises(AssertionError):
        P.sparse_quantize(g["ex"], hash_type="md5")
    with pytest.raises(AssertionError):
        P.sparse_quantize(g["ex"][:, 0])


def test_project_matches_reference(golden, cport):
    g = golden("project")
    m = P.getMapping()
    assert np.array_equal(m.intrinsics, g["intrinsics"])
    for v in range(g["poses"].shape[0]):
        pose, dmm = g["poses"][v], g["depth_mm"][v]
        assert np.array_equal(m.compute_mapping(pose, g["xyz"], dmm / 1000), g["mapping"][v])
        assert np.array_equal(m.compute_mapping(pose, g["xyz"], None), g["mapping_nodepth"][v])
        w2c = np.linalg.inv(pose)
        assert np.array_equal(cport.project(g["xyz"], w2c, _intr(g), dmm), g["mapping"][v])
        assert np.array_equal(cport.project(g["xyz"], w2c, _intr(g), dmm / 1000), g["mapping"][v])
        assert np.array_equal(cport.project(g["xyz"], w2c, _intr(g), None), g["mapping_nodepth"][v])
    # depth image smaller than the camera image, cut_bound = 0
    m0 = P.PointCloudToImageMapper((320, 240), 0.25, 0, g["intrinsics"])
    sd = g["small_depth_mm"]
    assert np.array_equal(m0.compute_mapping(g["poses"][0], g["xyz"], sd / 1000), g["mapping_small_cut0"])
    got = cport.project(g["xyz"], np.linalg.inv(g["poses"][0]), _intr(g), sd, cut=0)
    assert np.array_equal(got, g["mapping_small_cut0"])
    vis, xl, yl = P.compact_mapping(g["mapping"][0])
    assert vis.sum() == len(xl) == len(yl) > 400


def test_voxelize_matches_reference(golden, cport):
    g = golden("voxelize")
    from tests.golden.make_golden_params import vox_kwargs
    for tag in "abc":
        np.random.seed(int(g[f"seed_{tag}"]))
        vox = P.Voxelizer(**vox_kwargs(float(g[f"vs_{tag}"])))
        grid, feats, labels, inv, inds = vox.voxelize(g["xyz"], g["colors"].copy(), g["labels"].copy(),
                                                      return_ind=True)
        assert np.array_equal(grid, g[f"grid_{tag}"]) and grid.dtype == np.float64
        assert np.array_equal(inv, g[f"inv_{tag}"]) and np.array_equal(inds, g[f"inds_{tag}"])
        assert np.array_equal(feats, g[f"feats_{tag}"]) and np.array_equal(labels, g[f"labels_{tag}"])
        cgrid, cfirst, cinv = cport.voxelize(g["xyz"], g[f"rt_{tag}"])
        assert np.array_equal(cgrid, g[f"grid_{tag}"])
        assert np.array_equal(cfirst, g[f"inds_{tag}"]) and np.array_equal(cinv, g[f"inv_{tag}"])
    grid, _, _, inv, inds = P.Voxelizer(voxel_size=0.04).voxelize(g["xyz"], g["colors"].copy(),
                                                               g["labels"].copy(), return_ind=True)
    assert np.array_equal(grid, g["grid_plain"]) and np.array_equal(inv, g["inv_plain"])
    rt = np.eye(4)
    np.fill_diagonal(rt[:3, :3], 1 / 0.04)
    cgrid, cfirst, cinv = cport.voxelize(g["xyz"], rt)
    assert np.array_equal(cgrid, g["grid_plain"]) and np.array_equal(cfirst, g["inds_plain"])


def test_scatter_and_pool(golden, cport):
    g = golden("pool")
    xl, yl = torch.from_numpy(g["x_label"]), torch.from_numpy(g["y_label"])
    emb, pred3d = torch.from_numpy(g["emb"]), torch.from_numpy(g["pred3d"])
    for tag, masks in (("part", torch.from_numpy(g["part"]).float()), ("over", torch.from_numpy(g["over"]))):
        member = P.gather_masks(masks, xl, yl, "ge0.5")
        feat2d, counter = P.scatter_mask_embed(member, emb, pred3d)
        assert np.array_equal(feat2d.numpy(), g[f"feat2d_{tag}"])                  # bit-exact fp32
        covered = (counter.sum(1) >= 1)
        fused = pred3d.clone()
        fused[covered] = feat2d[covered] + 2.0 * pred3d[covered]
        assert np.array_equal(fused.numpy(), g[f"fused_{tag}"])
        assert np.array_equal(feat2d[covered].numpy(), g[f"pre_{tag}"])
        cout, ccnt = cport.scatter_member_f32(member.numpy(), g["emb"])
        assert np.array_equal(cout, g[f"feat2d_{tag}"])
        for mode, key in (("ge0.5", "ge"), ("sigmoid_ge0.5", "sig")):
            mem = P.gather_masks(masks, xl, yl, mode)
            mean, cnt = P.masked_mean_pool(pred3d, mem)
            assert np.array_equal(cnt.numpy(), g[f"cnt_{tag}_{key}"])
            assert np.array_equal(mean.numpy(), g[f"mean_{tag}_{key}"])
            s64, c64 = cport.pool_member_f64(g["pred3d"], mem.numpy())
            assert np.array_equal(c64, g[f"cnt_{tag}_{key}"])
            ref = g[f"mean_{tag}_{key}"]
            got = s64 / np.maximum(c64, 1)[:, None]
            err = np.abs(got - ref).max(1) / np.maximum(np.abs(ref).max(1), 1e-30)
            assert err.max() < 1e-5
    mem = P.gather_masks(torch.from_numpy(g["over"]), xl, yl, "sigmoid_gt0.5")
    sp = P.masked_score_pool(torch.from_numpy(g["score"]), mem)
    assert np.array_equal(sp.numpy(), g["score_pool"])
    assert np.array_equal((mem.sum(1) > 0).numpy(), g["score_keep"])
    # partition form through the label API
    label = torch.from_numpy(g["part"]).float()[:, xl, yl].argmax(0).numpy().astype(np.int32)
    s64, c64 = cport.pool_label_f64(g["pred3d"], label, g["part"].shape[0])
    assert np.array_equal(c64, g["cnt_part_ge"])


def test_logits(golden):
    g = golden("logits")
    for tag, t in (("b15", 20), ("sn200", 201)):
        o = {"mask_embed": torch.from_numpy(g[f"me_{tag}"]), "text_embed": torch.from_numpy(g[f"te_{tag}"]),
             "null_embed": torch.from_numpy(g[f"ne_{tag}"]), "labels": [[str(i)] for i in range(t - 1)],
             "logit_scale": torch.tensor(1 / 0.07)}
        assert np.array_equal(P.cal_pred_logits(o).numpy(), g[f"logits_{tag}"])
    sizes = g["sizes_grp"].tolist()
    labels, s = [], 0
    for n in sizes:
        labels.append([str(s + j) for j in range(n)])
        s += n
    o = {"mask_embed": torch.from_numpy(g["me_grp"]), "text_embed": torch.from_numpy(g["te_grp"]),
         "null_embed": torch.from_numpy(g["ne_grp"]), "labels": labels, "logit_scale": torch.tensor(100.0)}
    assert np.array_equal(P.cal_pred_logits(o).numpy(), g["logits_grp"])
    raw = torch.from_numpy(g["ens_in"])
    assert np.array_equal(P.ensemble_logits_with_labels(raw, labels, "max").numpy(), g["ens_max"])
    assert np.array_equal(P.ensemble_logits_with_labels(raw, labels, "mean").numpy(), g["ens_mean"])
    with pytest.raises(AssertionError):
        P.ensemble_logits_with_labels(raw, labels[:-1], "max")


def test_nn_fill_and_segment_max_oracle_vs_reference_statements(golden):
    """The restated exhaustive search returns what the reference's KDTree(k=1) statements returned
    (tests/golden/make_golden_nnfill.py), wherever the nearest neighbour is unique; segment maxima equal."""
    from oracle import ref_port
    g = golden("nnfill")
    match = ref_port.nn_fill_match(g["xyz"], g["counter"])
    false_idx = g["false_idx"]
    assert np.array_equal(false_idx, np.nonzero(g["counter"] == 0)[0])
    same = match[false_idx] == g["match"]
    if not same.all():       # only exact distance ties may differ
        xyz = g["xyz"].astype(np.float64)
        bad = false_idx[~same]
        d_mine = ((xyz[bad] - xyz[match[bad]]) ** 2).sum(1)
        d_ref = ((xyz[bad] - xyz[g["match"][~same]]) ** 2).sum(1)
        assert np.array_equal(d_mine, d_ref)
    assert np.array_equal(g["pred"][match], g["filled"])
    seen = np.nonzero(g["counter"] != 0)[0]
    assert np.array_equal(match[seen], seen)
    assert np.array_equal(ref_port.segment_max(g["feat"], g["idx"]), g["segmax"])


def test_voxelizer_clip_matches_reference(golden):
    """Voxelizer.clip (dataset/voxelizer.py:60-79; dead at every shipped config but part of the API):
    the host shim returns the reference's crop predicate, with and without an explicit centre."""
    from tests.golden.make_golden_params import vox_kwargs
    from xmask3d_b200.voxelizer import Voxelizer
    g = golden("clip")
    kw = vox_kwargs(0.05)
    kw["clip_bound"] = tuple(map(tuple, g["clip"]))
    vox = Voxelizer(**kw)
    assert np.array_equal(vox.clip(g["xyz"]), g["inside_default"])
    got = vox.clip(g["xyz"], center=g["centre"].copy(), trans_aug_ratio=np.array([0.1, -0.05, 0.0]))
    assert np.array_equal(got, g["inside_centre"])
