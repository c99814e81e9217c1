"""Generate tests/golden/*.npz by EXECUTING THE REFERENCE'S OWN FUNCTIONS (imported
unmodified from /root/reference through refimport.py).  Build-container only; the fixtures
are committed so that the GPU box (which has no /root/reference) can replay them.

    python tests/golden/make_golden.py

Every array an oracle/CUDA test needs (inputs and reference outputs) is stored, so replay
needs no RNG agreement.  Sizes are kept small (< 3 MB total).
"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))

import refimport  # noqa: E402
from xmask3d_b200 import synthetic as syn  # noqa: E402

R = refimport.load()
L = R.Point3DLoader


def save(name, **kw):
    path = os.path.join(HERE, name + ".npz")
    np.savez_compressed(path, **kw)
    print(f"{name}: {os.path.getsize(path) / 1024:.0f} KiB")


def vox_kwargs(voxel_size=0.02):
    return dict(voxel_size=voxel_size, clip_bound=None, use_augmentation=True,
                scale_augmentation_bound=L.SCALE_AUGMENTATION_BOUND,
                rotation_augmentation_bound=L.ROTATION_AUGMENTATION_BOUND,
                translation_augmentation_ratio_bound=L.TRANSLATION_AUGMENTATION_RATIO_BOUND)


def golden_hash():
    kat = np.array([[0, 0, 0], [1, 2, 3], [3, 1, 2], [511, 0, 77], [1023, 1023, 255]], dtype=np.float64)
    rng = np.random.default_rng(3)
    rnd = np.floor(rng.uniform(0, 600, (4000, 3)))
    ex = np.array([[3, 1, 2], [0, 0, 0], [3, 1, 2], [5, 5, 5], [0, 0, 0]], dtype=np.float64)
    inds, inv = R.sparse_quantize(ex, return_index=True)
    labels = rng.integers(0, 20, 4000)
    dup = np.floor(rng.uniform(0, 12, (4000, 3)))            # heavy collisions
    li, ll = R.sparse_quantize(dup, labels=labels.copy(), return_index=True,
                               set_ignore_label_when_collision=True)
    ri, rinv = R.sparse_quantize(dup - 5.0, return_index=True, hash_type="ravel")
    save("hash",
         kat=kat, kat_fnv=R.fnv_hash_vec(kat), rnd=rnd, rnd_fnv=R.fnv_hash_vec(rnd),
         rnd_ravel=R.ravel_hash_vec(rnd - 100.0),
         ex=ex, ex_inds=inds, ex_inv=inv,
         dup=dup, dup_labels=labels, dup_lab_inds=li, dup_lab_out=ll,
         dup_ravel_inds=ri, dup_ravel_inv=rinv,
         )
    # quantization_size case stored separately with its exact inputs
    rng2 = np.random.default_rng(11)
    qc = rng2.uniform(-3, 3, (3000, 3))
    qi, qinv = R.sparse_quantize(qc, return_index=True, quantization_size=0.05)
    save("quant", coords=qc, inds=qi, inv=qinv)


def golden_project_voxelize():
    sc = syn.make_scene(424242, 24_000)
    m = R.getMapping()
    poses, depths, maps, maps_nodepth = [], [], [], []
    for v in range(3):
        vw = syn.make_view(sc, v)
        poses.append(vw.pose)
        depths.append(vw.depth_mm)
        maps.append(m.compute_mapping(vw.pose, sc.xyz, vw.depth_m))
        maps_nodepth.append(m.compute_mapping(vw.pose, sc.xyz, None))
    # cut_bound=0 mapper with a depth image SMALLER than the image (exercises the
    # in-depth-bounds branch, fusion_util.py:105-135) and one that no inside point hits.
    m0 = R.PointCloudToImageMapper((320, 240), 0.25, 0, m.intrinsics)
    small_depth = (depths[0][:200, :260]).copy()
    map_small = m0.compute_mapping(poses[0], sc.xyz, small_depth / 1000)
    save("project", xyz=sc.xyz, poses=np.stack(poses), depth_mm=np.stack(depths),
         mapping=np.stack(maps), mapping_nodepth=np.stack(maps_nodepth),
         intrinsics=m.intrinsics, small_depth_mm=small_depth, mapping_small_cut0=map_small)

    vis = maps[0][:, 2] == 1
    pts, col, lab = sc.xyz[vis], sc.colors[vis], sc.labels[vis]
    out = {}
    for tag, vs, seed in (("a", 0.02, 5557), ("b", 0.05, 5558), ("c", 0.01, 5559)):
        np.random.seed(seed)
        vox = R.Voxelizer(**vox_kwargs(vs))
        st = np.random.get_state()
        m_v, m_r = vox.get_transformation_matrix()
        np.random.set_state(st)
        g, f, l, inv, inds = vox.voxelize(pts, col.copy(), lab.copy(), return_ind=True)
        out.update({f"rt_{tag}": m_r @ m_v, f"mr_{tag}": m_r, f"grid_{tag}": g, f"feats_{tag}": f,
                    f"labels_{tag}": l, f"inv_{tag}": inv, f"inds_{tag}": inds,
                    f"seed_{tag}": seed, f"vs_{tag}": vs})
    # no-augmentation voxelizer (plain scale matrix)
    vox = R.Voxelizer(voxel_size=0.04)
    g, f, l, inv, inds = vox.voxelize(pts, col.copy(), lab.copy(), return_ind=True)
    out.update(grid_plain=g, inv_plain=inv, inds_plain=inds)
    save("voxelize", xyz=pts, colors=col, labels=lab, **out)


class _Cfg:
    caption_contra_2d_pre = True


def golden_pool_scatter():
    torch.manual_seed(7)
    k, h, w, n, c = 12, 60, 80, 1500, 64
    rng = np.random.default_rng(5)
    x_label = torch.from_numpy(rng.integers(0, h, n))
    y_label = torch.from_numpy(rng.integers(0, w, n))
    part = torch.from_numpy(syn.make_partition_masks(9, k, h, w)).float()       # 0/1 floats
    over = torch.from_numpy(syn.make_overlap_logits(10, k, h, w, lo=32))
    emb = torch.randn(k, c)
    pred3d = torch.randn(n, c)
    ident2 = lambda a, b: a + 2.0 * b          # stand-in for the learned fuser (out of scope)
    ident = lambda a: a
    res = {}
    for tag, masks in (("part", part), ("over", over)):
        o, o2, o3, opre = R.mask_mapper([x_label], [y_label], [masks], [emb], [pred3d], ident2, ident, ident, _Cfg)
        res[f"fused_{tag}"], res[f"feat2d_{tag}"] = o[0].numpy(), o2[0].numpy()
        res[f"pre_{tag}"] = opre[0].numpy()
        # pooled means exactly as models/utils/criterion.py:83-85 + 152-157 computes them
        for mode, fn in (("ge", lambda t: t >= 0.5), ("sig", lambda t: t.sigmoid() >= 0.5)):
            m3 = fn(masks[:, x_label, y_label].clone())
            means = []
            for row in m3:
                means.append(torch.mean(pred3d[row], dim=0) if row.any() else torch.zeros(c))
            res[f"mean_{tag}_{mode}"] = torch.stack(means).numpy()
            res[f"cnt_{tag}_{mode}"] = m3.sum(1).numpy()
    # scalar pool, models/xmask3d.py:356-367
    score = torch.sigmoid(torch.randn(n)).view(1, -1)
    m3 = over[:, x_label, y_label].sigmoid() > 0.5
    res["score"] = score.numpy()
    res["score_pool"] = (torch.sum(score * m3, dim=1) / (torch.sum(m3, dim=1) + 1e-10)).numpy()
    res["score_keep"] = (torch.sum(m3, dim=1) > 0).numpy()
    save("pool", x_label=x_label.numpy(), y_label=y_label.numpy(), part=part.numpy().astype(np.uint8),
         over=over.numpy(), emb=emb.numpy(), pred3d=pred3d.numpy(), **res)


def golden_logits():
    out = {}
    for tag, (b, k, t) in {"b15": (2, 50, 20), "sn200": (1, 64, 201)}.items():
        me, te, ne = syn.make_embeddings(100 + t, b, k, t)
        labels = [[str(i)] for i in range(t - 1)]
        o = {"mask_embed": torch.from_numpy(me), "text_embed": torch.from_numpy(te),
             "null_embed": torch.from_numpy(ne), "labels": labels, "logit_scale": torch.tensor(1 / 0.07)}
        out[f"me_{tag}"], out[f"te_{tag}"], out[f"ne_{tag}"] = me, te, ne
        out[f"logits_{tag}"] = R.XMASK3d.cal_pred_logits(None, o).numpy()
    # synonym groups (ensemble max / mean), helper.py:72-97
    me, te, ne = syn.make_embeddings(7, 2, 10, 13)
    sizes = [1, 3, 2, 1, 5]
    labels, s = [], 0
    for g in sizes:
        labels.append([str(s + j) for j in range(g)])
        s += g
    o = {"mask_embed": torch.from_numpy(me), "text_embed": torch.from_numpy(te),
         "null_embed": torch.from_numpy(ne), "labels": labels, "logit_scale": torch.tensor(100.0)}
    out.update(me_grp=me, te_grp=te, ne_grp=ne, sizes_grp=np.array(sizes),
               logits_grp=R.XMASK3d.cal_pred_logits(None, o).numpy())
    raw = torch.randn(3, 7, 12)
    out.update(ens_in=raw.numpy(),
               ens_max=R.ensemble_logits_with_labels(raw, labels, "max").numpy(),
               ens_mean=R.ensemble_logits_with_labels(raw, labels, "mean").numpy())
    save("logits", **out)


if __name__ == "__main__":
    golden_hash()
    golden_project_voxelize()
    golden_pool_scatter()
    golden_logits()
