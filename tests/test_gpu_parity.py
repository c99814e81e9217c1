"""Parity of the CUDA path (through the C ABI) against (a) the golden fixtures produced by the
reference's own functions and (b) the CPU oracle on seeded synthetic inputs.

Bars: bit-exact for every integer / index / visibility output and for the float32 scatter-mean;
pooled features within 1e-5 (vector-wise max|a-b| / max|b| per mask, float32); identical logits
argmax.  Needs a B200: run with `pytest -m gpu`.
"""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

from xmask3d_b200 import synthetic as syn


@pytest.fixture(scope="module")
def dev():
    assert torch.cuda.is_available(), "GPU tests need a CUDA device"
    return torch.device("cuda", 0)


def _intr(k):
    return (k[0, 0], k[1, 1], k[0, 2], k[1, 2])


# ----------------------------------------------------------------------------- stage 2
def test_project_golden(golden, dev):
    from xmask3d_b200.fusion_util import PointCloudToImageMapper
    from xmask3d_b200.mapping_util import getMapping
    g = golden("project")
    m = getMapping()
    assert np.array_equal(m.intrinsics, g["intrinsics"])
    poses, dmm = g["poses"], g["depth_mm"]
    # float64 metres exactly as the loader hands them over (imread(png) / 1000)
    got = m.compute_mapping_batch(list(poses), g["xyz"], [d / 1000 for d in dmm])
    for v in range(poses.shape[0]):
        assert got[v].dtype == np.int64 and np.array_equal(got[v], g["mapping"][v])
    # raw uint16 millimetres (the fast path: depth image staged in shared memory)
    got = m.compute_mapping_batch(list(poses), g["xyz"], list(dmm))
    for v in range(poses.shape[0]):
        assert np.array_equal(got[v], g["mapping"][v])
    for v in range(poses.shape[0]):
        assert np.array_equal(m.compute_mapping(poses[v], g["xyz"], None), g["mapping_nodepth"][v])
    # depth image smaller than the camera image, cut_bound = 0 (fusion_util.py:105-135)
    m0 = PointCloudToImageMapper((320, 240), 0.25, 0, g["intrinsics"])
    sd = g["small_depth_mm"]
    assert np.array_equal(m0.compute_mapping(poses[0], g["xyz"], sd / 1000), g["mapping_small_cut0"])
    assert np.array_equal(m0.compute_mapping(poses[0], g["xyz"], sd), g["mapping_small_cut0"])


def _scene_views(seed, n_points, n_views):
    sc = syn.make_scene(seed, n_points)
    return sc, [syn.make_view(sc, v) for v in range(n_views)]


def test_project_vs_oracle_and_compaction(cport, dev):
    from xmask3d_b200 import ops
    intr = syn.scannet_intrinsics()
    scenes = [_scene_views(77, 150_000, 3), _scene_views(78, 41_237, 2)]      # ragged sizes, unaligned tail
    xyz = np.concatenate([s.xyz for s, _ in scenes])
    pt_off, n_pts, w2c, depth = [], [], [], []
    o = 0
    for sc, views in scenes:
        for vw in views:
            pt_off.append(o)
            n_pts.append(sc.xyz.shape[0])
            w2c.append(np.linalg.inv(vw.pose))
            depth.append(vw.depth_mm)
        o += sc.xyz.shape[0]
    views, out_off = ops.make_views(np.stack(w2c), intr, pt_off, n_pts, depth[0].shape)
    dten = torch.from_numpy(np.stack(depth).view(np.int16)).to(dev)
    pr = ops.project_batch(torch.from_numpy(xyz).to(dev), views, out_off, dten, want_mapping=True)
    assert int(pr.status.item()) == 0
    mapping = pr.mapping.cpu().numpy()
    vis = pr.vis.cpu().numpy()
    n_vis = pr.n_vis.cpu().numpy()
    vis_off = pr.vis_off.cpu().numpy()
    vis_idx, rowcol, xyz_vis = pr.vis_idx.cpu().numpy(), pr.rowcol.cpu().numpy(), pr.xyz_vis.cpu().numpy()
    i = 0
    for sc, vws in scenes:
        for vw in vws:
            ref = cport.project(sc.xyz, np.linalg.inv(vw.pose), intr, vw.depth_mm)
            a, b = out_off[i], out_off[i + 1]
            assert np.array_equal(mapping[a:b], ref), f"view {i}"
            assert np.array_equal(vis[a:b], ref[:, 2].astype(np.uint8))
            idx = np.nonzero(ref[:, 2])[0]
            assert n_vis[i] == len(idx) and vis_off[i + 1] - vis_off[i] == len(idx)
            s = slice(vis_off[i], vis_off[i + 1])
            assert np.array_equal(vis_idx[s], idx)                      # point order preserved
            assert np.array_equal(rowcol[s], ref[idx][:, :2])           # x_label (row), y_label (col)
            assert np.array_equal(xyz_vis[s], sc.xyz[idx])
            assert len(idx) > 400
            i += 1
    # capacity too small: flagged, never overrun
    pr2 = ops.project_batch(torch.from_numpy(xyz).to(dev), views, out_off, dten, cap_vis=1000)
    assert int(pr2.status.item()) & 1
    assert np.array_equal(pr2.vis_idx.cpu().numpy()[:1000], vis_idx[:1000])


def test_project_edge_cases(cport, dev):
    from xmask3d_b200 import ops
    intr = syn.scannet_intrinsics()
    rng = np.random.default_rng(5)
    # points behind the camera, at z ~ 0, NaN / inf coordinates, far outside the image
    xyz = rng.uniform(-3, 3, (5000, 3)).astype(np.float32)
    xyz[:50, 2] = 0.0
    xyz[50:60] = np.nan
    xyz[60:70] = np.inf
    xyz[70:80] = -np.inf
    xyz[80:90] = 1e30
    pose = np.eye(4)
    depth = np.full((240, 320), 2000, np.uint16)
    views, out_off = ops.make_views(np.linalg.inv(pose)[None], intr, [0], [5000], depth.shape)
    pr = ops.project_batch(torch.from_numpy(xyz).to(dev), views, out_off,
                           torch.from_numpy(depth.view(np.int16)).to(dev), want_mapping=True)
    with np.errstate(all="ignore"):
        ref = cport.project(xyz, np.linalg.inv(pose), intr, depth)
    assert np.array_equal(pr.mapping.cpu().numpy(), ref)
    # a view that sees nothing and an empty scene
    pose2 = np.eye(4)
    pose2[2, 3] = 100.0
    views, out_off = ops.make_views(np.stack([np.linalg.inv(pose2)]), intr, [0], [5000], depth.shape)
    pr = ops.project_batch(torch.from_numpy(xyz).to(dev), views, out_off,
                           torch.from_numpy(depth.view(np.int16)).to(dev))
    assert int(pr.n_vis[0].item()) == 0 and int(pr.vis_off[1].item()) == 0
    views, out_off = ops.make_views(np.stack([np.linalg.inv(pose)]), intr, [0], [0], depth.shape)
    pr = ops.project_batch(torch.from_numpy(xyz).to(dev), views, out_off,
                           torch.from_numpy(depth.view(np.int16)).to(dev))
    assert int(pr.vis_off[1].item()) == 0


# ----------------------------------------------------------------------------- stage 1
def test_voxelize_golden(golden, dev):
    from tests.golden.make_golden_params import vox_kwargs
    from xmask3d_b200.voxelizer import Voxelizer, voxelize_views
    g = golden("voxelize")
    for tag in "abc":
        grid, inds, inv = voxelize_views([g["xyz"]], [g[f"rt_{tag}"]])[0]
        assert grid.dtype == np.float64 and np.array_equal(grid, g[f"grid_{tag}"])
        assert np.array_equal(inds, g[f"inds_{tag}"]) and np.array_equal(inv, g[f"inv_{tag}"])
        # the full reference signature, drawing the matrix from the same np.random state
        np.random.seed(int(g[f"seed_{tag}"]))
        vox = Voxelizer(**vox_kwargs(float(g[f"vs_{tag}"])))
        grid, feats, labels, inv, inds = vox.voxelize(g["xyz"], g["colors"].copy(), g["labels"].copy(),
                                                      return_ind=True)
        assert np.array_equal(grid, g[f"grid_{tag}"]) and np.array_equal(inv, g[f"inv_{tag}"])
        assert np.array_equal(inds, g[f"inds_{tag}"]) and inv.dtype == np.int64
        assert np.array_equal(feats, g[f"feats_{tag}"]) and np.array_equal(labels, g[f"labels_{tag}"])
    grid, _, _, inv = Voxelizer(voxel_size=0.04).voxelize(g["xyz"], g["colors"].copy(), g["labels"].copy())
    assert np.array_equal(grid, g["grid_plain"]) and np.array_equal(inv, g["inv_plain"])
    link = np.arange(g["xyz"].shape[0]) * 3
    out = Voxelizer(voxel_size=0.04).voxelize(g["xyz"], g["colors"].copy(), g["labels"].copy(), link=link)
    assert np.array_equal(out[4], link[g["inds_plain"]])
    with pytest.raises(AssertionError):
        Voxelizer().voxelize(g["xyz"][:, :2], g["colors"], g["labels"])


def _random_rt(rng, voxel_size):
    from xmask3d_b200.voxelizer import M
    rot = M(np.array([0, 0, 1.0]), rng.uniform(-np.pi, np.pi)) @ M(np.array([1.0, 0, 0]), rng.uniform(-0.05, 0.05))
    rt = np.eye(4)
    rt[:3, :3] = rot * (rng.uniform(0.9, 1.1) / voxel_size)
    return rt


def test_voxelize_vs_oracle_batched(cport, dev):
    """Ragged batch through every sort path: tiny (bitonic), ~35k (2048-bucket sample sort),
    >65k unique keys (4096 buckets), a 1-point and an all-duplicates segment."""
    from xmask3d_b200 import ops
    rng = np.random.default_rng(11)
    sc = syn.make_scene(5, 150_000)
    big = syn.make_scene(6, 400_000, room=(12.0, 10.0, 3.0))
    segs = [(sc.xyz[:37_000], _random_rt(rng, 0.02)),
            (sc.xyz[:900], _random_rt(rng, 0.02)),
            (big.xyz, _random_rt(rng, 0.01)),
            (sc.xyz[:1], _random_rt(rng, 0.02)),
            (np.repeat(sc.xyz[:3], 2000, 0), _random_rt(rng, 0.02)),
            (sc.xyz[50_000:50_000 + 4097], _random_rt(rng, 0.005)),
            (sc.xyz, _random_rt(rng, 0.05))]
    n = [s.shape[0] for s, _ in segs]
    off = np.concatenate([[0], np.cumsum(n)]).astype(np.int64)
    xyz = torch.from_numpy(np.concatenate([s for s, _ in segs])).to(dev)
    rt = torch.from_numpy(np.stack([r[:3, :4] for _, r in segs])).to(dev)
    for collate in (False, True):
        u = ops.voxelize_batch(xyz, torch.from_numpy(off).to(dev), rt, collate=collate)
        assert int(u.status.item()) == 0
        m, uoff = u.m.cpu().numpy(), u.uniq_off.cpu().numpy()
        first, inv, vox = u.first.cpu().numpy(), u.inverse.cpu().numpy(), u.voxel_xyz.cpu().numpy()
        for i, (pts, r) in enumerate(segs):
            rgrid, rfirst, rinv = cport.voxelize(pts, r)
            assert m[i] == len(rfirst), f"segment {i}"
            a = int(uoff[i])
            assert np.array_equal(first[a:a + m[i]], rfirst), f"segment {i}"
            assert np.array_equal(vox[a:a + m[i]].astype(np.float64), rgrid), f"segment {i}"
            shift = a if collate else 0
            assert np.array_equal(inv[off[i]:off[i + 1]] - shift, rinv), f"segment {i}"
        assert uoff[-1] == m.sum()
    assert m[2] > 65536 and 4096 < m[0] < 65536 and m[3] == 1 and m[4] == 3


@pytest.mark.parametrize("mode,unit_pts", [(0, 0), (0, 1200), (0, 64), (1, 0)])
def test_voxelize_paths_vs_oracle(cport, dev, mode, unit_pts):
    """The shared-memory unit path (default; several key-range units per segment when unit_pts is
    small), the multi-kernel path (mode 1) and the overflow fallback all give the oracle's result:
    ragged batch with an empty, a 1-point, an all-duplicates, a very fine and a coarse segment."""
    from xmask3d_b200 import ops
    rng = np.random.default_rng(12)
    sc = syn.make_scene(7, 150_000)
    if unit_pts == 64:          # 64-point units: keep every segment below 32 units
        segs = [(sc.xyz[:2000], _random_rt(rng, 0.02)), (sc.xyz[:0], _random_rt(rng, 0.02)),
                (sc.xyz[3000:3001], _random_rt(rng, 0.02)), (np.repeat(sc.xyz[:3], 600, 0), _random_rt(rng, 0.02)),
                (sc.xyz[5000:5000 + 1025], _random_rt(rng, 0.005)), (sc.xyz[:64], _random_rt(rng, 0.02)),
                (sc.xyz[:65], _random_rt(rng, 0.02))]
    else:
        segs = [(sc.xyz[:37_000], _random_rt(rng, 0.02)), (sc.xyz[:0], _random_rt(rng, 0.02)),
                (sc.xyz[:900], _random_rt(rng, 0.02)), (sc.xyz[:1], _random_rt(rng, 0.02)),
                (np.repeat(sc.xyz[:3], 2000, 0), _random_rt(rng, 0.02)),
                (sc.xyz[50_000:50_000 + 8193], _random_rt(rng, 0.005)),
                (sc.xyz[:14_000], _random_rt(rng, 0.02)), (sc.xyz[:0], _random_rt(rng, 0.02))]
        if unit_pts == 0:
            segs.append((sc.xyz, _random_rt(rng, 0.05)))
    n = [s.shape[0] for s, _ in segs]
    off = np.concatenate([[0], np.cumsum(n)]).astype(np.int64)
    xyz = torch.from_numpy(np.concatenate([s for s, _ in segs])).to(dev)
    rt = torch.from_numpy(np.stack([r[:3, :4] for _, r in segs])).to(dev)
    if True:                                     # the path is a per-call argument (re-entrant, no process-global switch)
        for collate in (False, True):
            u = ops.voxelize_batch(xyz, torch.from_numpy(off).to(dev), rt, collate=collate, cap=xyz.shape[0] + 777,
                                   mode=mode, unit_pts=unit_pts)
            assert int(u.status.item()) == 0
            assert ops.voxel_path_info(u) == ((1, 0) if mode == 1 else (0, 0))
            m, uoff = u.m.cpu().numpy(), u.uniq_off.cpu().numpy()
            first, inv, vox = u.first.cpu().numpy(), u.inverse.cpu().numpy(), u.voxel_xyz.cpu().numpy()
            for i, (pts, r) in enumerate(segs):
                if len(pts) == 0:
                    assert m[i] == 0 and uoff[i + 1] == uoff[i]
                    continue
                rgrid, rfirst, rinv = cport.voxelize(pts, r)
                assert m[i] == len(rfirst), f"segment {i}"
                a = int(uoff[i])
                assert np.array_equal(first[a:a + m[i]], rfirst), f"segment {i}"
                assert np.array_equal(vox[a:a + m[i]].astype(np.float64), rgrid), f"segment {i}"
                shift = a if collate else 0
                assert np.array_equal(inv[off[i]:off[i + 1]] - shift, rinv), f"segment {i}"
            assert uoff[-1] == m.sum()


def test_voxelize_fast_only_path(cport, dev):
    """XM3D_VOX_FAST_ONLY: the shared-memory units alone (the fallback kernels are not launched) give the oracle's
    result; a batch that needs the multi-kernel path — a segment above 224 k points, a unit that overflows — is flagged
    (FLAG_VOX_FALLBACK) instead of being recomputed."""
    from xmask3d_b200 import _lib as L, ops
    rng = np.random.default_rng(21)
    sc = syn.make_scene(7, 150_000)
    segs = [(sc.xyz[:37_000], _random_rt(rng, 0.02)), (sc.xyz[:0], _random_rt(rng, 0.02)), (sc.xyz[40_000:41_000], _random_rt(rng, 0.02))]
    off = np.concatenate([[0], np.cumsum([len(s_) for s_, _ in segs])]).astype(np.int64)
    xyz = torch.from_numpy(np.concatenate([s_ for s_, _ in segs])).to(dev)
    rt = torch.from_numpy(np.stack([r[:3, :4] for _, r in segs])).to(dev)
    u = ops.voxelize_batch(xyz, torch.from_numpy(off).to(dev), rt, collate=True, mode=2)
    assert int(u.status.item()) == 0
    m, uoff = u.m.cpu().numpy(), u.uniq_off.cpu().numpy()
    for i, (p, r) in enumerate(segs):
        if len(p) == 0:
            continue
        rgrid, rfirst, rinv = cport.voxelize(p, r)
        a = int(uoff[i])
        assert m[i] == len(rfirst) and np.array_equal(u.first.cpu().numpy()[a:a + m[i]], rfirst)
        assert np.array_equal(u.voxel_xyz.cpu().numpy()[a:a + m[i]].astype(np.float64), rgrid)
        assert np.array_equal(u.inverse.cpu().numpy()[off[i]:off[i + 1]] - a, rinv)
    big = syn.make_scene(6, 300_000, room=(12.0, 10.0, 3.0))
    u = ops.voxelize_batch(torch.from_numpy(big.xyz).to(dev), torch.tensor([0, 300_000], device=dev),
                           torch.from_numpy(_random_rt(rng, 0.01)[None, :3, :4]).to(dev), mode=2)
    assert int(u.status.item()) & L.FLAG_VOX_FALLBACK            # not eligible: > 224 k points in one segment
    n = 20 * 1024
    pts = syn.make_scene(8, 60_000).xyz[:n].copy()
    pts[np.arange(1024) * 20] = pts[0]                              # defeats the key-range split: a unit overflows
    u = ops.voxelize_batch(torch.from_numpy(pts).to(dev), torch.tensor([0, n], device=dev),
                           torch.from_numpy(_random_rt(rng, 0.004)[None, :3, :4]).to(dev), mode=2)
    assert int(u.status.item()) & L.FLAG_VOX_FALLBACK


def test_voxelize_unit_overflow_falls_back(cport, dev):
    """A segment whose 1024 sample positions all hold the same voxel defeats the key-range split:
    one unit receives ~20 k distinct keys, overflows its 8704-key table, and the batch is recomputed by
    the multi-kernel path (same result, ctl = (0, 1)).  Keys-only variant through unique_batch too."""
    from xmask3d_b200 import ops
    rng = np.random.default_rng(13)
    sc = syn.make_scene(8, 60_000)
    n = 20 * 1024
    pts = sc.xyz[:n].copy()
    pts[np.arange(1024) * 20] = pts[0]
    segs = [(sc.xyz[30_000:36_000], _random_rt(rng, 0.02)), (pts, _random_rt(rng, 0.004))]
    off = np.concatenate([[0], np.cumsum([len(s) for s, _ in segs])]).astype(np.int64)
    xyz = torch.from_numpy(np.concatenate([s for s, _ in segs])).to(dev)
    rt = torch.from_numpy(np.stack([r[:3, :4] for _, r in segs])).to(dev)
    u = ops.voxelize_batch(xyz, torch.from_numpy(off).to(dev), rt)
    assert ops.voxel_path_info(u) == (0, 1)
    m, uoff = u.m.cpu().numpy(), u.uniq_off.cpu().numpy()
    for i, (p, r) in enumerate(segs):
        rgrid, rfirst, rinv = cport.voxelize(p, r)
        a = int(uoff[i])
        assert m[i] == len(rfirst)
        assert np.array_equal(u.first.cpu().numpy()[a:a + m[i]], rfirst)
        assert np.array_equal(u.voxel_xyz.cpu().numpy()[a:a + m[i]].astype(np.float64), rgrid)
        assert np.array_equal(u.inverse.cpu().numpy()[off[i]:off[i + 1]], rinv)
    assert m[1] > 2 * 8704
    # keys given (np.unique drop-in): fast path, several units per segment, adversarial keys
    ksegs = [np.sort(rng.integers(0, 2 ** 63, 30_000, dtype=np.int64)).astype(np.uint64) * np.uint64(2),
             np.full(5000, 12345, np.uint64), rng.integers(0, 40, 10_000).astype(np.uint64),
             np.array([2 ** 64 - 2, 0, 2 ** 63, 0], np.uint64), np.zeros(0, np.uint64)]
    koff = np.concatenate([[0], np.cumsum([len(s) for s in ksegs])]).astype(np.int64)
    keys = torch.from_numpy(np.concatenate(ksegs).view(np.int64)).to(dev)
    for unit_pts in (0, 300):
        ku = ops.unique_batch(keys, torch.from_numpy(koff).to(dev), collate=True, unit_pts=unit_pts)
        if unit_pts == 0:
            assert ops.voxel_path_info(ku) == (0, 0)
        km, kuoff = ku.m.cpu().numpy(), ku.uniq_off.cpu().numpy()
        for i, ks in enumerate(ksegs):
            rf, ri_, rc = cport.unique_u64(ks) if len(ks) else (np.zeros(0, np.int64),) * 3
            a = int(kuoff[i])
            assert km[i] == len(rf)
            assert np.array_equal(ku.first.cpu().numpy()[a:a + km[i]], rf)
            assert np.array_equal(ku.inverse.cpu().numpy()[koff[i]:koff[i + 1]] - a, ri_)


def test_sparse_quantize_and_hashes_golden(golden, cport, dev):
    from xmask3d_b200.voxelization_utils import fnv_hash_vec, ravel_hash_vec, sparse_quantize
    g = golden("hash")
    assert np.array_equal(fnv_hash_vec(g["kat"]), g["kat_fnv"])
    assert np.array_equal(fnv_hash_vec(g["rnd"]), g["rnd_fnv"])
    assert np.array_equal(ravel_hash_vec(g["rnd"] - 100.0), g["rnd_ravel"])
    inds, inv = sparse_quantize(g["ex"], return_index=True)
    assert inds.tolist() == [3, 0, 1] and inv.tolist() == [1, 2, 1, 0, 2]
    li, ll = sparse_quantize(g["dup"], labels=g["dup_labels"].copy(), return_index=True,
                             set_ignore_label_when_collision=True)
    assert np.array_equal(li, g["dup_lab_inds"]) and np.array_equal(ll, g["dup_lab_out"])
    ri, rinv = sparse_quantize(g["dup"] - 5.0, return_index=True, hash_type="ravel")
    assert np.array_equal(ri, g["dup_ravel_inds"]) and np.array_equal(rinv, g["dup_ravel_inv"])
    q = golden("quant")
    qi, qinv = sparse_quantize(q["coords"], return_index=True, quantization_size=0.05)       # negative cells
    assert np.array_equal(qi, q["inds"]) and np.array_equal(qinv, q["inv"])
    with pytest.raises(AssertionError):
        sparse_quantize(g["ex"], hash_type="md5")
    with pytest.raises(AssertionError):
        sparse_quantize(g["ex"][:, 0])
    # forced 64-bit collisions and adversarial (sorted / constant) keys through unique_batch
    from xmask3d_b200 import ops
    rng = np.random.default_rng(2)
    segs = [np.sort(rng.integers(0, 2 ** 63, 70_000, dtype=np.int64)).astype(np.uint64) * np.uint64(2),
            np.full(5000, 12345, np.uint64),
            rng.integers(0, 40, 10_000).astype(np.uint64),
            np.array([2 ** 64 - 2, 0, 2 ** 63, 0], np.uint64)]
    off = np.concatenate([[0], np.cumsum([len(s) for s in segs])]).astype(np.int64)
    keys = torch.from_numpy(np.concatenate(segs).view(np.int64)).to(dev)
    u = ops.unique_batch(keys, torch.from_numpy(off).to(dev), want_counts=True)
    m, uoff = u.m.cpu().numpy(), u.uniq_off.cpu().numpy()
    for i, s in enumerate(segs):
        rf, ri_, rc = cport.unique_u64(s)
        a = int(uoff[i])
        assert m[i] == len(rf)
        assert np.array_equal(u.first.cpu().numpy()[a:a + m[i]], rf)
        assert np.array_equal(u.counts.cpu().numpy()[a:a + m[i]], rc)
        assert np.array_equal(u.inverse.cpu().numpy()[off[i]:off[i + 1]], ri_)


# ----------------------------------------------------------------------------- stage 3
def test_gather_pool_scatter_golden(golden, cport, dev):
    from xmask3d_b200 import ops
    from xmask3d_b200.fuser import mask_mapper
    g = golden("pool")
    xl, yl = torch.from_numpy(g["x_label"]), torch.from_numpy(g["y_label"])
    emb, pred3d = torch.from_numpy(g["emb"]).to(dev), torch.from_numpy(g["pred3d"]).to(dev)

    class Cfg:
        caption_contra_2d_pre = True
    for tag, masks in (("part", torch.from_numpy(g["part"]).float()), ("over", torch.from_numpy(g["over"]))):
        fused, f2d, f3d, pre = mask_mapper([xl], [yl], [masks.to(dev)], [emb], [pred3d],
                                           lambda a, b: a + 2.0 * b, lambda a: a, lambda a: a, Cfg)
        assert np.array_equal(f2d[0].cpu().numpy(), g[f"feat2d_{tag}"])          # bit-exact float32
        assert np.array_equal(fused[0].cpu().numpy(), g[f"fused_{tag}"])
        assert np.array_equal(pre[0].cpu().numpy(), g[f"pre_{tag}"])
        assert np.array_equal(f3d[0].cpu().numpy(), g["pred3d"])
        seg = torch.tensor([0, len(xl)], dtype=torch.int64, device=dev)
        rowcol = torch.stack([xl, yl], 1).to(dev, torch.int32)
        for mode, key in (("ge0.5", "ge"), ("sigmoid_ge0.5", "sig")):
            member, counts = ops.gather_masks(masks.to(dev)[None], rowcol, seg, mode=mode, want_counts=True)
            s, cnt, mean = ops.pool(pred3d, seg, masks.shape[0], member=member)
            assert np.array_equal(cnt[0].cpu().numpy(), g[f"cnt_{tag}_{key}"])
            assert np.array_equal(counts[0].cpu().numpy(), g[f"cnt_{tag}_{key}"])
            ref = g[f"mean_{tag}_{key}"]
            err = np.abs(mean[0].cpu().numpy() - ref).max(1) / np.maximum(np.abs(ref).max(1), 1e-30)
            assert err.max() < 1e-5, err.max()
    from xmask3d_b200.pooling import masked_score_pool
    sp, keep = masked_score_pool(torch.from_numpy(g["score"]).to(dev), [xl], [yl],
                                 torch.from_numpy(g["over"]).to(dev)[None])
    assert np.array_equal(keep[0].cpu().numpy(), g["score_keep"])
    np.testing.assert_allclose(sp[0].cpu().numpy(), g["score_pool"], rtol=1e-5, atol=1e-7)


def test_sigmoid_threshold_near_zero(dev):
    """`sigmoid(x) >= 0.5` / `> 0.5` (models/utils/criterion.py:83-85, models/xmask3d.py:356-357) is decided by the sign of x
    outside |x| <= 1e-6 and by the exact float32 formula inside: identical to torch's float32 sigmoid on the CPU for
    every |x| >= 1e-7 (below that the CPU's vectorised exp and the GPU's expf may round differently)."""
    from xmask3d_b200 import ops
    mags = np.array([1e-7, 1.2e-7, 2.5e-7, 5e-7, 9.9e-7, 1.0e-6, 1.01e-6, 2e-6, 1e-5, 1e-3, 0.5, 3.0, 20.0, 100.0], np.float32)
    vals = np.concatenate([mags, -mags, [np.inf, -np.inf, np.nan]]).astype(np.float32)
    k, h, w = len(vals), 4, 8
    masks = torch.from_numpy(np.broadcast_to(vals[:, None, None], (k, h, w)).copy())[None]
    rowcol = torch.tensor([[1, 2], [3, 7]], dtype=torch.int32, device=dev)
    seg = torch.tensor([0, 2], dtype=torch.int64, device=dev)
    for mode, ref in (("sigmoid_ge0.5", torch.from_numpy(vals).sigmoid() >= 0.5), ("sigmoid_gt0.5", torch.from_numpy(vals).sigmoid() > 0.5)):
        member, _ = ops.gather_masks(masks.to(dev), rowcol, seg, mode=mode)
        bits = member.cpu().numpy().view(np.uint32)
        got = np.array([(bits[0, m // 32] >> (m % 32)) & 1 for m in range(k)], bool)
        assert np.array_equal(got, ref.numpy()), (mode, vals[got != ref.numpy()])
        assert np.array_equal(bits[0], bits[1])


def test_mask_mapper_ragged_lists(dev):
    """The caller's real input (models/utils/criterion.py:262-340): per-scene lists with DIFFERENT numbers of masks,
    bool partition masks for some scenes, float32 zero masks ("nothing kept") for others, one scene whose masks
    hit no point at all (fuser.py:19-20 guard) — bit-exact against the restatement of fuser.py:6-53 that
    tests/test_reference_live.py pins to the reference on exactly such lists."""
    from oracle import ref_port
    from xmask3d_b200.fuser import mask_mapper
    g = torch.Generator().manual_seed(11)
    h, w, c = 240, 320, 64
    xs, ys, masks, embs, preds = [], [], [], [], []
    for i, k in enumerate([7, 3, 12, 5, 4]):
        n = [900, 1, 2500, 333, 50][i]
        xs.append(torch.randint(0, h, (n,), generator=g))
        ys.append(torch.randint(0, w, (n,), generator=g))
        if i == 2:
            masks.append(torch.zeros(k, h, w))                                   # float zeros: nothing kept
        elif i == 4:
            m = torch.zeros(k, h, w, dtype=torch.bool)
            m[:, 0, 0] = True                                                    # bool masks that miss every point
            xs[-1].clamp_(min=1)
            masks.append(m)
        else:
            lab = torch.randint(0, k + 2, (h, w), generator=g)                   # partition; labels >= k = background
            masks.append(lab.unsqueeze(0) == torch.arange(k).view(k, 1, 1))
        embs.append(torch.randn(k, c, generator=g))
        preds.append(torch.randn(n, c, generator=g))
    ident = torch.nn.Identity()
    fuse = lambda a, b: a * 0.5 + b                          # noqa: E731

    class Cfg:
        caption_contra_2d_pre = True
    ref = ref_port.mask_mapper_ref(xs, ys, masks, embs, preds, fuse, ident, ident, True)
    got = mask_mapper(xs, ys, [m.to(dev) for m in masks], [e.to(dev) for e in embs], [p.to(dev) for p in preds],
                      fuse, ident, ident, Cfg)
    for a, b in zip(ref, got):
        assert len(a) == len(b)
        for u, v in zip(a, b):
            assert torch.equal(u, v.cpu())


def test_voxelize_float64_coords(dev):
    """The augmented training path hands the voxelizer FLOAT64 coordinates (ElasticDistortion returns
    coords + interp(coords) * magnitude, dataset/augmentation.py:171; data_loader.py:252-256): homo_coords then stays
    float64 and the matmul sees the unrounded values.  Device result vs the numpy restatement of voxelizer.py:81-132
    on the same float64 input — and it must differ from what rounding the input to float32 first would give."""
    from oracle import ref_port
    from tests.golden.make_golden_params import vox_kwargs
    from xmask3d_b200.voxelizer import Voxelizer
    rng = np.random.default_rng(3)
    sc = syn.make_scene(9, 60_000)
    coords64 = sc.xyz.astype(np.float64) + rng.normal(0, 0.01, sc.xyz.shape)     # not representable in float32
    kw = vox_kwargs(0.02)
    for seed in (1, 2):
        np.random.seed(seed)
        ref = ref_port.Voxelizer(**kw).voxelize(coords64.copy(), sc.colors.copy(), sc.labels.copy(), return_ind=True)
        np.random.seed(seed)
        got = Voxelizer(**kw).voxelize(coords64.copy(), sc.colors.copy(), sc.labels.copy(), return_ind=True)
        for a, b in zip(ref, got):
            assert np.asarray(a).dtype == np.asarray(b).dtype and np.array_equal(a, b)
    np.random.seed(2)
    lossy = Voxelizer(**kw).voxelize(coords64.astype(np.float32), sc.colors.copy(), sc.labels.copy(), return_ind=True)
    assert lossy[0].shape != got[0].shape or not np.array_equal(lossy[3], got[3]), \
        "rounding the coordinates to float32 first must change some voxel assignment at 60k points"


@pytest.mark.parametrize("k,c", [(50, 768), (100, 768), (7, 64), (130, 256), (3, 1), (96, 128), (33, 256)])
def test_pool_scatter_vs_oracle(cport, dev, k, c):
    """Ragged batch (incl. an empty segment), overlapping members and partition labels."""
    from xmask3d_b200 import ops
    rng = np.random.default_rng(k * 1000 + c)
    n = [3000, 0, 1777, 1, 5200, 320]          # 320 = two full tiles of the point-major kernel
    off = np.concatenate([[0], np.cumsum(n)]).astype(np.int64)
    total = int(off[-1])
    feat = rng.standard_normal((total, c), dtype=np.float32)
    member_b = rng.random((len(n), k, max(n))) < 0.12
    label = rng.integers(-1, k, total).astype(np.int32)
    words = (k + 31) // 32
    packed = np.zeros((total, words), np.uint32)
    for s in range(len(n)):
        mb = member_b[s][:, :n[s]]
        for m in range(k):
            packed[off[s]:off[s + 1], m // 32] |= (mb[m].astype(np.uint32) << np.uint32(m % 32))
    seg = torch.from_numpy(off).to(dev)
    f = torch.from_numpy(feat).to(dev)
    mem = torch.from_numpy(packed.view(np.int32)).to(dev)
    s_sum, cnt, mean = ops.pool(f, seg, k, member=mem)
    l_sum, l_cnt, l_mean = ops.pool(f, seg, k, label=torch.from_numpy(label).to(dev))
    emb = rng.standard_normal((len(n), k, c), dtype=np.float32)
    out, counter = ops.scatter(torch.from_numpy(emb).to(dev), seg, total, member=mem)
    lout, lcounter = ops.scatter(torch.from_numpy(emb).to(dev), seg, total, label=torch.from_numpy(label).to(dev))
    for s in range(len(n)):
        a, b = off[s], off[s + 1]
        mb = member_b[s][:, :n[s]]
        r_sum, r_cnt = cport.pool_member_f64(feat[a:b], mb) if n[s] else (np.zeros((k, c)), np.zeros(k, np.int64))
        assert np.array_equal(cnt[s].cpu().numpy(), r_cnt)
        scale = np.maximum(np.abs(r_sum).max(1), 1e-30)
        assert (np.abs(s_sum[s].cpu().numpy() - r_sum).max(1) / scale).max() < 1e-5
        r_mean = r_sum / np.maximum(r_cnt, 1)[:, None]
        assert (np.abs(mean[s].cpu().numpy() - r_mean).max(1) / np.maximum(np.abs(r_mean).max(1), 1e-30)).max() < 1e-5
        rl_sum, rl_cnt = cport.pool_label_f64(feat[a:b], label[a:b], k) if n[s] else (np.zeros((k, c)), np.zeros(k, np.int64))
        assert np.array_equal(l_cnt[s].cpu().numpy(), rl_cnt)
        assert (np.abs(l_sum[s].cpu().numpy() - rl_sum).max(1) / np.maximum(np.abs(rl_sum).max(1), 1e-30)).max() < 1e-5
        if n[s]:
            r_out, r_counter = cport.scatter_member_f32(mb, emb[s])
            assert np.array_equal(out[a:b].cpu().numpy(), r_out)                    # bit-exact
            assert np.array_equal(counter[a:b].cpu().numpy(), r_counter)
            onehot = (label[a:b][None, :] == np.arange(k)[:, None])
            rl_out, _ = cport.scatter_member_f32(onehot, emb[s])
            assert np.array_equal(lout[a:b].cpu().numpy(), rl_out)
    # determinism: same bits on a second run
    s2, _, _ = ops.pool(f, seg, k, member=mem)
    assert torch.equal(s2, s_sum)
    # row_index indirection (pred_3d[inds_reconstruct], models/xmask3d.py:152)
    perm = torch.randperm(total, device=dev).to(torch.int32)
    inv = torch.empty_like(perm)
    inv[perm.long()] = torch.arange(total, device=dev, dtype=torch.int32)
    path = "rows" if (c % 128 == 0 and k <= 96) else "pair_lists"      # kernels that take row_index
    s3, _, _ = ops.pool(f[perm.long()], seg, k, member=mem, row_index=inv, path=path)
    s4, _, _ = ops.pool(f, seg, k, member=mem, path=path)
    assert torch.equal(s3, s4)
    assert float(((s4 - s_sum).abs().amax(-1) / s_sum.abs().amax(-1).clamp_min(1e-30)).max()) < 1e-5


@pytest.mark.parametrize("path", ["pair_lists", "rows", "mma"])
@pytest.mark.parametrize("ns,k,c,dens", [([3000, 0, 1777, 1, 5200, 320], 50, 768, 0.12), ([3000, 64, 65, 63], 96, 128, 0.3),
                                         ([2500, 700], 33, 256, 0.5), ([70_000], 50, 256, 0.3),
                                         ([5000, 4000], 100, 384, 0.1), ([1000], 1, 128, 1.0), ([129], 64, 128, 0.0)])
def test_pool_paths_vs_oracle(cport, dev, path, ns, k, c, dens):
    """Every pooling kernel — sorted pair lists, point-major CUDA cores, tensor cores (tcgen05 tf32 hi/lo with the
    membership bits as a 0/1 operand, csrc/pool_mma.cu) — against the float64 oracle of
    models/utils/criterion.py:148-157 on overlapping masks: counts exact, sums / means <= 1e-5 vector-wise,
    deterministic, rows of a following segment (here 1e30) never leak into a segment's last tile."""
    from xmask3d_b200 import ops
    if path == "rows" and k > 96:
        pytest.skip("point-major kernel holds at most 96 accumulator rows")
    rng = np.random.default_rng(k * 7 + c)
    off = np.concatenate([[0], np.cumsum(ns)]).astype(np.int64)
    total = int(off[-1])
    feat = rng.standard_normal((total + 64, c), dtype=np.float32)
    feat[total:] = 1e30                                   # rows past the last segment
    if len(ns) > 1 and ns[1] > 0:
        feat[off[1]:off[1] + 8] *= 1e6                    # a loud neighbour right after segment 0
    member_b = [rng.random((k, n)) < dens for n in ns]
    packed = np.zeros((total + 64, (k + 31) // 32), np.uint32)
    packed[total:] = 0xFFFFFFFF                           # bits past the last segment must be ignored
    for s_, mb in enumerate(member_b):
        for m in range(k):
            packed[off[s_]:off[s_ + 1], m // 32] |= (mb[m].astype(np.uint32) << np.uint32(m % 32))
    seg = torch.from_numpy(off).to(dev)
    f = torch.from_numpy(feat).to(dev)
    mem = torch.from_numpy(packed.view(np.int32)).to(dev)
    pairs = int(sum(mb.sum() for mb in member_b))
    s_sum, cnt, mean = ops.pool(f, seg, k, member=mem, cap=total, cap_pairs=max(pairs, total + 2), path=path)
    for s_, mb in enumerate(member_b):
        a, b = off[s_], off[s_ + 1]
        r_sum, r_cnt = cport.pool_member_f64(feat[a:b], mb) if ns[s_] else (np.zeros((k, c)), np.zeros(k, np.int64))
        assert np.array_equal(cnt[s_].cpu().numpy(), r_cnt), f"segment {s_}"
        scale = np.maximum(np.abs(r_sum).max(1), 1e-30)
        assert (np.abs(s_sum[s_].cpu().numpy() - r_sum).max(1) / scale).max() < 1e-5, f"segment {s_}"
        r_mean = r_sum / np.maximum(r_cnt, 1)[:, None]
        assert (np.abs(mean[s_].cpu().numpy() - r_mean).max(1) / np.maximum(np.abs(r_mean).max(1), 1e-30)).max() < 1e-5
    s2, c2, m2 = ops.pool(f, seg, k, member=mem, cap=total, cap_pairs=max(pairs, total + 2), path=path)
    assert torch.equal(s2, s_sum) and torch.equal(c2, cnt) and torch.equal(m2, mean)


@pytest.mark.parametrize("n_seg,k,c,max_n", [(2000, 50, 256, 300), (5000, 128, 128, 150), (300, 64, 1024, 2000)])
def test_pool_mma_many_small_segments(dev, n_seg, k, c, max_n):
    """Work-item ring of the tensor-core kernel under stress: thousands of segments of 0 .. max_n points (most items
    shorter than a few 64-point tiles, many empty, more segments than the size-ordered hand-out covers in the second
    case), K = 128 with all four membership words, eight 128-channel slices — against the pair-list kernel (itself
    pinned to the float64 oracle above): counts exact, sums / means <= 1e-5 vector-wise, deterministic."""
    from xmask3d_b200 import ops
    g = torch.Generator(device="cpu").manual_seed(n_seg + k)
    n = torch.randint(0, max_n + 1, (n_seg,), generator=g)
    n[torch.rand(n_seg, generator=g) < 0.1] = 0
    off = torch.cat([torch.zeros(1, dtype=torch.int64), n.cumsum(0)]).to(dev)
    total = int(off[-1])
    f = torch.randn(total, c, generator=g).to(dev)
    words = (k + 31) // 32
    mem = torch.randint(-2 ** 31, 2 ** 31 - 1, (total, words), generator=g, dtype=torch.int64).to(torch.int32)
    mem &= torch.randint(-2 ** 31, 2 ** 31 - 1, (total, words), generator=g, dtype=torch.int64).to(torch.int32)   # density 1/4
    mem = mem.to(dev)
    cap_pairs = total * k
    a_sum, a_cnt, a_mean = ops.pool(f, off, k, member=mem, cap_pairs=cap_pairs, path="mma")
    b_sum, b_cnt, b_mean = ops.pool(f, off, k, member=mem, cap_pairs=cap_pairs, path="pair_lists")
    assert torch.equal(a_cnt, b_cnt)
    for x, y in ((a_sum, b_sum), (a_mean, b_mean)):
        err = (x - y).abs().amax(-1) / y.abs().amax(-1).clamp_min(1e-30)
        assert float(err.max()) < 1e-5
    a2, c2, m2 = ops.pool(f, off, k, member=mem, cap_pairs=cap_pairs, path="mma")
    assert torch.equal(a2, a_sum) and torch.equal(c2, a_cnt) and torch.equal(m2, a_mean)


def test_pool_path_selection(dev):
    """XM3D_POOL_AUTO: overlapping memberships (cap_pairs > cap + 1) go to the tensor-core kernel when it is eligible;
    forcing a path that is not eligible fails loudly (no silent fallback)."""
    from xmask3d_b200 import _lib as L, ops
    f = torch.randn(300, 96, device=dev)
    seg = torch.tensor([0, 300], device=dev)
    mem = torch.randint(0, 2 ** 20, (300, 1), device=dev, dtype=torch.int32)
    with pytest.raises(L.Xm3dError):
        ops.pool(f, seg, 20, member=mem, path="mma")              # c % 128 != 0
    with pytest.raises(L.Xm3dError):
        ops.pool(f, seg, 20, member=mem, path="rows")
    a, _, _ = ops.pool(f, seg, 20, member=mem)                    # auto -> pair lists
    b, _, _ = ops.pool(f, seg, 20, member=mem, path="pair_lists")
    assert torch.equal(a, b)
    f = torch.randn(300, 128, device=dev)
    a, _, _ = ops.pool(f, seg, 20, member=mem)                    # auto -> tensor cores (overlap, eligible)
    b, _, _ = ops.pool(f, seg, 20, member=mem, path="mma")
    assert torch.equal(a, b)


@pytest.mark.parametrize("bad", [float("nan"), float("inf"), float("-inf"), "-nan"])
@pytest.mark.parametrize("k", [20, 100])
def test_pool_mma_flags_nonfinite_features(dev, bad, k):
    """A NaN / Inf feature multiplied by a 0 membership would reach masks it does not belong to inside the tensor core:
    the kernel raises XM3D_FLAG_NONFINITE (and only then), and the CUDA-core path gives the reference's result (NaN only
    in the masks that hold the point)."""
    from xmask3d_b200 import _lib as L, ops
    g = torch.Generator(device="cpu").manual_seed(5)
    n, c = 1000, 256
    f = torch.randn(n, c, generator=g).to(dev)
    seg = torch.tensor([0, 400, n], device=dev)
    words = (k + 31) // 32
    bits = (torch.rand(n, k, generator=g) < 0.2)
    bits[777] = False
    bits[777, 3] = True                                           # the bad point belongs to mask 3 of segment 1 only
    mem = np.zeros((n, words), dtype=np.uint32)
    for m in range(k):
        mem[:, m // 32] |= bits[:, m].numpy().astype(np.uint32) << np.uint32(m % 32)
    mem = torch.from_numpy(mem.view(np.int32)).to(dev)
    st = torch.zeros(1, dtype=torch.int32, device=dev)
    ops.pool(f, seg, k, member=mem, path="mma", status=st, cap_pairs=n * k)
    assert int(st.item()) == 0                                    # finite features: no flag
    if bad == "-nan":
        f[777, 130:131] = torch.from_numpy(np.array([0xffc00001], dtype=np.uint32).view(np.float32)).to(dev)
    else:
        f[777, 130] = bad
    ops.pool(f, seg, k, member=mem, path="mma", status=st, cap_pairs=n * k)
    assert int(st.item()) & L.FLAG_NONFINITE
    st.zero_()
    s, cnt, _ = ops.pool(f, seg, k, member=mem, path="rows" if k <= 96 else "pair_lists", status=st, cap_pairs=n * k)
    assert int(st.item()) == 0
    nonfin = ~torch.isfinite(s)
    assert int(nonfin.sum()) == 1 and bool(nonfin[1, 3, 130])


# ----------------------------------------------------------------------------- stage 4
def test_logits_golden(golden, dev):
    from xmask3d_b200.logits import cal_pred_logits, ensemble_logits_with_labels
    g = golden("logits")
    for tag, t in (("b15", 20), ("sn200", 201)):
        o = {"mask_embed": torch.from_numpy(g[f"me_{tag}"]).to(dev), "text_embed": torch.from_numpy(g[f"te_{tag}"]).to(dev),
             "null_embed": torch.from_numpy(g[f"ne_{tag}"]).to(dev), "labels": [[str(i)] for i in range(t - 1)],
             "logit_scale": torch.tensor(1 / 0.07)}
        got, amax = cal_pred_logits(o, want_argmax=True)
        ref = g[f"logits_{tag}"]
        got = got.cpu().numpy()
        assert got.shape == ref.shape
        assert np.array_equal(got.argmax(-1), ref.argmax(-1))                      # identical argmax
        assert np.array_equal(amax.cpu().numpy(), ref.argmax(-1))
        assert np.abs(got - ref).max() < 2e-5 * max(1.0, np.abs(ref).max()), np.abs(got - ref).max()
    sizes = g["sizes_grp"].tolist()
    labels, s = [], 0
    for n in sizes:
        labels.append([str(s + j) for j in range(n)])
        s += n
    o = {"mask_embed": torch.from_numpy(g["me_grp"]).to(dev), "text_embed": torch.from_numpy(g["te_grp"]).to(dev),
         "null_embed": torch.from_numpy(g["ne_grp"]).to(dev), "labels": labels, "logit_scale": torch.tensor(100.0)}
    got = cal_pred_logits(o).cpu().numpy()
    assert np.array_equal(got.argmax(-1), g["logits_grp"].argmax(-1))
    assert np.abs(got - g["logits_grp"]).max() < 2e-5 * np.abs(g["logits_grp"]).max()
    raw = torch.from_numpy(g["ens_in"]).to(dev)
    assert np.array_equal(ensemble_logits_with_labels(raw, labels, "max").cpu().numpy(), g["ens_max"])
    np.testing.assert_allclose(ensemble_logits_with_labels(raw, labels, "mean").cpu().numpy(), g["ens_mean"], rtol=1e-5, atol=1e-7)
    with pytest.raises(AssertionError):
        ensemble_logits_with_labels(raw, labels[:-1], "max")


# ----------------------------------------------------------------------------- whole path
def test_pipeline_end_to_end_vs_oracle(cport, dev):
    """project -> voxelize -> masks at points -> pool, chained on the device with no host round
    trip, against the oracle run stage by stage; per-point features gathered through vis_idx, and
    the scene-level all-reduce helper in its single-process form."""
    from oracle import ref_port as P
    from xmask3d_b200 import dist as xd
    from xmask3d_b200.pipeline import Batch, CorrespondencePipeline
    k, c, n_views = 9, 32, 3
    sc = syn.make_scene(91, 60_000)
    views = [syn.make_view(sc, v) for v in range(n_views)]
    rng = np.random.default_rng(3)
    rts = [_random_rt(rng, 0.02) for _ in views]
    batch = Batch(sc.xyz, np.array([0, 60_000], np.int64), np.zeros(n_views, np.int64),
                  np.stack([np.linalg.inv(v.pose) for v in views]), np.stack([v.depth_mm for v in views]),
                  np.stack([r[:3, :4] for r in rts]), syn.scannet_intrinsics())
    pipe = CorrespondencePipeline(batch, k, c, dev, pairs_per_point=1.0)
    pipe.upload(torch.from_numpy(batch.xyz), torch.from_numpy(batch.depth_mm.view(np.int16)))
    masks_np = np.stack([syn.make_partition_masks(40 + v, k) for v in range(n_views)])
    feat = rng.standard_normal((60_000, c), dtype=np.float32)
    out = pipe.run(torch.from_numpy(masks_np).to(dev), torch.from_numpy(feat).to(dev), feat_per_point=True)
    assert int(out["proj"].status.item()) == 0 and int(out["vox"].status.item()) == 0
    voff = out["proj"].vis_off.cpu().numpy()
    uoff = out["vox"].uniq_off.cpu().numpy()
    m = P.getMapping()
    tot_ref, cnt_ref = np.zeros((k, c)), np.zeros(k, np.int64)
    for v, vw in enumerate(views):
        mapping = m.compute_mapping(vw.pose, sc.xyz, vw.depth_m)
        vis, xl, yl = P.compact_mapping(mapping)
        a, b = voff[v], voff[v + 1]
        assert np.array_equal(out["proj"].vis_idx.cpu().numpy()[a:b], np.nonzero(vis)[0])
        assert np.array_equal(out["proj"].rowcol.cpu().numpy()[a:b], np.stack([xl, yl], 1))
        g, first, inv = cport.voxelize(sc.xyz[vis], rts[v])
        assert np.array_equal(out["vox"].first.cpu().numpy()[uoff[v]:uoff[v + 1]], first)
        assert np.array_equal(out["vox"].inverse.cpu().numpy()[a:b] - uoff[v], inv)          # collated offsets
        assert np.array_equal(out["vox"].voxel_xyz.cpu().numpy()[uoff[v]:uoff[v + 1]].astype(np.float64), g)
        member = masks_np[v][:, xl, yl]
        s64, c64 = cport.pool_member_f64(feat[vis], member)
        assert np.array_equal(out["cnt"][v].cpu().numpy(), c64)
        assert (np.abs(out["sum"][v].cpu().numpy() - s64).max(1) / np.maximum(np.abs(s64).max(1), 1e-30)).max() < 1e-5
        tot_ref += s64
        cnt_ref += c64
    tot, cnt = xd.allreduce_mask_sums(out["sum"], out["cnt"])
    assert np.array_equal(cnt.cpu().numpy(), cnt_ref)
    mean = xd.finalize_mean(tot, cnt).cpu().numpy()
    ref_mean = tot_ref / np.maximum(cnt_ref, 1)[:, None]
    assert (np.abs(mean - ref_mean).max(1) / np.maximum(np.abs(ref_mean).max(1), 1e-30)).max() < 1e-5


# ----------------------------------------------------------------------------- full-size properties
def _u64_order(t):
    """int64 tensor holding uint64 bits -> int64 with the same ordering as the unsigned values."""
    return t ^ torch.tensor(-2 ** 63, dtype=torch.int64, device=t.device)


def test_full_size_properties(dev):
    """BASELINE configs[1] sizes (150k-point scenes x 20 views, C=768, K=50): size-independent
    invariants of every stage — no oracle, the CPU path would take minutes."""
    from xmask3d_b200 import ops
    from xmask3d_b200.pipeline import Batch, CorrespondencePipeline
    from xmask3d_b200.voxelizer import Voxelizer
    from bench import LOADER_VOX
    n_scenes, n_views, n_pts, k, c = 2, 20, 150_000, 50, 768
    xyz, off, vs, w2c, depth, rts = [], [0], [], [], [], []
    for s in range(n_scenes):
        sc = syn.make_scene(1000 + s, n_pts)
        xyz.append(sc.xyz)
        off.append(off[-1] + n_pts)
        for v in range(n_views):
            vw = syn.make_view(sc, v)
            vs.append(s)
            w2c.append(np.linalg.inv(vw.pose))
            depth.append(vw.depth_mm)
            np.random.seed(5557 + 1000 * s + v)
            rts.append(Voxelizer(voxel_size=0.02, **LOADER_VOX).draw_rigid_transformation()[0][:3, :4])
    batch = Batch(np.concatenate(xyz), np.array(off, np.int64), np.array(vs, np.int64), np.stack(w2c), np.stack(depth),
                  np.stack(rts), syn.scannet_intrinsics())
    pipe = CorrespondencePipeline(batch, k, c, dev)
    pipe.upload(torch.from_numpy(batch.xyz), torch.from_numpy(batch.depth_mm.view(np.int16)))
    pr = pipe.project()
    total = int(pr.vis_off[-1].item())
    pipe.set_cap(total)
    V = batch.n_views
    masks = torch.from_numpy(np.stack([syn.make_partition_masks(500 + v, k) for v in range(V)])).to(dev)
    feat = torch.randn(total, c, device=dev, generator=torch.Generator(device=dev).manual_seed(3))
    out = pipe.run(masks, feat)
    pr, vox = out["proj"], out["vox"]
    assert int(pr.status.item()) == 0 and int(vox.status.item()) == 0 and int(out["pool_status"].item()) == 0
    voff, nv = pr.vis_off.cpu().numpy(), pr.n_vis.cpu().numpy()
    assert voff[0] == 0 and np.all(np.diff(voff) == nv) and voff[-1] == total and nv.max() > 20_000
    seg = torch.repeat_interleave(torch.arange(V, device=dev), torch.from_numpy(nv).to(dev))
    # projection: visibility bytes agree with the compaction; pixels inside the cut image; order kept
    assert int(pr.vis.sum().item()) == total
    rc = pr.rowcol[:total]
    assert int(rc[:, 0].min()) >= 10 and int(rc[:, 0].max()) < 230 and int(rc[:, 1].min()) >= 10 and int(rc[:, 1].max()) < 310
    d = pr.vis_idx[1:total] - pr.vis_idx[:total - 1]
    same = seg[1:] == seg[:-1]
    assert bool((d[same] > 0).all())
    scene_of_view = torch.from_numpy(batch.view_scene).to(dev)
    src = pr.vis_idx[:total].long() + scene_of_view[seg] * n_pts
    assert torch.equal(pr.xyz_vis[:total], pipe.xyz[src])
    # voxelization: keys strictly ascending inside every segment; inverse / first consistent;
    # every point's own grid coordinates equal its voxel's
    m, uoff = vox.m.cpu().numpy(), vox.uniq_off.cpu().numpy()
    M = int(uoff[-1])
    assert np.all(np.diff(uoff) == m) and M == m.sum() and np.all(m <= nv) and np.all(m[nv > 0] >= 1) and M > 0.6 * total
    useg = torch.repeat_interleave(torch.arange(V, device=dev), torch.from_numpy(m).to(dev))
    keys = _u64_order(ops.fnv_hash(vox.voxel_xyz[:M].double()))
    assert bool((keys[1:] > keys[:-1])[useg[1:] == useg[:-1]].all())
    inv = vox.inverse[:total].long()                                     # collated (global unique index)
    assert bool((useg[inv] == seg).all())
    local = torch.arange(total, device=dev) - torch.from_numpy(voff).to(dev)[seg]
    first = vox.first[:M].long()
    assert bool((first[inv] <= local).all())
    assert torch.equal(inv[first + torch.from_numpy(voff).to(dev)[useg]], torch.arange(M, device=dev))
    rt = pipe.rt[seg]                                                    # [total,3,4] float64
    grid = torch.floor((rt[:, :, :3] * pr.xyz_vis[:total].double().unsqueeze(1)).sum(-1) + rt[:, :, 3])
    gmin = torch.stack([grid[voff[v]:voff[v + 1]].min(0).values for v in range(V)])
    # (float64 sum order differs from the FMA chain only in the last bit; compare where it is safe)
    frac_ok = ((rt[:, :, :3] * pr.xyz_vis[:total].double().unsqueeze(1)).sum(-1) + rt[:, :, 3])
    safe = ((frac_ok - torch.floor(frac_ok)).abs() > 1e-9).all(1) & ((torch.ceil(frac_ok) - frac_ok).abs() > 1e-9).all(1)
    assert float(safe.float().mean()) > 0.999
    assert torch.equal((grid - gmin[seg])[safe].long(), vox.voxel_xyz[:M].long()[inv][safe])
    # pooling: partition masks -> counts add up to the visible points, sums add up to the column sums
    cnt, s = out["cnt"], out["sum"]
    assert np.array_equal(cnt.sum(1).cpu().numpy(), nv)
    colsum = torch.zeros(V, c, device=dev, dtype=torch.float64).index_add_(0, seg, feat.double())
    err = (s.double().sum(1) - colsum).abs().max(1).values / colsum.abs().max(1).values
    assert float(err.max()) < 1e-5
    mean_chk = (out["mean"].double() * cnt.unsqueeze(-1)).sum(1)
    assert float(((mean_chk - colsum).abs().max(1).values / colsum.abs().max(1).values).max()) < 1e-5
    # idempotence: a second pass returns identical bits (deterministic reductions)
    out2 = pipe.run(masks, feat)
    assert torch.equal(out2["sum"], s) and torch.equal(out2["vox"].inverse[:total], vox.inverse[:total])


def test_logits_full_size_vs_torch(dev):
    """configs[2]: 160 x 100 mask embeddings vs 200 + null text embeddings (ScanNet200), against a
    float32 torch matmul on the same device: identical argmax, |diff| at float32 rounding level."""
    import torch.nn.functional as F
    from xmask3d_b200 import ops
    g = torch.Generator(device=dev).manual_seed(5)
    for rows, t in ((160 * 100, 201), (160 * 50, 20), (129, 33), (1, 2)):
        me = torch.randn(rows, 768, device=dev, generator=g)
        te = torch.randn(t - 1, 768, device=dev, generator=g)
        ne = torch.randn(1, 768, device=dev, generator=g)
        got, amax = ops.logits(me, te, ne, [1] * (t - 1), 1 / 0.07, want_argmax=True)
        torch.backends.cuda.matmul.allow_tf32 = False
        mn = F.normalize(me.double(), dim=-1)
        ref = (1 / 0.07) * torch.cat([mn @ F.normalize(te.double(), dim=-1).t(), mn @ F.normalize(ne.double(), dim=-1).t()], -1)
        assert got.shape == ref.shape
        assert float((got.double() - ref).abs().max()) < 2e-5 * float(ref.abs().max())
        top2 = ref.topk(2, -1).values
        clear = (top2[:, 0] - top2[:, 1]) > 1e-4                     # ignore numerical ties of the reference itself
        assert torch.equal(got.argmax(-1)[clear], ref.argmax(-1)[clear])
        assert torch.equal(amax.long()[clear], ref.argmax(-1)[clear])
    # embedding width that is not a multiple of the 32-float k-block, and mean ensembling
    me = torch.randn(70, 100, device=dev, generator=g)
    te = torch.randn(9, 100, device=dev, generator=g)
    ne = torch.randn(1, 100, device=dev, generator=g)
    got = ops.logits(me, te, ne, [2, 3, 4], 10.0, ensemble="mean")
    mn = F.normalize(me.double(), dim=-1)
    raw = 10.0 * (mn @ F.normalize(te.double(), dim=-1).t())
    ref = torch.stack([raw[:, :2].mean(1), raw[:, 2:5].mean(1), raw[:, 5:].mean(1),
                       10.0 * (mn @ F.normalize(ne.double(), dim=-1).t())[:, 0]], 1)
    assert float((got.double() - ref).abs().max()) < 1e-4


# ----------------------------------------------------------------------------- error behaviour
def test_abi_error_reporting(dev):
    """Bad arguments come back as negative status codes with a message, never as a crash; capacity
    overflows are flagged in device memory and nothing is overrun."""
    import ctypes as C
    from xmask3d_b200 import _lib as L, ops
    lib = L.lib()
    sm, major = C.c_int32(0), C.c_int32(0)
    assert lib.xm3d_device_info(C.byref(sm), C.byref(major), None) == 0 and sm.value >= 100 and major.value == 10
    assert lib.xm3d_pool_batch(None, 768, None, None, None, 1, 50, None, 0, 0, 0, None, None, None, None, 0, None, None) == -1
    assert b"xm3d_pool_batch" in lib.xm3d_last_error()
    with pytest.raises(L.Xm3dError):                      # workspace too small
        feat = torch.zeros(8, 4, device=dev)
        seg = torch.tensor([0, 8], dtype=torch.int64, device=dev)
        lab = torch.zeros(8, dtype=torch.int32, device=dev)
        out = torch.zeros(1, 2, 4, device=dev)
        L.check(lib.xm3d_pool_batch(C.c_void_p(feat.data_ptr()), 4, None, None, C.c_void_p(lab.data_ptr()), 1, 2,
                                    C.c_void_p(seg.data_ptr()), 8, 8, 0, C.c_void_p(out.data_ptr()), None, None,
                                    C.c_void_p(out.data_ptr()), 16, None, None))
    with pytest.raises(L.Xm3dError):                      # more than 256 masks per segment
        ops.pool(torch.zeros(8, 4, device=dev), torch.tensor([0, 8], device=dev), 300,
                 label=torch.zeros(8, dtype=torch.int32, device=dev))
    with pytest.raises(L.Xm3dError):                      # empty label group
        lib_rc = ops.logits(torch.randn(4, 8, device=dev), torch.randn(3, 8, device=dev), torch.randn(1, 8, device=dev),
                            [3, 0], 1.0) if False else None
        L.check(lib.xm3d_logits(None, 4, 8, None, 3, None, None, 2, 0, 1.0, None, None, None, 0, None))
    # pooling with more memberships than cap_pairs: flagged, sums zero, no overrun
    n, k, c = 5000, 40, 64
    member = torch.full((n, 2), -1, dtype=torch.int32, device=dev)            # every point in every mask
    status = torch.zeros(1, dtype=torch.int32, device=dev)
    s, cnt, _ = ops.pool(torch.ones(n, c, device=dev), torch.tensor([0, n], device=dev), k, member=member,
                         cap_pairs=1000, status=status)
    assert int(status.item()) & L.FLAG_PAIR_OVERFLOW and float(s.abs().sum()) == 0.0
    s, cnt, _ = ops.pool(torch.ones(n, c, device=dev), torch.tensor([0, n], device=dev), k, member=member)
    assert torch.all(cnt == n) and torch.all(s == n)
    # voxelize with more points than `cap`: flagged, nothing processed
    xyz = torch.rand(1000, 3, device=dev)
    rt = torch.eye(4, dtype=torch.float64, device=dev)[:3].reshape(1, 3, 4) * 50
    u = ops.voxelize_batch(xyz, torch.tensor([0, 1000], device=dev), rt, cap=500)
    assert int(u.status.item()) & L.FLAG_VIS_OVERFLOW and int(u.m.sum().item()) == 0


def test_pack_i16(dev):
    """Host-bound int16 copies of x/y labels and voxel coordinates: exact below 32768, device row count honoured,
    out-of-range values clamped and flagged."""
    from xmask3d_b200 import _lib as L, ops
    g = torch.Generator(device=dev).manual_seed(1)
    src = torch.randint(-3000, 32767, (100_003, 3), device=dev, generator=g, dtype=torch.int32)
    rows = torch.tensor([77_777], dtype=torch.int64, device=dev)
    out = torch.full((100_003, 3), -7, dtype=torch.int16, device=dev)
    st = torch.zeros(1, dtype=torch.int32, device=dev)
    ops.pack_i16(src, rows, out=out, status=st)
    assert torch.equal(out[:77_777].to(torch.int32), src[:77_777]) and bool((out[77_777:] == -7).all()) and int(st.item()) == 0
    src[5, 1] = 40_000
    src[9, 0] = -50_000
    o2 = ops.pack_i16(src, status=st)
    assert int(st.item()) == L.FLAG_I16_RANGE and int(o2[5, 1]) == 32767 and int(o2[9, 0]) == -32768


def test_vote_accumulation(cport, dev):
    """Cross-view votes (run/infer.py:642-647, :658) on the projection's compaction outputs."""
    from oracle import ref_port as P
    from xmask3d_b200 import ops
    t, n_views = 19, 5
    scenes = [syn.make_scene(61, 30_000), syn.make_scene(62, 17_001)]
    intr = syn.scannet_intrinsics()
    xyz = np.concatenate([s.xyz for s in scenes])
    pt_off, n_pts, w2c, depth, view_scene = [], [], [], [], []
    o = 0
    for si, sc in enumerate(scenes):
        for v in range(n_views):
            vw = syn.make_view(sc, v)
            pt_off.append(o); n_pts.append(sc.xyz.shape[0]); w2c.append(np.linalg.inv(vw.pose)); depth.append(vw.depth_mm)
            view_scene.append(si)
        o += sc.xyz.shape[0]
    views, out_off = ops.make_views(np.stack(w2c), intr, pt_off, n_pts, depth[0].shape)
    pr = ops.project_batch(torch.from_numpy(xyz).to(dev), views, out_off,
                           torch.from_numpy(np.stack(depth).view(np.int16)).to(dev), want_mapping=True)
    total = int(pr.vis_off[-1].item())
    rng = np.random.default_rng(8)
    cls = rng.integers(0, t, total).astype(np.int32)
    votes = torch.zeros(xyz.shape[0], t, dtype=torch.int32, device=dev)
    counter = torch.zeros(xyz.shape[0], dtype=torch.int32, device=dev)
    ops.accumulate_votes(pr.vis_idx, pr.vis_off, torch.tensor(pt_off, dtype=torch.int64, device=dev),
                         torch.from_numpy(cls).to(dev), votes, counter, cap=total)
    pred = ops.vote_argmax(votes, counter).cpu().numpy()
    ref_votes = np.zeros((xyz.shape[0], t), np.int64)
    ref_cnt = np.zeros(xyz.shape[0], np.int64)
    mapping, voff = pr.mapping.cpu().numpy(), pr.vis_off.cpu().numpy()
    for v in range(len(views)):
        m = mapping[out_off[v]:out_off[v + 1]][:, 2] == 1
        sl = slice(pt_off[v], pt_off[v] + n_pts[v])
        P.accumulate_votes(ref_votes[sl], ref_cnt[sl], m, cls[voff[v]:voff[v + 1]])
    assert np.array_equal(votes.cpu().numpy(), ref_votes) and np.array_equal(counter.cpu().numpy(), ref_cnt)
    seen = ref_cnt > 0
    assert np.array_equal(pred[seen], P.vote_argmax(ref_votes)[seen]) and np.all(pred[~seen] == -1)
    assert seen.sum() > 10_000 and ref_cnt.max() >= 3


# ----------------------------------------------------------------------------- training drop-in
def test_autograd_wrappers(dev):
    """Gradients of cal_pred_logits and mask_mapper's scatter-mean against plain torch autograd of
    the reference formulas (float32; tolerance 1e-4 relative to the gradient's largest entry)."""
    import torch.nn.functional as F
    from xmask3d_b200.fuser import mask_mapper
    from xmask3d_b200.logits import cal_pred_logits
    g = torch.Generator(device=dev).manual_seed(2)
    me = torch.randn(2, 11, 64, device=dev, generator=g, requires_grad=True)
    te = torch.randn(7, 64, device=dev, generator=g, requires_grad=True)
    ne = torch.randn(1, 64, device=dev, generator=g, requires_grad=True)
    ls = torch.tensor(14.2857, device=dev, requires_grad=True)
    labels = [["a"], ["b", "c"], ["d"], ["e", "f", "g"]]
    w = torch.randn(2, 11, 5, device=dev, generator=g)
    out = cal_pred_logits({"mask_embed": me, "text_embed": te, "null_embed": ne, "labels": labels, "logit_scale": ls})
    (out * w).sum().backward()
    got = [t.grad.clone() for t in (me, te, ne, ls)]
    for t in (me, te, ne, ls):
        t.grad = None
    m = F.normalize(me, dim=-1)
    pred = ls * (m @ F.normalize(te, dim=-1).t())
    ref = torch.cat([torch.stack([p.max(-1).values for p in torch.split(pred, [1, 2, 1, 3], -1)], -1),
                     ls * (m @ F.normalize(ne, dim=-1).t())], -1)
    assert float((out - ref).abs().max()) < 1e-4
    (ref * w).sum().backward()
    for a, t in zip(got, (me, te, ne, ls)):
        assert float((a - t.grad).abs().max()) <= 1e-4 * max(1.0, float(t.grad.abs().max()))
    # scatter-mean inside mask_mapper
    k, h, wd, n, c = 6, 24, 32, 500, 16
    rngs = np.random.default_rng(4)
    xl, yl = torch.from_numpy(rngs.integers(0, h, n)), torch.from_numpy(rngs.integers(0, wd, n))
    masks = (torch.rand(k, h, wd, device=dev, generator=g) > 0.7).float()
    emb = torch.randn(k, c, device=dev, generator=g, requires_grad=True)
    pred3d = torch.randn(n, c, device=dev, generator=g)

    class Cfg:
        caption_contra_2d_pre = False
    wt = torch.randn(n, c, device=dev, generator=g)
    fused, f2d, _, _ = mask_mapper([xl], [yl], [masks], [emb], [pred3d], lambda a, b: a * 2 + b, lambda a: a, lambda a: a, Cfg)
    ((fused[0] + f2d[0]) * wt).sum().backward()
    got = emb.grad.clone()
    emb.grad = None
    member = masks[:, xl.to(dev), yl.to(dev)] >= 0.5                           # reference loop, differentiable
    feat = torch.zeros(n, c, device=dev)
    cnt = torch.zeros(n, 1, device=dev)
    for mk, e in zip(member, emb):
        if mk.sum() == 0:
            continue
        feat = feat + mk.unsqueeze(1).float() * e
        cnt = cnt + mk.unsqueeze(1).float()
    cnt = torch.where(cnt == 0, torch.full_like(cnt, 1e-5), cnt)
    feat = feat / cnt
    cov = cnt.squeeze(1) >= 1
    final = torch.where(cov.unsqueeze(1), feat * 2 + pred3d, pred3d)
    ((final + feat) * wt).sum().backward()
    assert float((got - emb.grad).abs().max()) <= 1e-4 * float(emb.grad.abs().max())


def test_point_logits_vs_torch(dev):
    """Per-point logits / base-novel blend / argmax (run/infer.py:557, 606-640) against float64 torch."""
    import torch.nn.functional as F
    from xmask3d_b200 import ops
    g = torch.Generator(device=dev).manual_seed(9)
    for n, t, c in ((40_000, 19, 768), (1000, 200, 768), (129, 7, 100), (1, 3, 64)):
        feat = torch.randn(n, c, device=dev, generator=g) * 3
        te = torch.randn(t, c, device=dev, generator=g)
        ref = (1 / 0.07) * (F.normalize(feat.double(), dim=-1) @ F.normalize(te.double(), dim=-1).t())
        out, amax = ops.point_logits(feat, te, 1 / 0.07)
        assert out.shape == (n, t)
        assert float((out.double() - ref).abs().max()) < 2e-5 * float(ref.abs().max())
        top2 = ref.topk(min(2, t), -1).values
        clear = (top2[:, 0] - top2[:, -1]) > 1e-4 if t > 1 else torch.ones(n, dtype=torch.bool, device=dev)
        assert torch.equal(amax.long()[clear], ref.argmax(-1)[clear])
        # base / novel blending with the binary head
        binary = (torch.rand(n, device=dev, generator=g) > 0.5).float()
        is_base = torch.rand(t, device=dev, generator=g) > 0.4
        out_b, amax_b = ops.point_logits(feat, te, 1 / 0.07, binary=binary, is_base=is_base)
        novel = ref.clone(); novel[:, is_base] = -1e10
        base = ref.clone(); base[:, ~is_base] = -1e10
        ref_b = binary.double().unsqueeze(1) * base + (1 - binary.double().unsqueeze(1)) * novel
        live = ref_b > -1e9
        assert float((out_b.double() - ref_b)[live].abs().max() if live.any() else 0.0) < 2e-5 * float(ref.abs().max())
        assert bool(((out_b < -1e9) == ~live).all())
        ok = live.any(1)
        rb = ref_b.clone()
        t2 = rb.topk(min(2, t), -1).values
        clear_b = ok & (((t2[:, 0] - t2[:, -1]) > 1e-4) if t > 1 else ok)
        assert torch.equal(amax_b.long()[clear_b], rb.argmax(-1)[clear_b])
        # argmax-only call (no [n,T] write)
        none, amax2 = ops.point_logits(feat, te, 1 / 0.07, want_logits=False)
        assert none is None and torch.equal(amax2, amax)


@pytest.mark.parametrize("n,t,kf,c", [(30_000, 19, 23, 768), (5000, 200, 60, 768), (257, 7, 1, 128)])
def test_point_logits_fused_stream_ensemble(dev, n, t, kf, c):
    """The FUSED stream of the inference loop (run/infer.py:516-608): softmax over the classes, geometric-mean
    ensemble with the per-mask MaskCLIP probabilities in the log domain (base / novel ratios through
    category_overlapping_mask), -1e10 masking, blend with the binary head, argmax — in the epilogue of the
    tcgen05 kernel, against the statement-for-statement torch-CPU restatement (oracle/ref_port.py).
    Bars: identical argmax wherever the reference's own top-2 gap exceeds 1e-4; values within 1e-4 absolute
    (log-domain values of size ~10; the float32 pow / log / exp of the two implementations differ in the last bits)."""
    from oracle import ref_port
    from xmask3d_b200 import ops
    g = torch.Generator().manual_seed(n + t)
    feat = torch.randn(n, c, generator=g) * 2
    te = torch.randn(t, c, generator=g)
    open_emb = torch.randn(kf, c, generator=g)
    label = torch.randint(-1, kf, (n,), generator=g)                       # -1: the point is in no final mask
    final_mask_3d = label.unsqueeze(0) == torch.arange(kf).view(kf, 1)     # partition
    base_cat = sorted(torch.randperm(t, generator=g)[: max(1, (3 * t) // 4)].tolist())
    novel_cat = [i for i in range(t) if i not in base_cat]
    binary = (torch.rand(n, 1, generator=g) > 0.5).float()
    scale, rb, rn = 1 / 0.07, 0.65, 0.35
    ref_val, ref_arg = ref_port.fused_stream_ref(feat, te, scale, open_emb, final_mask_3d, base_cat, novel_cat, t, rb, rn,
                                                 binary)
    import torch.nn.functional as F
    probs = (scale * (F.normalize(open_emb, dim=-1) @ F.normalize(te, dim=-1).t())).softmax(-1)
    is_base = torch.zeros(t, dtype=torch.bool)
    is_base[base_cat] = True
    out, amax = ops.point_logits(feat.to(dev), te.to(dev), scale, binary=binary.to(dev), is_base=is_base.to(dev),
                                 mask_label=label.to(torch.int32).to(dev), mask_probs=probs.to(dev), base_ratio=rb,
                                 novel_ratio=rn)
    out, amax = out.cpu(), amax.cpu().long()
    live = ref_val > -1e9
    assert bool(((out < -1e9) == ~live).all())
    finite = live & torch.isfinite(ref_val)
    assert float((out - ref_val)[finite].abs().max()) < 1e-4
    top2 = torch.where(torch.isfinite(ref_val), ref_val, torch.full_like(ref_val, -1e10)).topk(min(2, t), -1).values
    clear = torch.isfinite(ref_val).all(1) & ((top2[:, 0] - top2[:, -1]) > 1e-4 if t > 1 else torch.ones(n, dtype=torch.bool))
    assert float(clear.float().mean()) > 0.98
    assert torch.equal(amax[clear], ref_arg[clear])
    # argmax-only call gives the same classes
    _, amax2 = ops.point_logits(feat.to(dev), te.to(dev), scale, binary=binary.to(dev), is_base=is_base.to(dev),
                                mask_label=label.to(torch.int32).to(dev), mask_probs=probs.to(dev), base_ratio=rb,
                                novel_ratio=rn, want_logits=False)
    assert torch.equal(amax2.cpu().long(), amax)


def test_nn_fill_and_segment_max(golden, dev):
    """Nearest-seen-neighbour fill (run/infer.py:651-656, 684-694) against the golden KDTree result and the
    exhaustive oracle on a ragged multi-scene batch (a scene without seen points, one without unseen
    points, duplicated points, an outlier far outside the seen bounding box); per-scene maxima."""
    from oracle import ref_port
    from xmask3d_b200 import ops
    g = golden("nnfill")
    xyz = torch.from_numpy(g["xyz"]).to(dev)
    counter = torch.from_numpy(g["counter"]).to(dev)
    match = ops.nn_fill_match(xyz, counter).cpu().numpy()
    false_idx = g["false_idx"]
    same = match[false_idx] == g["match"]
    if not same.all():
        x64 = g["xyz"].astype(np.float64)
        bad = false_idx[~same]
        assert np.array_equal(((x64[bad] - x64[match[bad]]) ** 2).sum(1), ((x64[bad] - x64[g["match"][~same]]) ** 2).sum(1))
    filled = ops.nn_fill(torch.from_numpy(g["pred"]).to(dev), xyz, counter).cpu().numpy()
    assert np.array_equal(filled, g["filled"])
    # ragged batch against the exhaustive restatement
    rng = np.random.default_rng(5)
    sc = syn.make_scene(43, 6000)
    a = sc.xyz[:2500].copy()
    a[7] = [40.0, -35.0, 20.0]                                   # unseen outlier
    ca = (rng.uniform(size=2500) < 0.4).astype(np.int32); ca[7] = 0
    b = sc.xyz[2500:3000]; cb = np.zeros(500, np.int32)          # nothing seen
    c = sc.xyz[3000:3700]; cc = np.ones(700, np.int32)           # everything seen
    d = np.repeat(sc.xyz[3700:3900], 3, 0); cd = np.tile(np.array([1, 0, 0], np.int32), 200)   # exact duplicates
    e = sc.xyz[4000:4001]; ce = np.zeros(1, np.int32)
    f = sc.xyz[4100:6000]; cf = (rng.uniform(size=1900) < 0.02).astype(np.int32); cf[0] = 1  # sparse seen set
    parts, cnts = [a, b, c, d, e, f], [ca, cb, cc, cd, ce, cf]
    off = np.concatenate([[0], np.cumsum([len(p) for p in parts])]).astype(np.int64)
    xyz_b = torch.from_numpy(np.concatenate(parts)).to(dev)
    cnt_b = torch.from_numpy(np.concatenate(cnts)).to(dev)
    got = ops.nn_fill_match(xyz_b, cnt_b, torch.from_numpy(off).to(dev)).cpu().numpy()
    for i, (p, c_) in enumerate(zip(parts, cnts)):
        ref = ref_port.nn_fill_match(p, c_)
        mine = got[off[i]:off[i + 1]] - off[i]
        if not (c_ != 0).any():
            assert np.array_equal(mine, np.arange(len(p))), f"scene {i}: nothing seen -> unchanged"
            continue
        assert np.array_equal(mine, ref), f"scene {i}"
    # per-scene maximum
    idx = g["idx"]
    soff = np.concatenate([[0], np.cumsum(np.bincount(idx))]).astype(np.int64)
    sm = ops.segment_max(torch.from_numpy(g["feat"]).to(dev), torch.from_numpy(soff).to(dev)).cpu().numpy()
    assert np.array_equal(sm, g["segmax"])


def test_voxelize_with_clip_bound_golden(golden, dev):
    """voxelize() with a clip_bound (dataset/voxelizer.py:87-102): the translation-augmentation draws come
    before the matrix draws, the cloud is cropped, then voxelized — same np.random state, same outputs."""
    from tests.golden.make_golden_params import vox_kwargs
    from xmask3d_b200.voxelizer import Voxelizer
    g = golden("clip")
    kw = vox_kwargs(0.05)
    kw["clip_bound"] = tuple(map(tuple, g["clip"]))
    np.random.seed(991)
    grid, feats, lab, inv, inds = Voxelizer(**kw).voxelize(g["xyz"], g["colors"].copy(), g["labels"].copy(),
                                                           return_ind=True)
    assert np.array_equal(grid, g["grid"]) and np.array_equal(inv, g["inv"]) and np.array_equal(inds, g["inds"])
    assert np.array_equal(feats, g["feats"]) and np.array_equal(lab, g["lab"])
