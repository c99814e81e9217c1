"""Bit-exact parity AT BENCH SCALE.

The projection decides most points with a float32 error-bound classifier (csrc/project.cu) and the
voxelization runs in shared-memory key-range units (csrc/voxelize.cu).  Both are proven here against
the scalar C oracle (oracle/xm3d_oracle.c, the restatement of models/utils/fusion_util.py:46-142 and
dataset/voxelizer.py:81-132 pinned to the reference by tests/golden + tests/test_reference_live.py):

  * BASELINE configs[1] in full — 8 scenes x 20 views x 150k points = 24 M point*views: the int64
    [N,3] mapping, visibility, compaction order, x/y labels of EVERY view, and first / inverse / voxel
    coordinates of EVERY view's visible points;
  * one 1 M-point scene at 1 cm voxels (configs[3] geometry), 3 views, segments of > 224 k points;
  * > 1.2e7 adversarial point*views built to sit on the classifier's decision boundaries: pixel
    half-integers, the cut border, |d - z| = vis_thres * d, z ~ 1e-8, and depth = 0 pixels.

Needs a B200 (`pytest -m gpu`); the oracle side takes ~20 s of host time.
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


def _check_projection(cport, pr, out_off, views_np, intr, cut=10, thres=0.25, want_compact=True):
    """views_np: list of (xyz float32 [N,3], w2c, depth_u16 or None).  Returns visible index lists."""
    mapping = pr.mapping.cpu().numpy()
    vis = pr.vis.cpu().numpy()
    n_vis = pr.n_vis.cpu().numpy()
    vis_off = pr.vis_off.cpu().numpy()
    if want_compact:
        vis_idx, rowcol, xyz_vis = pr.vis_idx.cpu().numpy(), pr.rowcol.cpu().numpy(), pr.xyz_vis.cpu().numpy()
    n_checked, n_visible = 0, 0
    for i, (xyz, w2c, depth) in enumerate(views_np):
        with np.errstate(all="ignore"):
            ref = cport.project(xyz, w2c, intr, depth, cut=cut, vis_thres=thres)
        a, b = int(out_off[i]), int(out_off[i + 1])
        if not np.array_equal(mapping[a:b], ref):
            bad = np.nonzero((mapping[a:b] != ref).any(1))[0]
            raise AssertionError(f"view {i}: {len(bad)} mapping rows differ, first {bad[:5]}: "
                                 f"got {mapping[a:b][bad[:5]]}, oracle {ref[bad[:5]]}, xyz {xyz[bad[:5]]}")
        assert np.array_equal(vis[a:b], ref[:, 2].astype(np.uint8)), f"view {i}"
        idx = np.nonzero(ref[:, 2])[0]
        assert n_vis[i] == len(idx) and vis_off[i + 1] - vis_off[i] == len(idx), f"view {i}"
        if want_compact:
            s = slice(int(vis_off[i]), int(vis_off[i + 1]))
            assert np.array_equal(vis_idx[s], idx), f"view {i}"
            assert np.array_equal(rowcol[s], ref[idx][:, :2]), f"view {i}"
            assert np.array_equal(xyz_vis[s], xyz[idx]), f"view {i}"
        n_checked += b - a
        n_visible += len(idx)
    return n_checked, n_visible


def _check_voxels(cport, u, seg_off, xyz_vis, rts, collate):
    m, uoff = u.m.cpu().numpy(), u.uniq_off.cpu().numpy()
    first, inv, vox = u.first.cpu().numpy(), u.inverse.cpu().numpy(), u.voxel_xyz.cpu().numpy()
    tot = 0
    for i in range(len(seg_off) - 1):
        a, b = int(seg_off[i]), int(seg_off[i + 1])
        if a == b:
            assert m[i] == 0
            continue
        rt = np.eye(4)
        rt[:3, :4] = rts[i]
        rgrid, rfirst, rinv = cport.voxelize(xyz_vis[a:b], rt)
        assert m[i] == len(rfirst), f"segment {i}: M {m[i]} vs {len(rfirst)}"
        ua = int(uoff[i])
        assert np.array_equal(first[ua:ua + m[i]], rfirst), f"segment {i}: first"
        assert np.array_equal(vox[ua:ua + m[i]].astype(np.float64), rgrid), f"segment {i}: voxel coordinates"
        assert np.array_equal(inv[a:b] - (ua if collate else 0), rinv), f"segment {i}: inverse"
        tot += int(m[i])
    assert uoff[-1] == m.sum()
    return tot


def test_configs1_full_batch_bit_exact(cport, dev):
    """configs[1] in full: what bench.py times is what is compared."""
    import argparse
    from xmask3d_b200 import ops
    from xmask3d_b200.pipeline import CorrespondencePipeline
    import bench
    args = argparse.Namespace(scenes=8, views=20, points=150_000, voxel=0.02, replicas=False)
    batch, scenes = bench.build_batch(args, 0)
    assert batch.point_views == 24_000_000
    pipe = CorrespondencePipeline(batch, 50, 768, dev)
    pipe.upload(torch.from_numpy(batch.xyz), torch.from_numpy(batch.depth_mm.view(np.int16)))
    # the projection of the pipeline (same call, same workspaces) + the int64 [N,3] drop-in output
    pr = ops.project_batch(pipe.xyz, pipe.views, pipe.out_off, pipe.depth, want_mapping=True, ws=pipe.ws_proj,
                           views_dev=pipe.views_dev)
    assert int(pr.status.item()) == 0
    views_np = [(scenes[int(s)]["xyz"], batch.w2c[v], batch.depth_mm[v]) for v, s in enumerate(batch.view_scene)]
    n_checked, n_visible = _check_projection(cport, pr, pipe.out_off, views_np, batch.intr)
    assert n_checked == 24_000_000 and n_visible > 2_000_000
    del pr.mapping
    # the voxelization exactly as the pipeline runs it (fast shared-memory units, collated inverse)
    pipe.set_cap(n_visible)
    pr2 = pipe.project()
    assert torch.equal(pr2.vis_idx[:n_visible], pr.vis_idx[:n_visible])
    u = ops.voxelize_batch(pr2.xyz_vis, pr2.vis_off, pipe.rt, cap=n_visible, collate=True, ws=pipe.ws_vox)
    assert int(u.status.item()) == 0
    assert ops.voxel_path_info(u) == (0, 0), "the bench configuration must run the shared-memory unit path"
    tot = _check_voxels(cport, u, pr2.vis_off.cpu().numpy(), pr2.xyz_vis.cpu().numpy(), batch.rt, collate=True)
    assert tot > 0.6 * n_visible


def test_configs3_dense_scene_bit_exact(cport, dev):
    """1 M points at 1 cm voxels (configs[3] geometry): segments beyond the fast path's 224 k-point limit
    and > 65536 voxels per segment, projection of 1 M-point views."""
    from xmask3d_b200 import ops
    from xmask3d_b200.voxelizer import Voxelizer
    from bench import LOADER_VOX
    n_pts = 1_000_000
    sc = syn.make_scene(2000, n_pts, room=(12.0, 10.0, 3.0))
    intr = syn.scannet_intrinsics()
    vws = [syn.make_view(sc, v) for v in range(3)]
    w2c = np.stack([np.linalg.inv(v.pose) for v in vws])
    views, out_off = ops.make_views(w2c, intr, [0] * 3, [n_pts] * 3, (240, 320))
    depth = torch.from_numpy(np.stack([v.depth_mm for v in vws]).view(np.int16)).to(dev)
    xyz = torch.from_numpy(sc.xyz).to(dev)
    pr = ops.project_batch(xyz, views, out_off, depth, want_mapping=True)
    assert int(pr.status.item()) == 0
    n_checked, n_visible = _check_projection(cport, pr, out_off, [(sc.xyz, w2c[i], vws[i].depth_mm) for i in range(3)], intr)
    assert n_checked == 3 * n_pts
    rts = []
    for v in range(3):
        np.random.seed(5557 + v)
        rts.append(Voxelizer(voxel_size=0.01, **LOADER_VOX).draw_rigid_transformation()[0][:3, :4])
    # (a) the views' visible points, (b) the WHOLE scene as one segment of 1 M points
    seg_off = pr.vis_off.cpu().numpy()
    u = ops.voxelize_batch(pr.xyz_vis, pr.vis_off, torch.from_numpy(np.stack(rts)).to(dev), cap=pr.xyz_vis.shape[0])
    assert int(u.status.item()) == 0
    _check_voxels(cport, u, seg_off, pr.xyz_vis.cpu().numpy(), rts, collate=False)
    whole_off = np.array([0, n_pts], np.int64)
    for collate in (False, True):
        u = ops.voxelize_batch(xyz, torch.from_numpy(whole_off).to(dev), torch.from_numpy(rts[0][None]).to(dev), collate=collate)
        assert int(u.status.item()) == 0
        m = _check_voxels(cport, u, whole_off, sc.xyz, rts[:1], collate=collate)
        assert m > 200_000


def _adversarial_view(rng, pose, intr, n, depth_mm, kind):
    """World-space float32 points whose exact projection lies on (or within float32 rounding of) a
    decision boundary of compute_mapping under `pose` (camera-to-world)."""
    fx, fy, cx, cy = intr
    h, w = depth_mm.shape
    if kind == "half_pixel":            # px, py within ~1e-6..1e-4 of k + 0.5 (rint boundary; ties to even)
        u = rng.integers(-2, w + 2, n) + 0.5
        v = rng.integers(-2, h + 2, n) + 0.5
        z = rng.choice([0.5, 1.0, 2.0, 4.0, 1.37, 2.9], n)
        u = u + rng.choice([0.0, 1e-7, -1e-7, 1e-6, -1e-6, 3e-5, -3e-5], n)
    elif kind == "cut_border":          # pixels right at cut_bound / dim - cut_bound, +- half a pixel
        u = rng.choice([9.5, 10.5, 309.5, 310.5, 9.0, 10.0, 309.0, 310.0], n) + rng.normal(0, 2e-6, n)
        v = rng.choice([9.5, 10.5, 229.5, 230.5, 9.0, 10.0, 229.0, 230.0, 100.25], n) + rng.normal(0, 2e-6, n)
        z = rng.uniform(0.3, 6.0, n)
    elif kind == "occlusion_edge":      # |d - z| == vis_thres * d up to rounding: z = d * (1 -+ 0.25)
        c = rng.integers(10, w - 10, n)
        r = rng.integers(10, h - 10, n)
        d = depth_mm[r, c] / 1000
        z = d * rng.choice([0.75, 1.25], n) * (1.0 + rng.choice([0.0, 1e-8, -1e-8, 2e-7, -2e-7, 1e-5, -1e-5], n))
        z = np.where(d > 0, z, rng.uniform(0.5, 3.0, n))
        u, v = c + rng.uniform(-0.49, 0.49, n), r + rng.uniform(-0.49, 0.49, n)
    elif kind == "z_zero":              # z ~ +-1e-8 .. 1e-4 around the camera plane
        z = rng.choice([1e-8, -1e-8, 3e-8, 1e-7, -1e-7, 1e-6, 1e-5, 1e-4, 0.0], n) * rng.uniform(0.5, 2.0, n)
        u, v = rng.uniform(0, w, n), rng.uniform(0, h, n)
    else:
        raise ValueError(kind)
    xc = (u - cx) * z / fx
    yc = (v - cy) * z / fy
    if kind == "z_zero":                # keep x, y of ordinary size as well (huge |px|) for half of them
        big = rng.random(n) < 0.5
        xc = np.where(big, rng.uniform(-2, 2, n), xc)
        yc = np.where(big, rng.uniform(-2, 2, n), yc)
    cam = np.stack([xc, yc, z, np.ones(n)], 0)
    return (pose @ cam)[:3].T.astype(np.float32)


def test_projection_adversarial_boundaries(cport, dev):
    """>= 1.2e7 point*views on the float32 classifier's decision boundaries (csrc/project.cu): every
    one must take the exact float64 path or be proven by the bound — bit-exact with the oracle."""
    from xmask3d_b200 import ops
    rng = np.random.default_rng(2024)
    intr = syn.scannet_intrinsics()
    sc = syn.make_scene(1003, 20_000)
    poses = [np.eye(4)]
    for v in range(5):
        poses.append(syn.make_view(sc, v).pose)
    per = 520_000
    xyz_all, views_np, pt_off, n_pts, w2c_all, depth_all = [], [], [], [], [], []
    o = 0
    for pi, pose in enumerate(poses):
        if pi == 0:
            depth = rng.integers(300, 5000, (240, 320)).astype(np.uint16)
            depth[rng.random((240, 320)) < 0.05] = 0
        else:
            depth = syn.raycast_depth_mm(sc, pose)
        pts = np.concatenate([_adversarial_view(rng, pose, intr, per, depth, kind)
                              for kind in ("half_pixel", "cut_border", "occlusion_edge", "z_zero")])
        w2c = np.linalg.inv(pose)
        xyz_all.append(pts)
        views_np.append((pts, w2c, depth))
        pt_off.append(o)
        n_pts.append(len(pts))
        w2c_all.append(w2c)
        depth_all.append(depth)
        o += len(pts)
    views, out_off = ops.make_views(np.stack(w2c_all), intr, pt_off, n_pts, (240, 320))
    dten = torch.from_numpy(np.stack(depth_all).view(np.int16)).to(dev)
    pr = ops.project_batch(torch.from_numpy(np.concatenate(xyz_all)).to(dev), views, out_off, dten, want_mapping=True)
    assert int(pr.status.item()) == 0
    n_checked, n_visible = _check_projection(cport, pr, out_off, views_np, intr)
    assert n_checked >= 12_000_000
    assert n_visible > 500_000, n_visible          # the boundary points are not all trivially rejected
    # the same points without a depth image (inside-mask only) and with float64 metre depth
    pr = ops.project_batch(torch.from_numpy(np.concatenate(xyz_all)).to(dev), views, out_off, None, want_mapping=True)
    _check_projection(cport, pr, out_off, [(p, w, None) for p, w, _ in views_np], intr)
    d64 = torch.from_numpy(np.stack(depth_all) / 1000).to(dev)
    pr = ops.project_batch(torch.from_numpy(np.concatenate(xyz_all)).to(dev), views, out_off, d64, want_mapping=True)
    _check_projection(cport, pr, out_off, [(p, w, d / 1000) for p, w, d in views_np], intr)
