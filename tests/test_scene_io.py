"""Formats feeding the path and the device-side batch layout."""
import os

import numpy as np
import pytest
import torch

from oracle import ref_port
from xmask3d_b200 import scene_io, synthetic as syn


def _write_scene(tmp, seed, n_points, n_views):
    sc = syn.make_scene(seed, n_points)
    views = [syn.make_view(sc, v) for v in range(n_views)]
    sp = os.path.join(tmp, f"scene{seed}.pth")
    scene_io.write_scene_pth(sp, sc.xyz, sc.colors, sc.labels)
    pp, dp = [], []
    for v, vw in enumerate(views):
        pp.append(os.path.join(tmp, f"scene{seed}_{v}.txt"))
        dp.append(os.path.join(tmp, f"scene{seed}_{v}.png"))
        scene_io.write_pose_txt(pp[-1], vw.pose)
        scene_io.write_depth_png(dp[-1], vw.depth_mm)
    return sc, views, scene_io.SceneFrames(sp, pp, dp)


def test_file_round_trips(tmp_path):
    sc, views, fr = _write_scene(str(tmp_path), 31, 20_000, 2)
    locs, feats, labels = scene_io.read_scene_pth(fr.scene_path)
    assert locs.dtype == np.float32 and np.array_equal(locs, sc.xyz)
    assert np.array_equal(feats, sc.colors) and np.array_equal(labels, sc.labels)
    for v, vw in enumerate(views):
        pose = scene_io.read_pose_txt(fr.pose_paths[v])
        assert pose.dtype == np.float64 and np.array_equal(pose, vw.pose)          # %.17g round-trips float64
        d = scene_io.read_depth_png(fr.depth_paths[v])
        assert d.dtype == np.uint16 and np.array_equal(d, vw.depth_mm)
        m = scene_io.depth_metres(d)                                             # imread(png) / 1000
        assert m.dtype == np.float64 and np.array_equal(m, vw.depth_mm / 1000)
    with pytest.raises(FileNotFoundError):
        scene_io.read_depth_png(os.path.join(str(tmp_path), "missing.png"))


def test_frame_batch_loader_matches_in_memory(tmp_path):
    items = [_write_scene(str(tmp_path), 40 + i, 5_000 + 777 * i, 2 + i) for i in range(2)]
    rts = {(si, fi): np.arange(12, dtype=np.float64).reshape(3, 4) + si + 0.5 * fi
           for si in range(2) for fi in range(3)}
    ld = scene_io.FrameBatchLoader(syn.scannet_intrinsics(), lambda si, fi: rts[(si, fi)], workers=4)
    batch, extras = ld.load([it[2] for it in items])
    ld.close()
    assert batch.n_views == 5 and batch.scene_off.tolist() == [0, 5000, 5000 + 5777]
    assert np.array_equal(batch.xyz, np.concatenate([it[0].xyz for it in items]))
    k = 0
    for si, (sc, views, _) in enumerate(items):
        assert np.array_equal(extras[si][1], sc.labels)
        for fi, vw in enumerate(views):
            assert batch.view_scene[k] == si
            assert np.array_equal(batch.w2c[k], np.linalg.inv(vw.pose))
            assert np.array_equal(batch.depth_mm[k], vw.depth_mm)
            assert np.array_equal(batch.rt[k], rts[(si, fi)])
            k += 1


@pytest.mark.gpu
def test_files_to_device_pipeline_and_collation(tmp_path):
    """files -> loader -> pinned staging ring -> project / voxelize -> device collation, against the
    in-memory path and the restated collation_fn."""
    from xmask3d_b200 import ops
    from xmask3d_b200.pipeline import CorrespondencePipeline
    from xmask3d_b200.voxelizer import Voxelizer
    dev = torch.device("cuda", 0)
    items = [_write_scene(str(tmp_path), 60 + i, 30_000 + 1234 * i, 2) for i in range(2)]
    rts = {}
    for si in range(2):
        for fi in range(2):
            np.random.seed(5557 + 1000 * si + fi)
            rt, _ = Voxelizer(voxel_size=0.02, use_augmentation=True, scale_augmentation_bound=(0.9, 1.1),
                              rotation_augmentation_bound=((-np.pi / 64, np.pi / 64), (-np.pi / 64, np.pi / 64),
                                                           (-np.pi, np.pi)),
                              translation_augmentation_ratio_bound=((-0.2, 0.2), (-0.2, 0.2), (0, 0))
                              ).draw_rigid_transformation()
            rts[(si, fi)] = rt
    ld = scene_io.FrameBatchLoader(syn.scannet_intrinsics(), lambda si, fi: rts[(si, fi)])
    batch, _ = ld.load([it[2] for it in items])
    ld.close()
    ring = scene_io.StagingRing(batch.xyz.shape[0], batch.n_views, device=dev)
    pipe = CorrespondencePipeline(batch, 4, 8, dev)
    for rep in range(3):                                   # walks around the ring
        xyz_d, dep_d, ready, slot = ring.stage(batch.xyz, batch.depth_mm)
        torch.cuda.current_stream().wait_event(ready)
        pipe.xyz, pipe.depth = xyz_d, dep_d
        pr = pipe.project()
        vox = ops.voxelize_batch(pr.xyz_vis, pr.vis_off, pipe.rt, cap=pipe.cap_vis, collate=True)
        ori, coords = ops.collate(pr, vox)
        ring.release(slot)
    torch.cuda.synchronize()
    vis_off = pr.vis_off.cpu().numpy()
    uniq_off = vox.uniq_off.cpu().numpy()
    n, M = int(vis_off[-1]), int(uniq_off[-1])
    # the same arrays computed from memory (no files, no ring)
    pipe2 = CorrespondencePipeline(batch, 4, 8, dev)
    pipe2.upload(torch.from_numpy(batch.xyz).pin_memory(), torch.from_numpy(batch.depth_mm.view(np.int16)).pin_memory())
    pr2 = pipe2.project()
    assert torch.equal(pr2.vis_off, pr.vis_off) and torch.equal(pr2.rowcol[:n], pr.rowcol[:n])
    # collation_fn restated on the per-view pieces
    xyz_vis = pr.xyz_vis.cpu().numpy()
    voxel = vox.voxel_xyz.cpu().numpy()
    inv = vox.inverse.cpu().numpy()
    loc_l, crd_l, ind_l = [], [], []
    for v in range(batch.n_views):
        a, b = vis_off[v], vis_off[v + 1]
        loc_l.append(xyz_vis[a:b])
        crd_l.append(voxel[uniq_off[v]:uniq_off[v + 1]])
        ind_l.append(inv[a:b] - uniq_off[v])               # per-sample inds_reconstruct (collate=True added the offset)
    want_loc, want_crd, want_ind = ref_port.collation_ref(loc_l, crd_l, ind_l)
    assert torch.equal(ori[:n].cpu(), want_loc)
    assert torch.equal(coords[:M].cpu(), want_crd)
    assert np.array_equal(inv[:n].astype(np.int64), want_ind.numpy())


def test_scene_cache_matches_loader_preprocessing(tmp_path):
    """The /dev/shm scene cache (dataset/point_loader.py:123-181): same keys, same preprocessing as the
    uncached branch (:182-188), idempotent init, attach returns private copies."""
    import numpy as np
    from xmask3d_b200 import scene_io
    rng = np.random.default_rng(0)
    paths = []
    raw = []
    for i in range(3):
        n = 500 + 37 * i
        locs = rng.uniform(-3, 3, (n, 3)).astype(np.float32)
        feats = rng.uniform(-1, 1, (n, 3)).astype(np.float32)
        labels = rng.integers(0, 20, n).astype(np.float64)
        labels[rng.uniform(size=n) < 0.1] = -100
        p = str(tmp_path / f"scene{i}.pth")
        scene_io.write_scene_pth(p, locs, feats, labels)
        paths.append(p)
        raw.append((locs, feats, labels))
    cache = scene_io.SceneCache("scannet_3d", "val", identifier=7, root=str(tmp_path))
    assert cache.key("locs", 2) == "scannet_3d_val_000007_locs_00000002"
    assert not cache.ready()
    assert cache.init(paths, workers=2) == 3 and cache.ready()
    assert cache.init(paths) == 0                                   # already there
    for i, (locs, feats, labels) in enumerate(raw):
        l2, f2, y2 = cache.attach(i)
        # the uncached branch of the loader, restated (dataset/point_loader.py:182-188)
        ref_labels = labels.copy()
        ref_labels[ref_labels == -100] = 255
        ref_labels = ref_labels.astype(np.uint8)
        assert np.array_equal(l2, locs) and l2.dtype == np.float32
        assert np.array_equal(f2, (feats + 1.0) * 127.5)
        assert np.array_equal(y2, ref_labels) and y2.dtype == np.uint8
        l2[0, 0] = 1e9                                              # a private copy: the cache is untouched
        assert cache.attach(i)[0][0, 0] == locs[0, 0]
    view = cache.attach(1, copy=False)[0]
    assert not view.flags.writeable
    cache.delete(3)
    assert not cache.ready()


def test_scene_without_colours(tmp_path):
    """dataset/point_loader.py:143-145: a scalar 0 colour entry means "no colours" -> zeros -> 127.5 per point,
    through read_scene_pth and the cache's preprocessing alike."""
    import torch
    from xmask3d_b200 import scene_io
    locs = np.random.default_rng(0).standard_normal((50, 3)).astype(np.float32)
    labels = np.arange(50) % 7
    p = str(tmp_path / "scene.pth")
    torch.save((locs, 0, labels), p)
    l2, f2, lab2 = scene_io.read_scene_pth(p)
    assert np.ndim(f2) == 0 and f2 == 0
    out_l, out_f, out_lab = scene_io.preprocess_scene(l2, f2, lab2)
    assert out_f.shape == locs.shape and np.all(out_f == 127.5)
    out_l, out_f, out_lab = scene_io.preprocess_scene(l2, np.asarray(0), lab2)        # 0-d array form
    assert out_f.shape == locs.shape and np.all(out_f == 127.5)
