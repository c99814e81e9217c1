"""Property tests (hypothesis) that pin the two restatements of the reference algorithm against each other:
the scalar C oracle (oracle/xm3d_oracle.c, exact IEEE operation sequence) and the numpy / torch port
(oracle/ref_port.py, op for op what the reference executes) must agree BIT FOR BIT on random inputs —
projection mappings with and without depth, FNV keys, np.unique maps, voxel grids.  CPU only."""
import numpy as np
from hypothesis import given, settings, strategies as st

from oracle import ref_port
from xmask3d_b200 import synthetic as syn


def _pose(rng):
    yaw, pitch = rng.uniform(0, 2 * np.pi), rng.uniform(-0.5, 0.1)
    cy, sy, cp, sp = np.cos(yaw), np.sin(yaw), np.cos(pitch), np.sin(pitch)
    rot = np.array([[cy, 0, sy], [0, 1, 0], [-sy, 0, cy]]) @ np.array([[1, 0, 0], [0, cp, -sp], [0, sp, cp]])
    pose = np.eye(4)
    pose[:3, :3] = rot
    pose[:3, 3] = rng.uniform(-1, 1, 3)
    return pose


@settings(max_examples=60, deadline=None, derandomize=True)
@given(seed=st.integers(0, 2 ** 31 - 1), n=st.integers(1, 3000), with_depth=st.booleans(), small_depth=st.booleans())
def test_projection_c_oracle_equals_numpy_port(cport, seed, n, with_depth, small_depth):
    rng = np.random.default_rng(seed)
    xyz = rng.uniform(-4, 4, (n, 3)).astype(np.float32)
    xyz[rng.random(n) < 0.02] = 0.0                                   # points at the camera-frame singularity
    pose = _pose(rng)
    mapper = ref_port.getMapping()
    depth_m = None
    if with_depth:
        h, w = (120, 160) if small_depth else (240, 320)              # a depth image smaller than the camera image
        depth_mm = rng.integers(0, 6000, (h, w)).astype(np.uint16)
        depth_mm[rng.random((h, w)) < 0.1] = 0
        depth_m = depth_mm / 1000.0
    ref = mapper.compute_mapping(pose, xyz, depth_m)
    got = cport.project(xyz, np.linalg.inv(pose), syn.scannet_intrinsics(), depth_m)
    assert np.array_equal(got, ref)
    if with_depth:                                                    # the uint16 entry divides by 1000 itself
        got16 = cport.project(xyz, np.linalg.inv(pose), syn.scannet_intrinsics(), depth_mm)
        assert np.array_equal(got16, ref)


@settings(max_examples=60, deadline=None, derandomize=True)
@given(seed=st.integers(0, 2 ** 31 - 1), n=st.integers(1, 4000), span=st.sampled_from([3, 40, 600]),
       voxel=st.sampled_from([0.02, 0.05, 0.2]))
def test_voxelize_c_oracle_equals_numpy_port(cport, seed, n, span, voxel):
    rng = np.random.default_rng(seed)
    # keys / unique maps on integer grids of very different density (heavy collisions .. none)
    grid = np.floor(rng.uniform(0, span, (n, 3)))
    keys = ref_port.fnv_hash_vec(grid)
    assert np.array_equal(cport.fnv(grid), keys)
    _, ref_first, ref_inv, ref_cnt = np.unique(keys, return_index=True, return_inverse=True, return_counts=True)
    first, inv, cnt = cport.unique_u64(keys)
    assert np.array_equal(first, ref_first) and np.array_equal(inv, ref_inv) and np.array_equal(cnt, ref_cnt)
    # the whole voxelize() with a drawn augmentation matrix
    xyz = rng.uniform(-3, 3, (n, 3)).astype(np.float32)
    np.random.seed(seed % (2 ** 31))
    vox = ref_port.Voxelizer(voxel_size=voxel, use_augmentation=True, scale_augmentation_bound=(0.9, 1.1),
                             rotation_augmentation_bound=((-0.05, 0.05), (-0.05, 0.05), (-np.pi, np.pi)))
    state = np.random.get_state()
    M_v, M_r = vox.get_transformation_matrix()
    np.random.set_state(state)
    colors = np.zeros((n, 3), np.float32)
    labels = np.zeros(n)
    g_ref, _, _, inv_ref, inds_ref = vox.voxelize(xyz, colors, labels, return_ind=True)
    g, first, inv = cport.voxelize(xyz, M_r @ M_v)
    assert np.array_equal(g, g_ref) and np.array_equal(first, inds_ref) and np.array_equal(inv, inv_ref)


@settings(max_examples=40, deadline=None, derandomize=True)
@given(seed=st.integers(0, 2 ** 31 - 1), n=st.integers(1, 600), k=st.integers(1, 40), c=st.sampled_from([1, 7, 64]),
       density=st.sampled_from([0.0, 0.05, 0.5]))
def test_scatter_and_pool_c_oracle_equals_torch_port(cport, seed, n, k, c, density):
    """mask -> point scatter-mean (fuser.py:22-34) bit for bit in float32, masked mean pooling
    (criterion.py:152-157) against the float64 sums of the C oracle."""
    import torch
    rng = np.random.default_rng(seed)
    member = rng.random((k, n)) < density
    emb = rng.standard_normal((k, c)).astype(np.float32)
    feat = rng.standard_normal((n, c)).astype(np.float32)
    out, counter = cport.scatter_member_f32(member, emb)
    ref_out, ref_counter = ref_port.scatter_mask_embed(torch.from_numpy(member), torch.from_numpy(emb),
                                                       torch.zeros(n, c))
    guard = member.copy()
    if not member.any():
        guard[0, 0] = True                                  # fuser.py:19-20, applied by the caller of the kernel
        out, counter = cport.scatter_member_f32(guard, emb)
    assert np.array_equal(out, ref_out.numpy())
    assert np.array_equal(np.where(counter == 0, np.float32(1e-5), counter), ref_counter.numpy()[:, 0])
    s64, cnt = cport.pool_member_f64(feat, member)
    mean, ref_cnt = ref_port.masked_mean_pool(torch.from_numpy(feat), torch.from_numpy(member))
    assert np.array_equal(cnt, ref_cnt.numpy())
    ref64 = s64 / np.maximum(cnt, 1)[:, None]
    assert np.abs(mean.numpy() - ref64).max() <= 1e-5 * max(1.0, np.abs(ref64).max())
