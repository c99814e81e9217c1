// Batch layout the model consumes — the loaders' `collation_fn`
// (reference dataset/data_loader.py:319-357) on the device, from the path's own outputs:
//   ori_coords [sum n, 4] float32   column 0 = batch item (scene / view) index, 1..3 = xyz of the visible points
//                                   (locals_3d with `[:, 0] *= i`, data_loader.py:339)
//   coords     [sum M, 4] int32     column 0 = batch item index, 1..3 = voxel coordinates (coords[i][:, 0] *= i, :338)
// inds_reconstruct with the cumulative voxel offset (:340-341) is what xm3d_voxelize_batch writes
// with collate = 1; x_label / y_label are the two columns of rowcol.  One thread per row, 16-byte stores.
#include "common.cuh"

namespace xm3d {

__global__ void __launch_bounds__(256)
collate_points_kernel(const float *__restrict__ xyz, const int64_t *__restrict__ seg_off, int n_seg, int64_t cap,
                      float4 *__restrict__ out) {
    int64_t total = seg_off[n_seg];
    if (total > cap) total = 0;
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= total) return;
    const int s = seg_of(seg_off, n_seg, i);
    out[i] = make_float4((float)s, xyz[3 * i], xyz[3 * i + 1], xyz[3 * i + 2]);
}

__global__ void __launch_bounds__(256)
collate_voxels_kernel(const int32_t *__restrict__ voxel, const int64_t *__restrict__ uniq_off, int n_seg, int64_t cap,
                      int4 *__restrict__ out) {
    int64_t total = uniq_off[n_seg];
    if (total > cap) total = 0;
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= total) return;
    const int s = seg_of(uniq_off, n_seg, i);
    out[i] = make_int4(s, voxel[3 * i], voxel[3 * i + 1], voxel[3 * i + 2]);
}

// Compact host-bound copies of the index maps: pixel rows / columns (< 32768) and voxel coordinates fit int16, which
// halves the device -> host bytes of x_label / y_label / coords (the end-to-end loop is PCIe bound).  `rows` is the
// device-side row count (vis_off[n_seg] / uniq_off[n_seg]); rows beyond it are not touched.  A value outside
// int16 raises XM3D_FLAG_I16_RANGE (and is clamped).
__global__ void __launch_bounds__(256)
pack_i16_kernel(const int32_t *__restrict__ src, const int64_t *__restrict__ rows, int64_t cap_rows, int width,
                int16_t *__restrict__ dst, int *status) {
    int64_t n = rows ? *rows : cap_rows;
    if (n > cap_rows) n = cap_rows;
    n *= width;
    bool bad = false;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        int v = src[i];
        if (v < -32768 || v > 32767) { bad = true; v = v < 0 ? -32768 : 32767; }
        dst[i] = (int16_t)v;
    }
    if (bad && status) atomicOr(status, XM3D_FLAG_I16_RANGE);
}

}  // namespace xm3d

using namespace xm3d;

extern "C" int xm3d_pack_i16(const int32_t *src, const int64_t *rows_dev, int64_t cap_rows, int32_t width, int16_t *dst,
                             int32_t *status, xm3d_stream_t stream_) {
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    XM3D_REQUIRE(cap_rows >= 0 && width > 0, "bad sizes");
    if (cap_rows == 0) return XM3D_OK;
    XM3D_REQUIRE(src && dst, "null pointer");
    int64_t blocks = (cap_rows * width + 255) / 256;
    const int64_t maxb = (int64_t)sm_count() * 16;
    if (blocks > maxb) blocks = maxb;
    pack_i16_kernel<<<(unsigned)blocks, 256, 0, stream>>>(src, rows_dev, cap_rows, width, dst, status);
    count_launches(1);
    return check_launch("xm3d_pack_i16");
}

extern "C" int xm3d_collate_batch(const float *xyz_vis, const int64_t *vis_off, const int32_t *voxel_xyz,
                                  const int64_t *uniq_off, int32_t n_seg, int64_t cap, float *ori_coords,
                                  int32_t *coords, xm3d_stream_t stream_) {
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    XM3D_REQUIRE(n_seg > 0 && cap >= 0, "bad sizes");
    XM3D_REQUIRE((ori_coords == nullptr) || (xyz_vis && vis_off), "ori_coords needs xyz_vis and vis_off");
    XM3D_REQUIRE((coords == nullptr) || (voxel_xyz && uniq_off), "coords needs voxel_xyz and uniq_off");
    XM3D_REQUIRE(reinterpret_cast<uintptr_t>(ori_coords) % 16 == 0 && reinterpret_cast<uintptr_t>(coords) % 16 == 0,
                 "outputs must be 16-byte aligned");
    if (cap == 0) return XM3D_OK;
    const unsigned blocks = (unsigned)((cap + 255) / 256);
    if (ori_coords) {
        collate_points_kernel<<<blocks, 256, 0, stream>>>(xyz_vis, vis_off, n_seg, cap, reinterpret_cast<float4 *>(ori_coords));
        count_launches(1);
    }
    if (coords) {
        collate_voxels_kernel<<<blocks, 256, 0, stream>>>(voxel_xyz, uniq_off, n_seg, cap, reinterpret_cast<int4 *>(coords));
        count_launches(1);
    }
    return check_launch("xm3d_collate_batch");
}
