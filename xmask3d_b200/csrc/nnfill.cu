// After the path at inference (second half):
//
//  * nearest seen neighbour of every point no view has seen — the reference builds
//    sklearn.neighbors.KDTree(scene_coords[counter != 0]) and queries the unseen points with k = 1
//    (run/infer.py:651-656, 684-694), then copies the neighbour's prediction.  Here: a uniform grid
//    over the seen points of every scene (counting sort by cell; cells are numbered block-major, 4 x 4 x 4
//    cells per block, so the points of a block are contiguous) and a two-level ring search per unseen
//    point: rings of cells up to radius 3, then rings of blocks (empty blocks cost two loads), each
//    stopping as soon as the best distance is inside the searched cube.  Distances are float64
//    sums of squared float64 differences in x, y, z order (what the KD tree computes on the float32
//    coordinates promoted to float64); equal distances resolve to the lowest point index.
//  * per-scene maximum of the sparse bottleneck features (models/xmask3d.py:154-159:
//    torch.max(imp_condition[_idx_ == scene_idx], dim=0)[0] per scene).
#include <math_constants.h>

#include "common.cuh"

namespace xm3d {

struct __align__(16) NnSeg {
    float mn[3];
    float inv_h;
    int g[3];
    int n_seen;
    float h;
    int pad[3];
};

constexpr int NN_THREADS = 256;
// cells of segment s live at cells[4 seg_off[s] + 64 s ...): at most 4 n + 64 of them
__device__ __forceinline__ int64_t nn_cell_base(const int64_t *seg_off, int s) { return 4 * seg_off[s] + 64 * (int64_t)s; }

// cell (cx, cy, cz) -> index: blocks of 4 x 4 x 4 cells, block-major (g[] are multiples of 4)
__device__ __forceinline__ int nn_cell_index(const NnSeg &S, int cx, int cy, int cz) {
    const int bx = S.g[0] >> 2, by = S.g[1] >> 2;
    const int blk = ((cz >> 2) * by + (cy >> 2)) * bx + (cx >> 2);
    return blk * 64 + ((cz & 3) << 4) + ((cy & 3) << 2) + (cx & 3);
}
__device__ __forceinline__ void nn_cell_coords(const NnSeg &S, float x, float y, float z, int &cx, int &cy, int &cz) {
    cx = (int)floorf((x - S.mn[0]) * S.inv_h); cy = (int)floorf((y - S.mn[1]) * S.inv_h); cz = (int)floorf((z - S.mn[2]) * S.inv_h);
    cx = min(max(cx, 0), S.g[0] - 1); cy = min(max(cy, 0), S.g[1] - 1); cz = min(max(cz, 0), S.g[2] - 1);
}
__device__ __forceinline__ int nn_cell_of(const NnSeg &S, float x, float y, float z) {
    int cx, cy, cz;
    nn_cell_coords(S, x, y, z, cx, cy, cz);
    return nn_cell_index(S, cx, cy, cz);
}

// one CTA per scene: bounding box and count of the seen points -> grid geometry
__global__ void __launch_bounds__(NN_THREADS)
nn_plan_kernel(const float *__restrict__ xyz, const int32_t *__restrict__ counter, const int64_t *__restrict__ seg_off,
               NnSeg *__restrict__ segs) {
    __shared__ float s_mn[NN_THREADS / 32][3], s_mx[NN_THREADS / 32][3];
    __shared__ int s_cnt[NN_THREADS / 32];
    const int s = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int64_t a = seg_off[s], b = seg_off[s + 1];
    float mn[3] = {CUDART_INF_F, CUDART_INF_F, CUDART_INF_F}, mx[3] = {-CUDART_INF_F, -CUDART_INF_F, -CUDART_INF_F};
    int cnt = 0;
    for (int64_t i = a + tid; i < b; i += NN_THREADS)
        if (counter[i] != 0) {
            ++cnt;
#pragma unroll
            for (int d = 0; d < 3; ++d) {
                const float v = xyz[i * 3 + d];
                if (v == v && fabsf(v) < 1e30f) { mn[d] = fminf(mn[d], v); mx[d] = fmaxf(mx[d], v); }
            }
        }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        cnt += __shfl_xor_sync(0xffffffffu, cnt, o);
#pragma unroll
        for (int d = 0; d < 3; ++d) {
            mn[d] = fminf(mn[d], __shfl_xor_sync(0xffffffffu, mn[d], o));
            mx[d] = fmaxf(mx[d], __shfl_xor_sync(0xffffffffu, mx[d], o));
        }
    }
    if (lane == 0) {
        s_cnt[warp] = cnt;
        for (int d = 0; d < 3; ++d) { s_mn[warp][d] = mn[d]; s_mx[warp][d] = mx[d]; }
    }
    __syncthreads();
    if (tid == 0) {
        NnSeg S;
        int n_seen = 0;
        for (int w = 0; w < NN_THREADS / 32; ++w) {
            n_seen += s_cnt[w];
            for (int d = 0; d < 3; ++d) { mn[d] = fminf(mn[d], s_mn[w][d]); mx[d] = fmaxf(mx[d], s_mx[w][d]); }
        }
        float ext[3];
        for (int d = 0; d < 3; ++d) {
            if (!(mn[d] <= mx[d])) { mn[d] = 0.f; mx[d] = 0.f; }
            ext[d] = fmaxf(mx[d] - mn[d], 1e-6f);
            S.mn[d] = mn[d];
        }
        // about two cells per seen point if they filled the box; surfaces leave most cells empty
        const int64_t budget = 4 * (b - a) + 64;
        float h = cbrtf(ext[0] * ext[1] * ext[2] / fmaxf(2.f * (float)n_seen, 1.f));
        h = fmaxf(h, 1e-6f);
        for (;;) {                                 // grow the cell until the grid fits the budget
            int64_t cells = 1;
            for (int d = 0; d < 3; ++d) {
                S.g[d] = ((int)fminf(floorf(ext[d] / h) + 1.f, 2048.f) + 3) & ~3;      // whole 4 x 4 x 4 blocks
                cells *= S.g[d];
            }
            if (cells <= budget) break;
            h *= 1.26f;
        }
        S.h = h; S.inv_h = 1.f / h; S.n_seen = n_seen;
        S.pad[0] = S.pad[1] = S.pad[2] = 0;
        segs[s] = S;
    }
}

// MODE 0: count the seen points of every cell; MODE 1: write them into cell order
template <int MODE>
__global__ void __launch_bounds__(NN_THREADS)
nn_bin_kernel(const float *__restrict__ xyz, const int32_t *__restrict__ counter, const int64_t *__restrict__ seg_off,
              int n_seg, int64_t n_total, const NnSeg *__restrict__ segs, int *__restrict__ cells,
              int *__restrict__ cursor, int *__restrict__ sorted) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_total || counter[i] == 0) return;
    const int s = seg_of(seg_off, n_seg, i);
    const NnSeg S = segs[s];
    const int c = nn_cell_of(S, xyz[i * 3], xyz[i * 3 + 1], xyz[i * 3 + 2]);
    const int64_t cb = nn_cell_base(seg_off, s);
    if (MODE == 0) atomicAdd(&cells[cb + c], 1);
    else sorted[seg_off[s] + atomicAdd(&cursor[cb + c], 1)] = (int)(i - seg_off[s]);
}

// one CTA per scene: counts -> exclusive starts (cells) and a copy as the fill cursor
__global__ void __launch_bounds__(1024)
nn_scan_kernel(const int64_t *__restrict__ seg_off, const NnSeg *__restrict__ segs, int *__restrict__ cells,
               int *__restrict__ cursor) {
    __shared__ int s_w[32];
    __shared__ int s_carry;
    const int s = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const NnSeg S = segs[s];
    const int ncell = S.g[0] * S.g[1] * S.g[2];
    int *c = cells + nn_cell_base(seg_off, s), *cur = cursor + nn_cell_base(seg_off, s);
    if (tid == 0) s_carry = 0;
    __syncthreads();
    for (int base = 0; base < ncell; base += 1024) {
        const int val = base + tid < ncell ? c[base + tid] : 0;
        int incl = val;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int t = __shfl_up_sync(0xffffffffu, incl, o);
            if (lane >= o) incl += t;
        }
        if (lane == 31) s_w[warp] = incl;
        __syncthreads();
        if (warp == 0) {
            const int w = s_w[lane];
            int wi = w;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const int t = __shfl_up_sync(0xffffffffu, wi, o);
                if (lane >= o) wi += t;
            }
            s_w[lane] = wi - w;
        }
        __syncthreads();
        const int excl = s_carry + s_w[warp] + incl - val;
        if (base + tid < ncell) { c[base + tid] = excl; cur[base + tid] = excl; }
        __syncthreads();
        if (tid == 1023) s_carry = excl + val;
        __syncthreads();
    }
}

// nearest seen point of every unseen point (ring search); seen points map to themselves
__global__ void __launch_bounds__(NN_THREADS)
nn_query_kernel(const float *__restrict__ xyz, const int32_t *__restrict__ counter, const int64_t *__restrict__ seg_off,
                int n_seg, int64_t n_total, const NnSeg *__restrict__ segs, const int *__restrict__ cells,
                const int *__restrict__ cursor, const int *__restrict__ sorted, int32_t *__restrict__ match) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_total) return;
    const int s = seg_of(seg_off, n_seg, i);
    const int64_t a = seg_off[s];
    if (counter[i] != 0) { match[i] = (int32_t)(i - a); return; }
    const NnSeg S = segs[s];
    if (S.n_seen == 0) { match[i] = -1; return; }
    const float fx = xyz[i * 3], fy = xyz[i * 3 + 1], fz = xyz[i * 3 + 2];
    const double px = fx, py = fy, pz = fz;
    int cx, cy, cz;
    nn_cell_coords(S, fx, fy, fz, cx, cy, cz);
    const int *cs = cells + nn_cell_base(seg_off, s), *ce = cursor + nn_cell_base(seg_off, s);
    const int *srt = sorted + a;
    const float *base = xyz + a * 3;
    double best = CUDART_INF;
    int best_i = -1;
    auto scan = [&](int q0, int q1) {
        for (int q = q0; q < q1; ++q) {
            const int j = srt[q];
            const double dx = px - (double)base[j * 3], dy = py - (double)base[j * 3 + 1], dz = pz - (double)base[j * 3 + 2];
            const double d = __dadd_rn(__dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy)), __dmul_rn(dz, dz));
            if (d < best || (d == best && j < best_i)) { best = d; best_i = j; }
        }
    };
    // Once every cell (block) within Chebyshev distance R of the query's own cell (block) has been scanned,
    // any other seen point differs by more than R cell (block) sizes along some axis — also when the query
    // lies outside the grid and its cell was clamped.  The cell size carries float32 rounding: 0.999.
    bool done = false;
    constexpr int NN_FINE_RINGS = 3;
    for (int r = 0; r <= NN_FINE_RINGS && !done; ++r) {
        if (r > 0) {
            const double reach = 0.999 * (double)(r - 1) * (double)S.h;
            if (best <= reach * reach) { done = true; break; }
        }
        const int z0 = max(cz - r, 0), z1 = min(cz + r, S.g[2] - 1);
        const int y0 = max(cy - r, 0), y1 = min(cy + r, S.g[1] - 1);
        for (int z = z0; z <= z1; ++z)
            for (int y = y0; y <= y1; ++y) {
                const bool face = (z == cz - r) || (z == cz + r) || (y == cy - r) || (y == cy + r);
                const int xs = (face || r == 0) ? 1 : 2 * r;     // interior rows: only the two end cells
                for (int x = cx - r; x <= cx + r; x += xs) {
                    if (x < 0 || x >= S.g[0]) continue;
                    const int c = nn_cell_index(S, x, y, z);
                    scan(cs[c], ce[c]);
                }
            }
    }
    if (!done) {
        {   // rings 0..3 of cells are complete
            const double reach = 0.999 * (double)NN_FINE_RINGS * (double)S.h;
            done = best <= reach * reach;
        }
        const int bxn = S.g[0] >> 2, byn = S.g[1] >> 2, bzn = S.g[2] >> 2;
        const int bx = cx >> 2, by = cy >> 2, bz = cz >> 2;
        const double bh = 4.0 * (double)S.h;
        const int rmax = max(bxn, max(byn, bzn));
        for (int r = 0; r <= rmax && !done; ++r) {
            if (r > 0) {
                const double reach = 0.999 * (double)(r - 1) * bh;
                if (best <= reach * reach) break;
            }
            const int z0 = max(bz - r, 0), z1 = min(bz + r, bzn - 1);
            const int y0 = max(by - r, 0), y1 = min(by + r, byn - 1);
            for (int z = z0; z <= z1; ++z)
                for (int y = y0; y <= y1; ++y) {
                    const bool face = (z == bz - r) || (z == bz + r) || (y == by - r) || (y == by + r);
                    const int xs = (face || r == 0) ? 1 : 2 * r;
                    for (int x = bx - r; x <= bx + r; x += xs) {
                        if (x < 0 || x >= bxn) continue;
                        // skip a block that cannot hold anything closer than the best so far (distance from
                        // the query to the block's box, shrunk by 0.1 % of a block for the float32 cell maths)
                        const double lo[3] = {(double)S.mn[0] + x * bh, (double)S.mn[1] + y * bh, (double)S.mn[2] + z * bh};
                        const double q3[3] = {px, py, pz};
                        double gap2 = 0.0;
                        const bool last[3] = {x == bxn - 1, y == byn - 1, z == bzn - 1};   // clamped coordinates end up here
#pragma unroll
                        for (int d = 0; d < 3; ++d) {
                            const double above = last[d] ? 0.0 : q3[d] - (lo[d] + bh);
                            const double g = fmax(fmax(lo[d] - q3[d], above), 0.0) - 0.001 * bh;
                            if (g > 0.0) gap2 += g * g;
                        }
                        if (gap2 >= best) continue;
                        const int blk = ((z * byn + y) * bxn + x) * 64;
                        scan(cs[blk], ce[blk + 63]);            // the 64 cells of a block are contiguous
                    }
                }
        }
    }
    match[i] = best_i;
}

// out[s, c] = max over the rows of segment s (NaN propagates as in torch.max); -inf for an empty segment
__global__ void __launch_bounds__(256)
segment_max_kernel(const float *__restrict__ feat, const int64_t *__restrict__ seg_off, int c, float *__restrict__ out) {
    __shared__ float s_part[8][32];
    const int s = blockIdx.y;
    const int ch = blockIdx.x * 32 + (threadIdx.x & 31);
    const int rg = threadIdx.x >> 5;                          // 8 row groups of 32 channels
    const int64_t a = seg_off[s], b = seg_off[s + 1];
    float m = -CUDART_INF_F;
    bool nan = false;
    if (ch < c)
        for (int64_t r = a + rg; r < b; r += 8) {
            const float v = __ldg(feat + r * c + ch);
            nan |= !(v == v);
            m = fmaxf(m, v);
        }
    s_part[rg][threadIdx.x & 31] = nan ? CUDART_NAN_F : m;
    __syncthreads();
    if (rg == 0 && ch < c) {
        float r = s_part[0][threadIdx.x];
        bool isn = !(r == r);
        for (int g = 1; g < 8; ++g) {
            const float v = s_part[g][threadIdx.x];
            isn |= !(v == v);
            r = fmaxf(r, v);
        }
        out[(int64_t)s * c + ch] = isn ? CUDART_NAN_F : r;
    }
}

struct NnWs { NnSeg *segs; int *cells, *cursor, *sorted; size_t cell_count; };
static NnWs carve_nn(void *ws, int n_seg, int64_t n_total, size_t *bytes) {
    Carver c(ws);
    NnWs w;
    w.cell_count = (size_t)(4 * n_total + 64 * (int64_t)n_seg);
    w.segs = c.take<NnSeg>((size_t)n_seg);
    w.cells = c.take<int>(w.cell_count);
    w.cursor = c.take<int>(w.cell_count);
    w.sorted = c.take<int>((size_t)n_total + 1);
    *bytes = c.off + 256;
    return w;
}

}  // namespace xm3d

using namespace xm3d;

extern "C" size_t xm3d_nn_fill_ws_bytes(int32_t n_seg, int64_t n_total) {
    size_t b = 0;
    carve_nn(nullptr, n_seg, n_total, &b);
    return b;
}

extern "C" int xm3d_nn_fill_batch(const float *xyz, const int32_t *counter, const int64_t *seg_off, int32_t n_seg,
                                  int64_t n_total, int32_t *match, void *ws, size_t ws_bytes, xm3d_stream_t stream_) {
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    XM3D_REQUIRE(n_seg > 0 && n_total >= 0, "bad sizes");
    if (n_total == 0) return XM3D_OK;
    XM3D_REQUIRE(xyz && counter && seg_off && match && ws, "null pointer");
    XM3D_REQUIRE(n_total < ((int64_t)1 << 29), "n_total must stay below 2^29");
    size_t need = 0;
    const NnWs w = carve_nn(ws, n_seg, n_total, &need);
    if (ws_bytes < need) {
        set_error("xm3d_nn_fill_batch: workspace too small (%zu < %zu)", ws_bytes, need);
        return XM3D_ERR_WORKSPACE;
    }
    const unsigned blocks = (unsigned)((n_total + NN_THREADS - 1) / NN_THREADS);
    cudaMemsetAsync(w.cells, 0, w.cell_count * sizeof(int), stream);
    nn_plan_kernel<<<n_seg, NN_THREADS, 0, stream>>>(xyz, counter, seg_off, w.segs);
    nn_bin_kernel<0><<<blocks, NN_THREADS, 0, stream>>>(xyz, counter, seg_off, n_seg, n_total, w.segs, w.cells, w.cursor, w.sorted);
    nn_scan_kernel<<<n_seg, 1024, 0, stream>>>(seg_off, w.segs, w.cells, w.cursor);
    nn_bin_kernel<1><<<blocks, NN_THREADS, 0, stream>>>(xyz, counter, seg_off, n_seg, n_total, w.segs, w.cells, w.cursor, w.sorted);
    nn_query_kernel<<<blocks, NN_THREADS, 0, stream>>>(xyz, counter, seg_off, n_seg, n_total, w.segs, w.cells, w.cursor, w.sorted, match);
    count_launches(5);
    return check_launch("xm3d_nn_fill_batch");
}

extern "C" int xm3d_segment_max(const float *feat, const int64_t *seg_off, int32_t n_seg, int32_t c, float *out,
                                xm3d_stream_t stream) {
    XM3D_REQUIRE(n_seg > 0 && c > 0, "bad sizes");
    XM3D_REQUIRE(feat && seg_off && out, "null pointer");
    segment_max_kernel<<<dim3((unsigned)((c + 31) / 32), (unsigned)n_seg), 256, 0, static_cast<cudaStream_t>(stream)>>>(
        feat, seg_off, c, out);
    count_launches(1);
    return check_launch("xm3d_segment_max");
}
