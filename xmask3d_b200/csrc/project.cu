// Stage 2 — point-to-pixel projection with the depth-occlusion test, batched over views.
//
// Replaces PointCloudToImageMapper.compute_mapping (reference models/utils/fusion_util.py:46-142)
// and the loaders' compaction of the visible points (dataset/data_loader_infer.py:174-182,
// 263-268).  One CTA owns PART consecutive points of one view:
//   * the 192-byte view record (world->camera rows, intrinsics) and, when it is a uint16
//     image that fits, the whole depth image are staged into shared memory with 1-D bulk
//     async copies (TMA engine) signalled through mbarriers; the camera-space transform of
//     all the CTA's points runs while the depth image is still in flight;
//   * xyz is streamed with 16-byte loads (4 points = 3 x float4 per thread);
//   * every float->int decision is the exact IEEE operation sequence numpy performs
//     (fp64 FMA chain for the dgemm dot products, individually rounded mul/div/add after
//     it, rint = half-to-even), see oracle/xm3d_oracle.c for the scalar restatement;
//   * visible points are compacted in point order inside the CTA (warp-shuffle scan) into a
//     part-local staging run; a one-CTA scan over the (view, part) counts then gives every
//     run its final offset and a relocation kernel writes vis_idx / rowcol / xyz_vis.
#include "common.cuh"

namespace xm3d {

// Four 256-thread CTAs per SM (depth read through L1/L2) or one 1024-thread CTA per SM with the view's
// depth image staged in shared memory: alone the kernel takes 227 vs 247 us, inside the step 1.773 / 1.787 ms
// vs 1.790 / 1.799 ms (two runs each) — the small CTAs also let the mask-bits pass of the pooling stream
// slip in next to the projection.
#ifndef XM3D_PROJ_THREADS
#define XM3D_PROJ_THREADS 256
#endif
#ifndef XM3D_PROJ_CTAS
#define XM3D_PROJ_CTAS 4
#endif
constexpr int PROJ_THREADS = XM3D_PROJ_THREADS;
constexpr int PROJ_CTAS = XM3D_PROJ_CTAS;            // persistent CTAs per SM (> 1: the depth image is not staged)
constexpr int PROJ_GROUP = 4;                       // consecutive points per thread per group
constexpr int PROJ_GROUPS = 2;                      // groups per thread
constexpr int PROJ_GROUP_PTS = PROJ_THREADS * PROJ_GROUP;     // 4096
constexpr int PROJ_PART = PROJ_GROUP_PTS * PROJ_GROUPS;       // 8192 points per CTA
constexpr int PROJ_MAX_SMEM_DEPTH = 160 * 1024;     // bytes of depth image we are willing to stage
constexpr int PROJ_SMEM_EXTRA = PROJ_PART * 4 + PROJ_PART * 2;   // candidate codes + queue

struct ProjParams {
    const float *xyz;
    const xm3d_view_t *views;   // device copy
    const void *depth;
    int depth_kind;
    double depth_scale;
    double img_w, img_h, cut;   // as doubles: bounds are tested on the rounded double
    double vis_thres;
    uint8_t *vis;
    int64_t *mapping;
    unsigned long long *stage;  // [total_pts] part-local compacted (idx<<32 | row<<16 | col)
    int *part_cnt;              // [n_views * parts]
    int *view_flag;             // [n_views] any inside point within the depth image (exact mode)
    int parts;
    const int *item_off;        // [n_views + 1] exclusive prefix of the parts of every view (device)
    int n_views;
    int smem_depth_bytes;
    int use_flag;               // 1: vis = flag ? inside&&ok : inside   (depth smaller than image)
    // float32 classifier constants (host-computed so that they are plain constant-bank operands)
    float f_lx, f_hx, f_ly, f_hy;   // rounding boundaries of the cut image: cut - 0.5, W - cut - 0.5, same for y
    float f_ecw, f_ech;             // 4e-7 (W + 1), 4e-7 (H + 1)
    float f_inv_scale, f_vt;        // 1 / depth_scale, visibility threshold
};

__device__ __forceinline__ double dot_row(const double *a, double x, double y, double z) {
    // numpy float64 matmul == one accumulator per output, FMA in k order (OpenBLAS dgemm kernel)
    double s = __dmul_rn(a[0], x);
    s = __fma_rn(a[1], y, s);
    s = __fma_rn(a[2], z, s);
    s = __fma_rn(a[3], 1.0, s);
    return s;
}

// Exact per-point arithmetic of compute_mapping (fusion_util.py:71-95): returns the packed code
// (bit31 inside, row<<16 | col) and the camera-space depth.
__device__ __forceinline__ uint32_t project_exact(const double *vw, float fx_, float fy_, float fz_,
                                                  const ProjParams &P, double *z_out) {
    const double x = (double)fx_, y = (double)fy_, z = (double)fz_;
    const double p0 = dot_row(vw + 0, x, y, z);                 // :71
    const double p1 = dot_row(vw + 4, x, y, z);
    const double p2 = dot_row(vw + 8, x, y, z);
    const double zdiv = (fabs(p2) < 1e-8) ? 1.0 : p2;           // :75-76
    const double px = __dadd_rn(__ddiv_rn(__dmul_rn(p0, vw[12]), zdiv), vw[14]);   // :78
    const double py = __dadd_rn(__ddiv_rn(__dmul_rn(p1, vw[13]), zdiv), vw[15]);   // :79
    const double rx = rint(px), ry = rint(py);                  // :82-83 (half to even)
    // :86-95 — tested on the integer-valued doubles (NaN / out-of-int64 values fail like the
    // INT64_MIN numpy's astype(int) produces for them)
    const bool inside = (p2 > 0.0) && (rx >= P.cut) && (ry >= P.cut) && (rx < P.img_w - P.cut) && (ry < P.img_h - P.cut);
    *z_out = p2;
    return inside ? (0x80000000u | ((uint32_t)(int)ry << 16) | (uint32_t)(int)rx) : 0u;
}

// Float32 classifier: decides a point WITHOUT the fp64 sequence whenever the float32 estimate is
// farther from every decision boundary of the exact arithmetic than a bound on its own error —
// behind the camera, outside the cut image, the rounded pixel (px, py more than `ex` away from a
// half-integer), and the occlusion test (|d - z| vs thres * d with a margin).  Everything else —
// borderline values, NaN / inf, anything near the z singularity — returns CLS_EXACT and is
// evaluated by project_exact, so the result is the exact path's by construction.
// Error bound of a float32 camera coordinate p_i = sum_j a_ij x_j: coefficient rounding plus three
// fused multiply-adds stay below 1e-6 * sum_j |a_ij| |x_j| (>= 8 ulp); |x_j| is bounded per work
// item by the largest |coordinate| of the CTA's points (one shared-memory max per axis).
// px32 = (p0 * fx) * (1 / p2) + cx:  |px32 - px| <= fx r e0 + |px - cx| (r e2 + RHO) + ecx, with
// r = 1 / p2, RHO = 1e-6 (>= 16 ulp: fx conversion, two products, the approximate reciprocal) and
// ecx = 4e-7 (|cx| + W + 1) for the conversion of cx and the final add.
constexpr uint32_t CLS_EXACT = 0xffffffffu;        // not a valid code: rows are < 32767
struct FilterConst {
    float a[12];             // world->camera rows
    float fx, fy, cx, cy;
    float e2, zmin;          // error bound of the float32 camera z for this work item; 64 e2 + 1e-3
    float e2r, fxe0, fye1, ecx, ecy;   // 1.05 e2 (1/(1-t) <= 1 + 1.04 t for t < 1/64), 1.02 fx e0, ...
    float lxc, hxc, lyc, hyc;          // boundaries relative to the principal point (coarse test)
};
struct DepthSrc {
    const unsigned short *sd;    // staged uint16 image or nullptr
    const void *g;               // global image (uint16 or float64)
    int64_t off;
    int kind, dh, dw;
    bool has;
};
// Coarse reject in multiplied form (no reciprocal): true only if the exact arithmetic is certain
// to give inside == false.  With e2 / p2 < 1/64 the classifier's ex * p2 is below
// fxe0 + 0.0166 |p0 fx| + ecx p2; the test doubles the last term for its own roundings.
__device__ __forceinline__ bool coarse_reject(const FilterConst &F, float x, float y, float z) {
    const float p2 = fmaf(F.a[8], x, fmaf(F.a[9], y, fmaf(F.a[10], z, F.a[11])));
    if (p2 + F.e2 <= 0.f) return true;                     // exact p2 <= 0: not in front
    if (!(p2 > F.zmin)) return false;                      // near the z singularity (or NaN): not decided here
    const float t = fmaf(F.a[0], x, fmaf(F.a[1], y, fmaf(F.a[2], z, F.a[3]))) * F.fx;
    const float s = fmaf(F.a[4], x, fmaf(F.a[5], y, fmaf(F.a[6], z, F.a[7]))) * F.fy;
    const float mx = fmaf(0.0166f, fabsf(t), fmaf(2.f * F.ecx, p2, F.fxe0));
    const float my = fmaf(0.0166f, fabsf(s), fmaf(2.f * F.ecy, p2, F.fye1));
    return fmaf(-F.lxc, p2, t) < -mx || fmaf(-F.hxc, p2, t) > mx ||
           fmaf(-F.lyc, p2, s) < -my || fmaf(-F.hyc, p2, s) > my;
}

template <bool FLAG_ONLY>
__device__ __forceinline__ uint32_t classify_fast(const ProjParams &P, const FilterConst &F, const DepthSrc &D, float x, float y, float z,
                                                  int use_flag, int vflag, bool *any_in_depth) {
    const float p2 = fmaf(F.a[8], x, fmaf(F.a[9], y, fmaf(F.a[10], z, F.a[11])));
    if (p2 + F.e2 <= 0.f) return 0u;                       // exact p2 <= 0: not in front
    if (!(p2 > F.zmin)) return CLS_EXACT;   // near the z singularity (or NaN); e2 / p2 < 1/64 below
    const float p0 = fmaf(F.a[0], x, fmaf(F.a[1], y, fmaf(F.a[2], z, F.a[3])));
    const float p1 = fmaf(F.a[4], x, fmaf(F.a[5], y, fmaf(F.a[6], z, F.a[7])));
    const float r = __fdividef(1.f, p2);
    const float u = (p0 * F.fx) * r, w = (p1 * F.fy) * r;
    const float px = u + F.cx, py = w + F.cy;
    const float rel = fmaf(F.e2r, r, 1.0e-6f);
    const float ex = fmaf(fabsf(u), rel, fmaf(F.fxe0, r, F.ecx));
    const float ey = fmaf(fabsf(w), rel, fmaf(F.fye1, r, F.ecy));
    if (px < P.f_lx - ex || px > P.f_hx + ex || py < P.f_ly - ey || py > P.f_hy + ey) return 0u;   // outside for sure
    const float nx = rintf(px), ny = rintf(py);
    if (!(fabsf(px - nx) < 0.5f - ex) || !(fabsf(py - ny) < 0.5f - ey)) return CLS_EXACT;   // rounding not certain
    // nx, ny are the exact path's rx, ry; the boundaries lx.. sit on half-integers
    if (!(nx > P.f_lx && nx < P.f_hx && ny > P.f_ly && ny < P.f_hy)) return 0u;
    const int ix = (int)nx, iy = (int)ny;
    uint32_t code = 0x80000000u | ((uint32_t)iy << 16) | (uint32_t)ix;
    if (D.has) {                                           // fusion_util.py:98-135
        const bool in_depth = iy < D.dh && ix < D.dw;
        bool ok = false;
        if (in_depth) {
            *any_in_depth = true;
            if (!FLAG_ONLY) {
                const size_t e = (size_t)iy * D.dw + ix;
                float d;
                if (D.kind == XM3D_DEPTH_U16) {
                    const unsigned short raw = D.sd ? D.sd[e] : __ldg(reinterpret_cast<const unsigned short *>(D.g) + D.off + e);
                    d = (float)raw * P.f_inv_scale;
                } else {
                    d = (float)__ldg(reinterpret_cast<const double *>(D.g) + D.off + e);
                }
                const float aa = fabsf(d - p2), bb = P.f_vt * d;
                const float m = F.e2 + 2.0e-6f * (fabsf(d) + p2 + fabsf(bb));
                if (aa <= bb - m) ok = true;
                else if (aa > bb + m) ok = false;
                else return CLS_EXACT;                     // occlusion test too close to call (or NaN / inf)
            }
        }
        if (!FLAG_ONLY) {
            const bool keep = use_flag ? (vflag ? ok : true) : ok;
            if (!keep) code = 0u;
        }
    }
    return code;
}

template <bool FLAG_ONLY>
__global__ void __launch_bounds__(PROJ_THREADS, PROJ_CTAS) project_kernel(const ProjParams P) {
    // dynamic shared memory: [depth image (P.smem_depth_bytes)] [codes u32 x PART] [queue u16 x PART]
    extern __shared__ __align__(128) unsigned char smem_raw[];
    __shared__ __align__(16) double s_view[24];      // the 192-byte record
    __shared__ uint64_t s_bar[2];
    __shared__ int s_warp_tot[PROJ_THREADS / 32];
    __shared__ int s_any;
    __shared__ int s_amax[3];                        // max |x|, |y|, |z| of the item's points (float bits)
    uint32_t *s_code = reinterpret_cast<uint32_t *>(smem_raw + P.smem_depth_bytes);
    unsigned short *s_queue = reinterpret_cast<unsigned short *>(smem_raw + P.smem_depth_bytes + PROJ_PART * 4);

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    // persistent CTA: a contiguous run of (view, part) work items; the view record and the depth
    // image stay in shared memory while consecutive items belong to the same view
    const int n_items = P.item_off[P.n_views];
    const int it0 = (int)((int64_t)n_items * blockIdx.x / gridDim.x);
    const int it1 = (int)((int64_t)n_items * (blockIdx.x + 1) / gridDim.x);
    if (it0 >= it1) return;
    if (tid == 0) {
        mbar_init(&s_bar[0], 1);
        mbar_init(&s_bar[1], 1);
        mbar_fence_init();
        s_any = 0;
        s_amax[0] = s_amax[1] = s_amax[2] = 0;
    }
    int v = 0;
    {   // view of the first item: largest v with item_off[v] <= it0
        int lo = 0, hi = P.n_views;
        while (hi - lo > 1) {
            const int mid = (lo + hi) >> 1;
            if (P.item_off[mid] <= it0) lo = mid; else hi = mid;
        }
        v = lo;
    }
    int cur_view = -1;
    uint32_t ph = 0;                              // parity of both barriers (they complete together)
    bool staged = false, has_depth = false;
    int dh = 0, dw = 0, n_pts = 0;
    int64_t depth_off = -1;
    const xm3d_view_t *gv = nullptr;
  for (int it = it0; it < it1; ++it) {
    while (it >= P.item_off[v + 1]) ++v;
    const int part = it - P.item_off[v];
    const int64_t part_start = (int64_t)part * PROJ_PART;
    __syncthreads();                               // previous item done with s_code / depth / s_any
    bool stage_depth = staged;
    if (v != cur_view) {
        cur_view = v;
        gv = P.views + v;
        n_pts = gv->n_pts;
        depth_off = gv->depth_off;
        dh = gv->depth_h; dw = gv->depth_w;
        has_depth = P.depth_kind != XM3D_DEPTH_NONE && depth_off >= 0;
        const size_t depth_bytes = (size_t)dh * dw * 2;
        stage_depth = !FLAG_ONLY && has_depth && P.depth_kind == XM3D_DEPTH_U16 && depth_bytes <= (size_t)P.smem_depth_bytes &&
                      (depth_bytes % 16 == 0) &&
                      ((reinterpret_cast<uintptr_t>(P.depth) + (size_t)depth_off * 2) % 16 == 0);
        staged = stage_depth;
        if (tid == 0) {
            mbar_expect_tx(&s_bar[0], (uint32_t)sizeof(xm3d_view_t));
            bulk_g2s(s_view, gv, (uint32_t)sizeof(xm3d_view_t), &s_bar[0]);
            if (stage_depth) {
                mbar_expect_tx(&s_bar[1], (uint32_t)depth_bytes);
                const char *src = reinterpret_cast<const char *>(P.depth) + (size_t)depth_off * 2;
                for (size_t o = 0; o < depth_bytes; o += 32768) {
                    const uint32_t chunk = (uint32_t)((depth_bytes - o < 32768) ? depth_bytes - o : 32768);
                    bulk_g2s(smem_raw + o, src + o, chunk, &s_bar[1]);
                }
            } else {
                mbar_expect_tx(&s_bar[1], 0);    // keep both barriers in the same phase
            }
        }
        __syncthreads();
        mbar_wait(&s_bar[0], ph);
        if (!stage_depth) mbar_wait(&s_bar[1], ph);
    }

    // ---- phase 1: stream the points (16-byte loads), float32 reject filter, queue the candidates
    const float *xyz = P.xyz + gv->pt_off * 3;
    const bool vec_ok = (reinterpret_cast<uintptr_t>(xyz) % 16 == 0);
    float c[PROJ_GROUPS][PROJ_GROUP * 3];
#pragma unroll
    for (int g = 0; g < PROJ_GROUPS; ++g) {
        const int64_t i0 = part_start + (int64_t)g * PROJ_GROUP_PTS + (int64_t)tid * PROJ_GROUP;
        if (vec_ok && i0 + PROJ_GROUP <= n_pts) {
            const float4 *q = reinterpret_cast<const float4 *>(xyz + i0 * 3);
            const float4 a = __ldg(q), b = __ldg(q + 1), d = __ldg(q + 2);
            c[g][0] = a.x; c[g][1] = a.y; c[g][2] = a.z; c[g][3] = a.w;
            c[g][4] = b.x; c[g][5] = b.y; c[g][6] = b.z; c[g][7] = b.w;
            c[g][8] = d.x; c[g][9] = d.y; c[g][10] = d.z; c[g][11] = d.w;
        } else {
#pragma unroll
            for (int j = 0; j < PROJ_GROUP * 3; ++j) {
                const int64_t e = i0 * 3 + j;
                c[g][j] = (e < (int64_t)n_pts * 3) ? __ldg(xyz + e) : 0.f;
            }
        }
    }
    // largest |coordinate| per axis over the item's points (non-negative floats order like ints)
    {
        float mx = 0.f, my = 0.f, mz = 0.f;
#pragma unroll
        for (int g = 0; g < PROJ_GROUPS; ++g)
#pragma unroll
            for (int j = 0; j < PROJ_GROUP; ++j) {
                mx = fmaxf(mx, fabsf(c[g][3 * j])); my = fmaxf(my, fabsf(c[g][3 * j + 1])); mz = fmaxf(mz, fabsf(c[g][3 * j + 2]));
                // fmaxf drops NaNs: keep them so that a NaN coordinate disables the filter
                if (!(c[g][3 * j] == c[g][3 * j]) || !(c[g][3 * j + 1] == c[g][3 * j + 1]) || !(c[g][3 * j + 2] == c[g][3 * j + 2]))
                    mx = __int_as_float(0x7fc00000);
            }
        int ix = __float_as_int(mx), iy = __float_as_int(my), iz = __float_as_int(mz);
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            ix = max(ix, __shfl_xor_sync(0xffffffffu, ix, o));
            iy = max(iy, __shfl_xor_sync(0xffffffffu, iy, o));
            iz = max(iz, __shfl_xor_sync(0xffffffffu, iz, o));
        }
        if (lane == 0) { atomicMax(&s_amax[0], ix); atomicMax(&s_amax[1], iy); atomicMax(&s_amax[2], iz); }
    }
    __syncthreads();
    FilterConst F;
#pragma unroll
    for (int j = 0; j < 12; ++j) F.a[j] = (float)s_view[j];
    F.fx = (float)s_view[12]; F.fy = (float)s_view[13]; F.cx = (float)s_view[14]; F.cy = (float)s_view[15];
    {
        const float GAM = 1.0e-6f;
        const float X = __int_as_float(s_amax[0]), Y = __int_as_float(s_amax[1]), Z = __int_as_float(s_amax[2]);
        const float e0 = GAM * fmaf(fabsf(F.a[0]), X, fmaf(fabsf(F.a[1]), Y, fmaf(fabsf(F.a[2]), Z, fabsf(F.a[3]))));
        const float e1 = GAM * fmaf(fabsf(F.a[4]), X, fmaf(fabsf(F.a[5]), Y, fmaf(fabsf(F.a[6]), Z, fabsf(F.a[7]))));
        F.e2 = GAM * fmaf(fabsf(F.a[8]), X, fmaf(fabsf(F.a[9]), Y, fmaf(fabsf(F.a[10]), Z, fabsf(F.a[11]))));
        F.zmin = fmaf(64.f, F.e2, 1.0e-3f);
        F.e2r = 1.05f * F.e2; F.fxe0 = 1.02f * F.fx * e0; F.fye1 = 1.02f * F.fy * e1;
    }
    // the classifier's inequalities assume fx, fy > 0 and sane intrinsics
    const bool filter_ok = s_view[12] > 0.0 && s_view[13] > 0.0 && s_view[12] < 1e6 && s_view[13] < 1e6 &&
                           fabs(s_view[14]) < 1e6 && fabs(s_view[15]) < 1e6;
    F.ecx = fmaf(4.0e-7f, fabsf(F.cx), P.f_ecw); F.ecy = fmaf(4.0e-7f, fabsf(F.cy), P.f_ech);
    F.lxc = P.f_lx - F.cx; F.hxc = P.f_hx - F.cx; F.lyc = P.f_ly - F.cy; F.hyc = P.f_hy - F.cy;
    DepthSrc D;
    D.has = has_depth; D.kind = P.depth_kind; D.dh = dh; D.dw = dw; D.off = depth_off; D.g = P.depth;
    D.sd = stage_depth ? reinterpret_cast<const unsigned short *>(smem_raw) : nullptr;
    const int vflag = (!FLAG_ONLY && P.use_flag) ? P.view_flag[v] : 0;
    if (stage_depth) mbar_wait(&s_bar[1], ph);     // no-op once the phase has completed

    // ---- phase 1: coarse float32 reject of every point (points arrive in scene order, so almost
    // every warp holds a few inside points: the full classifier would run for all of them); the
    // survivors are queued per warp.  Every warp owns its 256 points end to end: no block barrier
    // until the ordered compaction.
    unsigned short *wq = s_queue + warp * (PROJ_GROUPS * PROJ_GROUP * 32);
    int wcand = 0;
    bool any_in_depth = false;
#pragma unroll
    for (int g = 0; g < PROJ_GROUPS; ++g)
#pragma unroll
        for (int j = 0; j < PROJ_GROUP; ++j) {
            const int local = g * PROJ_GROUP_PTS + tid * PROJ_GROUP + j;
            const bool live = part_start + local < n_pts;
            const bool cand = live && !(filter_ok && coarse_reject(F, c[g][3 * j], c[g][3 * j + 1], c[g][3 * j + 2]));
            const unsigned vote = __ballot_sync(0xffffffffu, cand);
            if (cand) wq[wcand + __popc(vote & ((1u << lane) - 1u))] = (unsigned short)local;
            wcand += __popc(vote);
        }
    if (!FLAG_ONLY) {
#pragma unroll
        for (int g = 0; g < PROJ_GROUPS; ++g)
            *reinterpret_cast<uint4 *>(s_code + g * PROJ_GROUP_PTS + tid * PROJ_GROUP) = make_uint4(0u, 0u, 0u, 0u);
    }
    __syncwarp();

    // ---- phase 1b: the float32 classifier on the survivors, one per lane; the undecided ones are
    // queued again (the second queue overwrites entries that have already been consumed)
    const float *xyz_item = xyz + part_start * 3;
    int wcount = 0;
    for (int q0 = 0; q0 < wcand; q0 += 32) {
        const bool act = q0 + lane < wcand;
        const int local = act ? wq[q0 + lane] : 0;
        uint32_t code = 0u;
        if (act) {
            const float *pt = xyz_item + local * 3;
            code = filter_ok ? classify_fast<FLAG_ONLY>(P, F, D, __ldg(pt), __ldg(pt + 1), __ldg(pt + 2),
                                                        P.use_flag, vflag, &any_in_depth)
                             : CLS_EXACT;
        }
        const bool und = code == CLS_EXACT;
        if (!FLAG_ONLY && act && !und) s_code[local] = code;
        __syncwarp();
        const unsigned vote = __ballot_sync(0xffffffffu, und);
        if (und) wq[wcount + __popc(vote & ((1u << lane) - 1u))] = (unsigned short)local;
        wcount += __popc(vote);
    }
    __syncwarp();

    // ---- phase 2: the exact fp64 sequence for the undecided points, one per lane
    const unsigned short *sd = reinterpret_cast<const unsigned short *>(smem_raw);
    for (int q = lane; q < wcount; q += 32) {
        const int local = wq[q];
        const float *pt = xyz_item + local * 3;
        double zc;
        uint32_t code = project_exact(s_view, __ldg(pt), __ldg(pt + 1), __ldg(pt + 2), P, &zc);
        if (has_depth && (code >> 31)) {                         // fusion_util.py:98-135
            const int iy = (code >> 16) & 0x7fff, ix = code & 0xffff;
            const bool in_depth = iy < dh && ix < dw;            // both are >= 0 here (cut >= 0)
            bool ok = false;
            if (in_depth) {
                any_in_depth = true;
                double d;
                const size_t e = (size_t)iy * dw + ix;
                if (P.depth_kind == XM3D_DEPTH_U16) {
                    const unsigned short raw = stage_depth
                        ? sd[e] : __ldg(reinterpret_cast<const unsigned short *>(P.depth) + depth_off + e);
                    d = __ddiv_rn((double)raw, P.depth_scale);   // imread(png) / 1000
                } else {
                    d = __ldg(reinterpret_cast<const double *>(P.depth) + depth_off + e);
                }
                ok = fabs(__dsub_rn(d, zc)) <= __dmul_rn(P.vis_thres, d);     // :125
            }
            if (!FLAG_ONLY) {
                // exact mode: keep `inside` when no inside point of the view hits the depth image
                const bool keep = P.use_flag ? (vflag ? ok : true) : ok;
                if (!keep) code = 0u;
            }
        }
        if (!FLAG_ONLY) s_code[local] = code;
    }
    if (FLAG_ONLY) {
        if (any_in_depth) s_any = 1;
        __syncthreads();
        if (tid == 0) {
            if (s_any) atomicOr(&P.view_flag[v], 1);
            s_any = 0;                               // ordered before the next item by its leading barrier
            s_amax[0] = s_amax[1] = s_amax[2] = 0;
        }
        if (it + 1 < it1 && it + 1 >= P.item_off[v + 1]) ph ^= 1;      // next item starts a new view
        continue;
    }
    __syncthreads();

    // ---- phase 3: visibility bytes, optional int64 [N,3] mapping, part-local compaction (point order)
    const int64_t out0 = gv->out_off;
    uint32_t code[PROJ_GROUPS][PROJ_GROUP];
    int packed = 0;                                  // visible count of group g in bits [16g, 16g+16)
#pragma unroll
    for (int g = 0; g < PROJ_GROUPS; ++g) {
        const int64_t i0 = part_start + (int64_t)g * PROJ_GROUP_PTS + (int64_t)tid * PROJ_GROUP;
        const uint4 cd = *reinterpret_cast<const uint4 *>(s_code + g * PROJ_GROUP_PTS + tid * PROJ_GROUP);
        code[g][0] = cd.x; code[g][1] = cd.y; code[g][2] = cd.z; code[g][3] = cd.w;
        int cnt = 0;
#pragma unroll
        for (int j = 0; j < PROJ_GROUP; ++j) cnt += code[g][j] >> 31;
        packed |= cnt << (16 * g);
        if (i0 < n_pts) {
            uint8_t *vp = P.vis + out0 + i0;
            if (i0 + PROJ_GROUP <= n_pts && (reinterpret_cast<uintptr_t>(vp) % 4 == 0)) {
                *reinterpret_cast<uint32_t *>(vp) = (code[g][0] >> 31) | ((code[g][1] >> 31) << 8) |
                                                    ((code[g][2] >> 31) << 16) | ((code[g][3] >> 31) << 24);
            } else {
#pragma unroll
                for (int j = 0; j < PROJ_GROUP; ++j)
                    if (i0 + j < n_pts) vp[j] = (uint8_t)(code[g][j] >> 31);
            }
            if (P.mapping) {
                int64_t *mp = P.mapping + (out0 + i0) * 3;
#pragma unroll
                for (int j = 0; j < PROJ_GROUP; ++j)
                    if (i0 + j < n_pts) {
                        const bool vis = code[g][j] >> 31;
                        mp[3 * j + 0] = vis ? (int64_t)((code[g][j] >> 16) & 0x7fff) : 0;   // row (y)
                        mp[3 * j + 1] = vis ? (int64_t)(code[g][j] & 0xffff) : 0;           // col (x)
                        mp[3 * j + 2] = vis ? 1 : 0;
                    }
            }
        }
    }
    // one block-exclusive scan serves both groups (order = thread order = point order inside a
    // group; a group holds <= 4096 visible points, so the 16-bit fields never carry)
    static_assert(PROJ_GROUPS == 2 && PROJ_GROUP_PTS <= 32768, "packed scan layout");
    int incl = packed;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const int t = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o) incl += t;
    }
    if (lane == 31) s_warp_tot[warp] = incl;
    __syncthreads();
    if (warp == 0) {
        const int w = lane < PROJ_THREADS / 32 ? s_warp_tot[lane] : 0;
        int wi = w;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int t = __shfl_up_sync(0xffffffffu, wi, o);
            if (lane >= o) wi += t;
        }
        if (lane < PROJ_THREADS / 32) s_warp_tot[lane] = wi - w;               // exclusive warp offsets
        if (lane == 31) s_any = wi;              // totals of both groups
    }
    __syncthreads();
    const int excl = s_warp_tot[warp] + incl - packed;
    const int tot0 = s_any & 0xffff, tot1 = s_any >> 16;
    unsigned long long *st = P.stage + out0 + part_start;
#pragma unroll
    for (int g = 0; g < PROJ_GROUPS; ++g) {
        const int64_t i0 = part_start + (int64_t)g * PROJ_GROUP_PTS + (int64_t)tid * PROJ_GROUP;
        int pos = g == 0 ? (excl & 0xffff) : tot0 + (excl >> 16);
#pragma unroll
        for (int j = 0; j < PROJ_GROUP; ++j)
            if (code[g][j] >> 31) {
                const unsigned long long idx = (unsigned long long)(i0 + j);
                st[pos++] = (idx << 32) | (unsigned long long)(code[g][j] & 0x7fffffffu);
            }
    }
    if (tid == 0) {
        P.part_cnt[v * P.parts + part] = tot0 + tot1;
        s_amax[0] = s_amax[1] = s_amax[2] = 0;
    }
    if (it + 1 < it1 && it + 1 >= P.item_off[v + 1]) ph ^= 1;          // next item starts a new view
  }
}

// parts of every view (exclusive prefix) for the persistent kernel; part counts zeroed
__global__ void __launch_bounds__(1024, 1)
project_plan_kernel(const xm3d_view_t *__restrict__ views, int n_views, int parts, int *__restrict__ item_off,
                    int *__restrict__ part_cnt) {
    __shared__ int s_w[32];
    __shared__ int s_carry;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (tid == 0) s_carry = 0;
    for (int e = tid; e < n_views * parts; e += 1024) part_cnt[e] = 0;
    __syncthreads();
    for (int base = 0; base < n_views; base += 1024) {
        const int v = base + tid;
        const int val = v < n_views ? (views[v].n_pts + PROJ_PART - 1) / PROJ_PART : 0;
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
        if (v < n_views) item_off[v] = excl;
        __syncthreads();
        if (tid == 1023) s_carry = excl + val;
        __syncthreads();
    }
    if (tid == 0) item_off[n_views] = s_carry;
}

// Exclusive scan over the (view-major) part counts: one CTA, chunks of 1024 with a carry.
__global__ void __launch_bounds__(1024, 1)
project_scan_kernel(const int *__restrict__ part_cnt, int n_views, int parts, int64_t *__restrict__ part_off,
                    int *__restrict__ n_vis, int64_t *__restrict__ vis_off, int64_t cap_vis, int *status) {
    __shared__ int64_t s_warp[32];
    __shared__ int64_t s_carry;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int total = n_views * parts;
    if (tid == 0) s_carry = 0;
    __syncthreads();
    for (int base = 0; base < total; base += 1024) {
        const int e = base + tid;
        const int64_t val = (e < total) ? part_cnt[e] : 0;
        int64_t incl = val;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int64_t t = __shfl_up_sync(0xffffffffu, incl, o);
            if (lane >= o) incl += t;
        }
        if (lane == 31) s_warp[warp] = incl;
        __syncthreads();
        if (warp == 0) {
            const int64_t w = s_warp[lane];
            int64_t wi = w;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const int64_t t = __shfl_up_sync(0xffffffffu, wi, o);
                if (lane >= o) wi += t;
            }
            s_warp[lane] = wi - w;
        }
        __syncthreads();
        const int64_t excl = s_carry + s_warp[warp] + incl - val;
        if (e < total) {
            part_off[e] = excl;
            if (e % parts == 0) vis_off[e / parts] = excl;
        }
        __syncthreads();
        if (tid == 1023) s_carry = excl + val;
        __syncthreads();
    }
    if (tid == 0) {
        vis_off[n_views] = s_carry;
        if (s_carry > cap_vis && status) atomicOr(status, XM3D_FLAG_VIS_OVERFLOW);
    }
    __syncthreads();
    for (int v = tid; v < n_views; v += 1024) {
        const int64_t a = vis_off[v];
        const int64_t b = (v + 1 < n_views) ? part_off[(v + 1) * parts] : s_carry;
        n_vis[v] = (int)(b - a);
    }
}

__global__ void __launch_bounds__(256)
project_emit_kernel(const float *__restrict__ xyz, const xm3d_view_t *__restrict__ views,
                    const unsigned long long *__restrict__ stage, const int *__restrict__ part_cnt,
                    const int64_t *__restrict__ part_off, int parts, int64_t cap_vis,
                    int32_t *__restrict__ vis_idx, int32_t *__restrict__ rowcol, float *__restrict__ xyz_vis) {
    const int v = blockIdx.y, part = blockIdx.x;
    const int cnt = part_cnt[v * parts + part];
    if (cnt == 0) return;
    const xm3d_view_t *gv = views + v;
    const int64_t dst0 = part_off[v * parts + part];
    const unsigned long long *st = stage + gv->out_off + (int64_t)part * PROJ_PART;
    const float *src = xyz + gv->pt_off * 3;
    for (int j = threadIdx.x; j < cnt; j += blockDim.x) {
        const int64_t d = dst0 + j;
        if (d >= cap_vis) break;
        const unsigned long long e = st[j];
        const uint32_t idx = (uint32_t)(e >> 32), rc = (uint32_t)e;
        if (vis_idx) vis_idx[d] = (int32_t)idx;
        if (rowcol) *reinterpret_cast<int2 *>(rowcol + 2 * d) = make_int2((int)(rc >> 16), (int)(rc & 0xffff));
        if (xyz_vis) {
            const float *p = src + (size_t)idx * 3;
            xyz_vis[3 * d + 0] = __ldg(p);
            xyz_vis[3 * d + 1] = __ldg(p + 1);
            xyz_vis[3 * d + 2] = __ldg(p + 2);
        }
    }
}

static int parts_for(int max_pts) {
    const int p = (max_pts + PROJ_PART - 1) / PROJ_PART;
    return p < 1 ? 1 : p;
}

}  // namespace xm3d

using namespace xm3d;

extern "C" size_t xm3d_project_ws_bytes(int32_t n_views, int64_t total_pts, int32_t max_pts_per_view) {
    Carver c(nullptr);
    const int parts = parts_for(max_pts_per_view);
    c.take<xm3d_view_t>(n_views);
    c.take<unsigned long long>(total_pts + PROJ_PART);
    c.take<int>((size_t)n_views * parts);
    c.take<int64_t>((size_t)n_views * parts);
    c.take<int>(n_views);
    c.take<int>(n_views + 1);
    return c.off + 256;
}

extern "C" int xm3d_project_batch(const float *xyz, const xm3d_view_t *views_host, const xm3d_view_t *views_dev,
                                  int32_t n_views, int64_t total_pts, const void *depth, int32_t depth_kind, double depth_scale,
                                  int32_t img_w, int32_t img_h, int32_t cut_bound, double vis_thres,
                                  uint8_t *vis, int64_t *mapping, int32_t *n_vis, int64_t *vis_off,
                                  int64_t cap_vis, int32_t *vis_idx, int32_t *rowcol, float *xyz_vis,
                                  void *ws, size_t ws_bytes, int32_t *status, xm3d_stream_t stream_) {
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    XM3D_REQUIRE(n_views >= 0 && total_pts >= 0, "negative size");
    XM3D_REQUIRE(n_views == 0 || (xyz && views_host && vis && n_vis && vis_off && ws), "null pointer");
    XM3D_REQUIRE(depth_kind == XM3D_DEPTH_NONE || depth_kind == XM3D_DEPTH_U16 || depth_kind == XM3D_DEPTH_F64,
                 "bad depth_kind");
    XM3D_REQUIRE(depth_kind == XM3D_DEPTH_NONE || depth != nullptr, "depth_kind set but depth is null");
    XM3D_REQUIRE(img_w > 0 && img_h > 0 && img_w <= 65535 && img_h <= 32767, "image size out of range");
    XM3D_REQUIRE(cut_bound >= 0, "cut_bound must be >= 0");
    if (n_views == 0) return XM3D_OK;

    int max_pts = 0;
    int64_t sum_pts = 0;
    bool covers = true, any_depth = false;
    size_t stage_bytes = 0;
    for (int v = 0; v < n_views; ++v) {
        const xm3d_view_t &w = views_host[v];
        XM3D_REQUIRE(w.n_pts >= 0 && w.pt_off >= 0 && w.out_off >= 0, "bad view record");
        XM3D_REQUIRE(w.out_off + w.n_pts <= total_pts, "view outputs exceed total_pts");
        max_pts = w.n_pts > max_pts ? w.n_pts : max_pts;
        sum_pts += w.n_pts;
        if (depth_kind != XM3D_DEPTH_NONE && w.depth_off >= 0) {
            XM3D_REQUIRE(w.depth_h > 0 && w.depth_w > 0, "bad depth size");
            any_depth = true;
            // every inside pixel lies in the depth image -> "any inside" implies "any in depth"
            if (w.depth_h < img_h - cut_bound || w.depth_w < img_w - cut_bound) covers = false;
            const size_t b = (size_t)w.depth_h * w.depth_w * 2;
            if (depth_kind == XM3D_DEPTH_U16 && b <= (size_t)PROJ_MAX_SMEM_DEPTH && b > stage_bytes) stage_bytes = b;
        }
    }
    const int parts = parts_for(max_pts);
    if (ws_bytes < xm3d_project_ws_bytes(n_views, total_pts, max_pts)) {
        set_error("xm3d_project_batch: workspace too small");
        return XM3D_ERR_WORKSPACE;
    }
    Carver c(ws);
    xm3d_view_t *d_views = c.take<xm3d_view_t>(n_views);
    unsigned long long *stage = c.take<unsigned long long>(total_pts + PROJ_PART);
    int *part_cnt = c.take<int>((size_t)n_views * parts);
    int64_t *part_off = c.take<int64_t>((size_t)n_views * parts);
    int *view_flag = c.take<int>(n_views);
    int *item_off = c.take<int>(n_views + 1);

    if (views_dev) d_views = const_cast<xm3d_view_t *>(views_dev);
    else cudaMemcpyAsync(d_views, views_host, sizeof(xm3d_view_t) * n_views, cudaMemcpyHostToDevice, stream);

    ProjParams P;
    P.xyz = xyz; P.views = d_views; P.depth = depth; P.depth_kind = depth_kind; P.depth_scale = depth_scale;
    P.img_w = img_w; P.img_h = img_h; P.cut = cut_bound; P.vis_thres = vis_thres;
    P.vis = vis; P.mapping = mapping; P.stage = stage; P.part_cnt = part_cnt; P.view_flag = view_flag;
    if (PROJ_CTAS > 1) stage_bytes = 0;
    stage_bytes = (stage_bytes + 127) / 128 * 128;
    P.item_off = item_off; P.n_views = n_views;
    P.f_lx = (float)(cut_bound - 0.5); P.f_hx = (float)(img_w - cut_bound - 0.5);
    P.f_ly = (float)(cut_bound - 0.5); P.f_hy = (float)(img_h - cut_bound - 0.5);
    P.f_ecw = 4.0e-7f * (float)(img_w + 1); P.f_ech = 4.0e-7f * (float)(img_h + 1);
    P.f_inv_scale = (float)(1.0 / depth_scale); P.f_vt = (float)vis_thres;
    P.parts = parts; P.smem_depth_bytes = (int)stage_bytes; P.use_flag = (any_depth && !covers) ? 1 : 0;

    static std::atomic<uint64_t> attr_set{0};
    if (first_use_on_device(&attr_set)) {
        cudaFuncSetAttribute(project_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, PROJ_MAX_SMEM_DEPTH + PROJ_SMEM_EXTRA);
        cudaFuncSetAttribute(project_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, PROJ_MAX_SMEM_DEPTH + PROJ_SMEM_EXTRA);
    }
    dim3 grid(parts, n_views);
    project_plan_kernel<<<1, 1024, 0, stream>>>(d_views, n_views, parts, item_off, part_cnt); count_launches(1);
    const unsigned pgrid = (unsigned)sm_count() * PROJ_CTAS;       // persistent: PROJ_CTAS per SM
    if (P.use_flag) {
        cudaMemsetAsync(view_flag, 0, sizeof(int) * n_views, stream);
        project_kernel<true><<<pgrid, PROJ_THREADS, stage_bytes + PROJ_SMEM_EXTRA, stream>>>(P); count_launches(1);
    }
    project_kernel<false><<<pgrid, PROJ_THREADS, stage_bytes + PROJ_SMEM_EXTRA, stream>>>(P); count_launches(1);
    project_scan_kernel<<<1, 1024, 0, stream>>>(part_cnt, n_views, parts, part_off, n_vis, vis_off, cap_vis, status); count_launches(1);
    if (vis_idx || rowcol || xyz_vis) {
        project_emit_kernel<<<grid, 256, 0, stream>>>(xyz, d_views, stage, part_cnt, part_off, parts, cap_vis,
                                                      vis_idx, rowcol, xyz_vis); count_launches(1); }
    return check_launch("xm3d_project_batch");
}
