// After the path — mask preparation fused into one dense pass.
//
// Replaces, per (scene, view), the torch sequence both callers run before the masks reach the
// points (reference models/xmask3d.py:326-331, 356-358, 391-435; models/utils/criterion.py:239-244,
// 273-320):
//   up   = F.interpolate(mask_logits [K,hs,ws], size=(H,W), mode="bilinear", align_corners=False)
//   sig  = up.sigmoid()
//   (a)  mask_3d_full = sig[:, x_label, y_label] > 0.5                      -> per-pixel membership words
//   (b)  ids = (score[k] * sig[k]).argmax(0) over the kept masks            -> label image
//        mask_area[k] = (ids == k).sum(), original_area[k] = (sig[k] >= 0.5).sum(),
//        final mask k = (ids == k) & (sig[k] >= 0.5), kept iff it is non-empty
// The reference materialises [K,H,W] float32 four times (15 MB each per view) and runs a Python
// loop with three .item() synchronisations per mask.  Here one thread owns one output pixel, walks
// the K low-resolution planes (L1-resident, 64 KB each) and writes only the membership words, one
// int16 label and the per-mask areas; nothing of size K*H*W ever reaches HBM.
//
// Bilinear arithmetic = torch's CPU kernel bit for bit (checked in tests/ against F.interpolate):
//   src = max(fma(in/out, dst + 0.5, -0.5), 0); i0 = (int)src; i1 = i0 + (i0 < in-1); l1 = src - i0; l0 = 1 - l1
//   row = fma(lx0, v[i0], lx1 * v[i1]);   val = fma(ly0, row0, ly1 * row1)
#include "common.cuh"
#include "vec.cuh"

namespace xm3d {

constexpr int PREP_THREADS = 256;
constexpr int PREP_MAXK = 32 * MAX_WORDS;

struct PrepParams {
    const float *logits;       // [n_seg, k, hs, ws]
    const float *scores;       // [n_seg, k] or null
    const uint8_t *keep;       // [n_seg, k] or null (= all kept)
    int k, hs, ws, h, w, words, thr_mode;
    uint32_t *pixbits;         // [n_seg, words, h*w] or null
    int16_t *label;            // [n_seg, h*w] or null
    int32_t *areas;            // [n_seg, k, 3] or null: argmax area, original area, intersection
    float *up;                 // [n_seg, k, h*w] or null
};

__device__ __forceinline__ float sigmoid_f32(float x) {                     // torch.sigmoid, float32
    return __fdiv_rn(1.0f, __fadd_rn(1.0f, expf(-x)));
}

__device__ __forceinline__ void source_index(float scale, int dst, int in, int &i0, int &i1, float &l0, float &l1) {
    float src = __fmaf_rn(scale, (float)dst + 0.5f, -0.5f);
    src = src < 0.f ? 0.f : src;
    i0 = (int)src;
    if (i0 > in - 1) i0 = in - 1;
    i1 = i0 + (i0 < in - 1 ? 1 : 0);
    l1 = src - (float)i0;
    l1 = l1 < 0.f ? 0.f : (l1 > 1.f ? 1.f : l1);
    l0 = 1.f - l1;
}

// Per-pixel state of the mask loop, shared by the two kernels below.
// BITS: -1 = no membership words, else the threshold mode; PART: label image / areas; UP: dense output.
template <int BITS, bool PART, bool UP>
struct PixelLoop {
    const PrepParams &P;
    int (*s_area)[3];
    const float *sc;
    const uint8_t *kp;
    int s, p, hw, lane;
    bool live, want_areas;
    float best = -INFINITY;
    bool best_in = false;
    int id = -1;
    uint32_t word = 0u;

    __device__ __forceinline__ void consume(int m, float val) {
        if (UP) { if (live) P.up[((size_t)s * P.k + m) * hw + p] = val; }
        // sigmoid(x) vs 0.5 is the sign of x except within a few ulp of 0, where the float32
        // 1/(1+exp(-x)) rounds to exactly 0.5: only those values go through the exact formula
        bool ge_half = val > 0.f, gt_half = ge_half;
        if ((BITS == XM3D_THR_SIGMOID_GE_HALF || BITS == XM3D_THR_SIGMOID_GT_HALF || PART) && !(fabsf(val) > 1.0e-6f)) {
            const float sg = sigmoid_f32(val);
            ge_half = sg >= 0.5f; gt_half = sg > 0.5f;
        }
        if (BITS >= 0) {
            const bool hit = BITS == XM3D_THR_GE_HALF ? (val >= 0.5f) : (BITS == XM3D_THR_SIGMOID_GE_HALF ? ge_half : gt_half);
            word |= (hit ? 1u : 0u) << (m & 31);
            if ((m & 31) == 31 || m == P.k - 1) {
                if (live) P.pixbits[((size_t)s * P.words + (m >> 5)) * hw + p] = word;
                word = 0u;
            }
        }
        if (PART && (!kp || kp[m])) {
            // score * sigmoid is evaluated only when it can still beat the running maximum:
            // sigmoid <= 1, and <= 0.5 for a negative logit (both exact in float32)
            const float sm = sc ? sc[m] : 1.f;
            const float ub = val < 0.f ? 0.5f * sm : sm;
            if (sm < 0.f || ub > best) {
                const float sg = sigmoid_f32(val);
                const float prob = __fmul_rn(sm, sg);
                if (prob > best) { best = prob; id = m; best_in = ge_half; }      // strict: first maximum (torch.argmax)
            }
            if (want_areas) {
                const unsigned vote = __ballot_sync(0xffffffffu, live && ge_half);
                if (lane == 0 && vote) atomicAdd(&s_area[m][1], __popc(vote));
            }
        }
    }

    __device__ __forceinline__ void finish(int tid) {
        if (!PART) return;
        const bool in_mask = id >= 0 && best_in;
        if (P.label && live) P.label[(size_t)s * hw + p] = (int16_t)(in_mask ? id : -1);
        if (want_areas) {
            const int key = live ? id : -1;
            const unsigned grp = __match_any_sync(0xffffffffu, key);
            const unsigned inm = __ballot_sync(0xffffffffu, live && in_mask);
            if (key >= 0 && lane == __ffs(grp) - 1) {
                atomicAdd(&s_area[key][0], __popc(grp));
                const int ni = __popc(grp & inm);
                if (ni) atomicAdd(&s_area[key][2], ni);
            }
            __syncthreads();
            int32_t *ga = P.areas + (size_t)s * P.k * 3;
            for (int j = tid; j < P.k * 3; j += PREP_THREADS) {
                const int v = (&s_area[0][0])[j];
                if (v) atomicAdd(ga + j, v);
            }
        }
    }
};

// One CTA = a tile of 32 x 8 output pixels (a warp = 32 consecutive pixels of one row).
// TILED: the tile's source patch (<= PATCH_H x PATCH_W floats per plane) of up to PREP_KC planes is
// staged in shared memory first, so the per-(pixel, mask) loop is 4 LDS with immediate offsets +
// 6 float ops; otherwise (patch too large: strong down-sampling) the taps come straight from
// global memory.  (First version: one generic kernel reading global memory, 74 instructions per
// (pixel, mask), 22 of them 64-bit address arithmetic: 1.7 ms for 160 views x 50 masks.)
constexpr int PATCH_W = 16, PATCH_H = 8, PREP_KC = 64;
template <int BITS, bool PART, bool UP, bool TILED>
__global__ void __launch_bounds__(PREP_THREADS)
mask_prep_kernel(const PrepParams P) {
    __shared__ int s_area[PART ? PREP_MAXK : 1][3];
    __shared__ float s_patch[TILED ? PREP_KC * PATCH_H * PATCH_W : 1];
    const int s = blockIdx.y, tid = threadIdx.x, lane = tid & 31;
    const int hw = P.h * P.w;
    const int tiles_x = (P.w + 31) >> 5;
    const int ty = blockIdx.x / tiles_x, tx = blockIdx.x - ty * tiles_x;
    const int px = tx * 32 + lane, py = ty * (PREP_THREADS / 32) + (tid >> 5);
    const bool live = px < P.w && py < P.h;
    PixelLoop<BITS, PART, UP> L{P, s_area, P.scores ? P.scores + (size_t)s * P.k : nullptr,
                                P.keep ? P.keep + (size_t)s * P.k : nullptr, s, live ? py * P.w + px : 0, hw, lane,
                                live, PART && P.areas != nullptr};
    if (L.want_areas)
        for (int j = tid; j < P.k * 3; j += PREP_THREADS) (&s_area[0][0])[j] = 0;
    const float sy = (float)P.hs / (float)P.h, sx = (float)P.ws / (float)P.w;
    int y0, y1, x0, x1;
    float ly0, ly1, lx0, lx1;
    source_index(sy, live ? py : 0, P.hs, y0, y1, ly0, ly1);
    source_index(sx, live ? px : 0, P.ws, x0, x1, lx0, lx1);
    const size_t plane = (size_t)P.hs * P.ws;
    const float *base = P.logits + (size_t)s * P.k * plane;
    if (TILED) {
        // source patch of the tile (source_index is monotone in dst)
        int ya, yb, xa, xb, t0, t1;
        float f0, f1;
        source_index(sy, ty * (PREP_THREADS / 32), P.hs, ya, t1, f0, f1);
        source_index(sy, min(ty * (PREP_THREADS / 32) + PREP_THREADS / 32 - 1, P.h - 1), P.hs, t0, yb, f0, f1);
        source_index(sx, tx * 32, P.ws, xa, t1, f0, f1);
        source_index(sx, min(tx * 32 + 31, P.w - 1), P.ws, t0, xb, f0, f1);
        const int ph = yb - ya + 1, pw = xb - xa + 1;                 // <= PATCH_H, PATCH_W (checked by the host)
        const int t00 = ((live ? y0 - ya : 0) * PATCH_W + (live ? x0 - xa : 0));
        const int t01 = t00 + (x1 - x0), t10 = t00 + (y1 - y0) * PATCH_W, t11 = t10 + (x1 - x0);
        for (int m0 = 0; m0 < P.k; m0 += PREP_KC) {
            __syncthreads();                                           // previous chunk consumed (and s_area zeroed)
            const int kc = min(PREP_KC, P.k - m0);
            for (int e = tid; e < kc * (PATCH_H * PATCH_W); e += PREP_THREADS) {
                const int m = e / (PATCH_H * PATCH_W), r = (e / PATCH_W) % PATCH_H, c = e % PATCH_W;
                if (r < ph && c < pw)
                    s_patch[e] = __ldg(base + (size_t)(m0 + m) * plane + (size_t)(ya + r) * P.ws + (xa + c));
            }
            __syncthreads();
#pragma unroll 8
            for (int j = 0; j < kc; ++j) {
                const float *q = s_patch + j * (PATCH_H * PATCH_W);
                const float r0 = __fmaf_rn(lx0, q[t00], __fmul_rn(lx1, q[t01]));
                const float r1 = __fmaf_rn(lx0, q[t10], __fmul_rn(lx1, q[t11]));
                L.consume(m0 + j, __fmaf_rn(ly0, r0, __fmul_rn(ly1, r1)));
            }
        }
    } else {
        __syncthreads();
        const unsigned o00 = y0 * P.ws + x0, o01 = y0 * P.ws + x1, o10 = y1 * P.ws + x0, o11 = y1 * P.ws + x1;
        const float *pl = base;
        for (int m = 0; m < P.k; ++m, pl += plane) {
            const float r0 = __fmaf_rn(lx0, __ldg(pl + o00), __fmul_rn(lx1, __ldg(pl + o01)));
            const float r1 = __fmaf_rn(lx0, __ldg(pl + o10), __fmul_rn(lx1, __ldg(pl + o11)));
            L.consume(m, __fmaf_rn(ly0, r0, __fmul_rn(ly1, r1)));
        }
    }
    L.finish(tid);
}

template <int BITS, bool TILED>
static void launch_prep2(const PrepParams &P, bool part, bool up, dim3 grid, cudaStream_t stream) {
    if (part) {
        if (up) mask_prep_kernel<BITS, true, true, TILED><<<grid, PREP_THREADS, 0, stream>>>(P);
        else mask_prep_kernel<BITS, true, false, TILED><<<grid, PREP_THREADS, 0, stream>>>(P);
    } else {
        if (up) mask_prep_kernel<BITS, false, true, TILED><<<grid, PREP_THREADS, 0, stream>>>(P);
        else mask_prep_kernel<BITS, false, false, TILED><<<grid, PREP_THREADS, 0, stream>>>(P);
    }
}
template <int BITS>
static void launch_prep(const PrepParams &P, bool part, bool up, bool tiled, dim3 grid, cudaStream_t stream) {
    if (tiled) launch_prep2<BITS, true>(P, part, up, grid, stream);
    else launch_prep2<BITS, false>(P, part, up, grid, stream);
}

// per-point label = label image at the point's pixel (int16 -> int32; out of range -> -1)
__global__ void __launch_bounds__(256)
gather_labels_kernel(const int16_t *__restrict__ label_img, const int32_t *__restrict__ rowcol,
                     const int64_t *__restrict__ seg_off, int n_seg, int64_t cap, int h, int w,
                     int32_t *__restrict__ out) {
    int64_t total = seg_off[n_seg];
    if (total > cap) total = 0;
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= total) return;
    const int s = seg_of(seg_off, n_seg, i);
    const int2 rc = __ldg(reinterpret_cast<const int2 *>(rowcol) + i);
    const bool inb = rc.x >= 0 && rc.x < h && rc.y >= 0 && rc.y < w;
    out[i] = inb ? (int32_t)__ldg(label_img + (size_t)s * h * w + (size_t)rc.x * w + rc.y) : -1;
}

}  // namespace xm3d

using namespace xm3d;

extern "C" int xm3d_mask_prep_batch(const float *logits, int32_t n_seg, int32_t k, int32_t hs, int32_t ws, int32_t h,
                                    int32_t w, const float *scores, const uint8_t *keep, int32_t thr_mode,
                                    uint32_t *pixbits, int16_t *label, int32_t *areas, float *upsampled,
                                    xm3d_stream_t stream_) {
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    XM3D_REQUIRE(n_seg > 0 && k > 0 && hs > 0 && ws > 0 && h > 0 && w > 0, "bad sizes");
    XM3D_REQUIRE(k <= PREP_MAXK, "at most 256 masks per segment");
    XM3D_REQUIRE((int64_t)h * w < ((int64_t)1 << 31) && (int64_t)hs * ws < ((int64_t)1 << 31), "image too large");
    XM3D_REQUIRE(logits != nullptr, "null pointer");
    XM3D_REQUIRE(pixbits || label || areas || upsampled, "no output requested");
    XM3D_REQUIRE(thr_mode >= 0 && thr_mode <= 2, "bad thr_mode");
    PrepParams P;
    P.logits = logits; P.scores = scores; P.keep = keep; P.k = k; P.hs = hs; P.ws = ws; P.h = h; P.w = w;
    P.words = words_for(k); P.thr_mode = thr_mode; P.pixbits = pixbits; P.label = label; P.areas = areas; P.up = upsampled;
    if (areas) cudaMemsetAsync(areas, 0, sizeof(int32_t) * (size_t)n_seg * k * 3, stream);
    dim3 grid((unsigned)(((w + 31) / 32) * ((h + PREP_THREADS / 32 - 1) / (PREP_THREADS / 32))), (unsigned)n_seg);
    const bool part = label != nullptr || areas != nullptr, up = upsampled != nullptr;
    // the tile's source patch spans at most ceil(32 * ws / w) + 2 columns and ceil(8 * hs / h) + 2 rows
    const bool tiled = (32LL * ws + w - 1) / w + 2 <= PATCH_W && ((PREP_THREADS / 32) * (int64_t)hs + h - 1) / h + 2 <= PATCH_H;
    if (!pixbits) launch_prep<-1>(P, part, up, tiled, grid, stream);
    else if (thr_mode == XM3D_THR_GE_HALF) launch_prep<XM3D_THR_GE_HALF>(P, part, up, tiled, grid, stream);
    else if (thr_mode == XM3D_THR_SIGMOID_GE_HALF) launch_prep<XM3D_THR_SIGMOID_GE_HALF>(P, part, up, tiled, grid, stream);
    else launch_prep<XM3D_THR_SIGMOID_GT_HALF>(P, part, up, tiled, grid, stream);
    count_launches(1);
    return check_launch("xm3d_mask_prep_batch");
}

extern "C" int xm3d_gather_labels_batch(const int16_t *label_img, int32_t n_seg, int32_t h, int32_t w,
                                        const int32_t *rowcol, const int64_t *seg_off, int64_t cap,
                                        int32_t *point_label, xm3d_stream_t stream_) {
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    XM3D_REQUIRE(n_seg > 0 && h > 0 && w > 0 && cap >= 0, "bad sizes");
    XM3D_REQUIRE(label_img && rowcol && seg_off && point_label, "null pointer");
    if (cap == 0) return XM3D_OK;
    gather_labels_kernel<<<(unsigned)((cap + 255) / 256), 256, 0, stream>>>(label_img, rowcol, seg_off, n_seg, cap, h, w,
                                                                           point_label); count_launches(1);
    return check_launch("xm3d_gather_labels_batch");
}
