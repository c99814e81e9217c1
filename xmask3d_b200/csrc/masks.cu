// Stage 3 — masks at points, segmented mean pooling, mask -> point scatter-mean.
//
//   gather   mask[:, x_label, y_label] + threshold   (reference models/utils/fuser.py:16-17,
//            models/utils/criterion.py:83-85, models/xmask3d.py:356-358).  The [K,H,W] masks are
//            read exactly once, coalesced, and transposed into per-pixel membership words
//            [H*W, words]; each visible point then fetches its words with one small gather.
//   pool     sum[k,:] = sum of feat[i,:] over points i in mask k, cnt[k]  (criterion.py:148-157,
//            xmask3d.py:362-367).  The dominant HBM stream of the whole path (n x C float32).
//            One persistent CTA per SM owns a contiguous run of points and ALL k accumulator rows
//            of its channel slice in shared memory; thread t owns channels [VEC*t, VEC*t+VEC) of
//            every accumulator row, so there are no atomics, no inter-thread hazards and no
//            __syncthreads in the stream loop.  Feature rows are streamed with 16-byte
//            L1-bypassing loads through a two-half register ring (>= 8 rows in flight per
//            thread).  Per-(CTA, segment) partials are combined in a fixed order by a second
//            kernel (deterministic, no float atomics).
//   scatter  out[i,:] = (sum over masks m containing i, ascending m, of emb[m,:]) / count_i
//            (fuser.py:22-34; xmask3d.py:441-455) — float32 op order of the reference, bit-exact.
#include "common.cuh"

namespace xm3d {

constexpr int MAX_WORDS = 8;                  // k <= 256 masks per segment
constexpr int POOL_SMEM_MAX = 224 * 1024;     // accumulator bytes per CTA
constexpr int POOL_HALF = 8;                  // rows per half of the register ring
constexpr int POOL_MAX_THREADS = 256;         // <= 256 threads per CTA keeps 255 registers per thread

__device__ __forceinline__ bool mask_hit(float x, int thr_mode) {
    if (thr_mode == XM3D_THR_GE_HALF) return x >= 0.5f;
    const float sg = __fdiv_rn(1.0f, __fadd_rn(1.0f, expf(-x)));       // torch.sigmoid, float32
    return thr_mode == XM3D_THR_SIGMOID_GE_HALF ? (sg >= 0.5f) : (sg > 0.5f);
}

// ---- gather, step 1: [k,h,w] masks -> per-pixel membership words ---------------------------
template <typename T>
__global__ void __launch_bounds__(256)
pixel_bits_kernel(const T *__restrict__ masks, int thr_mode, int k, int hw, int words, int vec_ok,
                  uint32_t *__restrict__ pixbits) {
    const int s = blockIdx.y;
    const int p0 = (blockIdx.x * blockDim.x + threadIdx.x) * 4;
    if (p0 >= hw) return;
    const T *base = masks + (size_t)s * k * hw;
    uint32_t *out = pixbits + ((size_t)s * hw + p0) * words;
    const int np = min(4, hw - p0);
    for (int wd = 0; wd < words; ++wd) {
        uint32_t acc[4] = {0u, 0u, 0u, 0u};
        const int m_end = min(k, (wd + 1) * 32);
        for (int m = wd * 32; m < m_end; ++m) {
            const T *pl = base + (size_t)m * hw + p0;
            float x[4];
            if (vec_ok && np == 4) {
                if (sizeof(T) == 1) {
                    const uchar4 v = __ldg(reinterpret_cast<const uchar4 *>(pl));
                    x[0] = v.x; x[1] = v.y; x[2] = v.z; x[3] = v.w;
                } else {
                    const float4 v = __ldg(reinterpret_cast<const float4 *>(pl));
                    x[0] = v.x; x[1] = v.y; x[2] = v.z; x[3] = v.w;
                }
            } else {
#pragma unroll
                for (int j = 0; j < 4; ++j) x[j] = (j < np) ? (float)__ldg(pl + j) : 0.f;
            }
#pragma unroll
            for (int j = 0; j < 4; ++j)
                if (mask_hit(x[j], thr_mode)) acc[j] |= 1u << (m & 31);
        }
        for (int j = 0; j < np; ++j) out[(size_t)j * words + wd] = acc[j];
    }
}

// ---- gather, step 2: per-point words (+ optional per-mask counts) --------------------------
__global__ void __launch_bounds__(256)
point_bits_kernel(const uint32_t *__restrict__ pixbits, const int32_t *__restrict__ rowcol,
                  const int64_t *__restrict__ seg_off, int n_seg, int64_t cap, int k, int h, int w, int words,
                  uint32_t *__restrict__ member, int32_t *__restrict__ counts) {
    int64_t total = seg_off[n_seg];
    if (total > cap) total = 0;
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const bool valid = i < total;
    const unsigned act = __ballot_sync(0xffffffffu, valid);
    if (!valid) return;
    const int s = seg_of(seg_off, n_seg, i);
    const int2 rc = __ldg(reinterpret_cast<const int2 *>(rowcol) + i);
    const bool inb = rc.x >= 0 && rc.x < h && rc.y >= 0 && rc.y < w;
    const uint32_t *src = pixbits + ((size_t)s * h * w + (size_t)rc.x * w + rc.y) * words;
    const unsigned sameseg = __match_any_sync(act, s);
    for (int wd = 0; wd < words; ++wd) {
        const uint32_t b = inb ? __ldg(src + wd) : 0u;
        member[i * words + wd] = b;
        if (counts) {
            if (sameseg == act) {          // warp inside one segment: one atomic per mask per warp
                const int m_end = min(32, k - wd * 32);
                for (int m = 0; m < m_end; ++m) {
                    const unsigned vote = __ballot_sync(act, (b >> m) & 1u);
                    if (vote && lane_id() == (__ffs(act) - 1)) atomicAdd(&counts[(size_t)s * k + wd * 32 + m], __popc(vote));
                }
            } else {
                uint32_t t = b;
                while (t) {
                    const int m = __ffs(t) - 1;
                    t &= t - 1;
                    atomicAdd(&counts[(size_t)s * k + wd * 32 + m], 1);
                }
            }
        }
    }
}

// ---- pooling ------------------------------------------------------------------------------
template <int VEC> struct VecT;
template <> struct VecT<4> { using type = float4; };
template <> struct VecT<2> { using type = float2; };
template <> struct VecT<1> { using type = float; };

template <int VEC>
__device__ __forceinline__ typename VecT<VEC>::type ld_feat(const float *p);
template <>
__device__ __forceinline__ float4 ld_feat<4>(const float *p) { return ldg_stream4(p); }
template <>
__device__ __forceinline__ float2 ld_feat<2>(const float *p) {
    float2 r;
    asm volatile("ld.global.nc.L1::no_allocate.v2.f32 {%0,%1}, [%2];" : "=f"(r.x), "=f"(r.y) : "l"(p));
    return r;
}
template <>
__device__ __forceinline__ float ld_feat<1>(const float *p) {
    float r;
    asm volatile("ld.global.nc.L1::no_allocate.f32 %0, [%1];" : "=f"(r) : "l"(p));
    return r;
}
__device__ __forceinline__ void vadd(float4 &a, const float4 &b) { a.x += b.x; a.y += b.y; a.z += b.z; a.w += b.w; }
__device__ __forceinline__ void vadd(float2 &a, const float2 &b) { a.x += b.x; a.y += b.y; }
__device__ __forceinline__ void vadd(float &a, const float &b) { a += b; }
__device__ __forceinline__ void vzero(float4 &a) { a = make_float4(0.f, 0.f, 0.f, 0.f); }
__device__ __forceinline__ void vzero(float2 &a) { a = make_float2(0.f, 0.f); }
__device__ __forceinline__ void vzero(float &a) { a = 0.f; }

__host__ __device__ inline int64_t pool_chunk_start(int64_t total, int g, int G) {
    return total * (int64_t)g / G;      // total < 2^31, G < 2^12
}

struct PoolParams {
    const float *feat;
    int c;
    const int32_t *row_index;
    const uint32_t *member;      // [total, words] or null
    const int32_t *label;        // [total] or null
    const int64_t *seg_off;
    int n_seg, k, words;
    int64_t cap;
    float *part_sum;             // [n_seg + G, k, c]
    int32_t *part_cnt;           // [n_seg + G, k]
    int wc;                      // channels per CTA slice
};

template <int VEC, int WORDS, bool LABEL>
__global__ void __launch_bounds__(POOL_MAX_THREADS, 1) pool_kernel(const PoolParams P) {
    using V = typename VecT<VEC>::type;
    extern __shared__ __align__(16) unsigned char smem_acc[];
    V *acc = reinterpret_cast<V *>(smem_acc);           // [k][blockDim.x]
    const int tid = threadIdx.x, T = blockDim.x;
    const int G = gridDim.x, g = blockIdx.x;
    const int ch = blockIdx.y * P.wc + tid * VEC;        // first channel of this thread
    const bool active = ch < P.c && tid * VEC < P.wc;
    int64_t total = P.seg_off[P.n_seg];
    if (total > P.cap) total = 0;
    int64_t p = pool_chunk_start(total, g, G);
    const int64_t p_end = pool_chunk_start(total, g + 1, G);
    if (p >= p_end) return;
    int s = seg_of(P.seg_off, P.n_seg, p);
    const bool count_warp = (blockIdx.y == 0) && (tid < 32);
    const int k = P.k;

    while (p < p_end) {
        const int64_t s_end = P.seg_off[s + 1];
        const int64_t e = s_end < p_end ? s_end : p_end;
        if (e > p) {
            // ---- one (CTA, segment) piece: zero own accumulator columns
            for (int m = 0; m < k; ++m) vzero(acc[m * T + tid]);
            int cnt_reg[MAX_WORDS];
#pragma unroll
            for (int j = 0; j < MAX_WORDS; ++j) cnt_reg[j] = 0;

            V bufA[POOL_HALF], bufB[POOL_HALF];
            uint32_t bitA[POOL_HALF][WORDS], bitB[POOL_HALF][WORDS];

            auto load_half = [&](V (&buf)[POOL_HALF], uint32_t (&bits)[POOL_HALF][WORDS], int64_t base) {
#pragma unroll
                for (int j = 0; j < POOL_HALF; ++j) {
                    const int64_t i = base + j;
                    if (i < e) {
                        if (LABEL) {
                            bits[j][0] = (uint32_t)__ldg(P.label + i);
                        } else {
#pragma unroll
                            for (int w = 0; w < WORDS; ++w)
                                bits[j][w] = (w < P.words) ? __ldg(P.member + i * P.words + w) : 0u;
                        }
                        const int64_t row = P.row_index ? (int64_t)__ldg(P.row_index + i) : i;
                        if (active) buf[j] = ld_feat<VEC>(P.feat + row * P.c + ch);
                    }
                }
            };
            auto consume_half = [&](V (&buf)[POOL_HALF], uint32_t (&bits)[POOL_HALF][WORDS], int64_t base) {
#pragma unroll
                for (int j = 0; j < POOL_HALF; ++j) {
                    if (base + j < e) {
                        if (LABEL) {
                            const int m = (int)bits[j][0];
                            if (m >= 0 && m < k) {
                                if (active) { V a = acc[m * T + tid]; vadd(a, buf[j]); acc[m * T + tid] = a; }
                                if (count_warp && (m & 31) == tid) {
#pragma unroll
                                    for (int w = 0; w < MAX_WORDS; ++w) cnt_reg[w] += ((m >> 5) == w) ? 1 : 0;
                                }
                            }
                        } else {
#pragma unroll
                            for (int w = 0; w < WORDS; ++w) {
                                uint32_t b = bits[j][w];
                                if (count_warp) cnt_reg[w] += (b >> tid) & 1u;
                                while (b) {
                                    const int m = w * 32 + __ffs(b) - 1;
                                    b &= b - 1;
                                    if (active) { V a = acc[m * T + tid]; vadd(a, buf[j]); acc[m * T + tid] = a; }
                                }
                            }
                        }
                    }
                }
            };

            load_half(bufA, bitA, p);
            for (int64_t base = p; base < e; base += 2 * POOL_HALF) {
                load_half(bufB, bitB, base + POOL_HALF);
                consume_half(bufA, bitA, base);
                load_half(bufA, bitA, base + 2 * POOL_HALF);
                consume_half(bufB, bitB, base + POOL_HALF);
            }

            // ---- flush this piece
            const size_t slot = (size_t)s + g;
            if (active) {
                float *dst = P.part_sum + (slot * k) * P.c + ch;
                for (int m = 0; m < k; ++m)
                    *reinterpret_cast<V *>(dst + (size_t)m * P.c) = acc[m * T + tid];
            }
            if (count_warp) {
#pragma unroll
                for (int w = 0; w < MAX_WORDS; ++w) {
                    const int m = w * 32 + tid;
                    if (m < k) P.part_cnt[slot * k + m] = cnt_reg[w];
                }
            }
        }
        p = e;
        ++s;
    }
}

__global__ void __launch_bounds__(256)
pool_combine_kernel(const float *__restrict__ part_sum, const int32_t *__restrict__ part_cnt,
                    const int64_t *__restrict__ seg_off, int n_seg, int64_t cap, int k, int c, int G,
                    float *__restrict__ sum, int32_t *__restrict__ cnt, float *__restrict__ mean) {
    const int s = blockIdx.y;
    const int e = blockIdx.x * blockDim.x + threadIdx.x;      // element of [k, c]
    if (e >= k * c) return;
    int64_t total = seg_off[n_seg];
    if (total > cap) total = 0;
    const int64_t a = seg_off[s], b = total ? seg_off[s + 1] : a;
    float acc = 0.f;
    int n = 0;
    const int m = e / c;
    if (b > a) {
        // CTAs whose chunk intersects [a, b): chunk g = [total*g/G, total*(g+1)/G)
        int g_lo = (int)(a * (int64_t)G / total);
        while (g_lo > 0 && pool_chunk_start(total, g_lo, G) > a) --g_lo;
        while (g_lo + 1 < G && pool_chunk_start(total, g_lo + 1, G) <= a) ++g_lo;
        for (int g = g_lo; g < G; ++g) {
            const int64_t cs = pool_chunk_start(total, g, G), ce = pool_chunk_start(total, g + 1, G);
            if (cs >= b) break;
            const int64_t lo = cs > a ? cs : a, hi = ce < b ? ce : b;
            if (hi <= lo) continue;
            const size_t slot = (size_t)s + g;
            acc += part_sum[slot * k * c + e];                  // fixed ascending-g order
            n += part_cnt[slot * k + m];
        }
    }
    sum[(size_t)s * k * c + e] = acc;
    if (mean) mean[(size_t)s * k * c + e] = n > 0 ? __fdiv_rn(acc, (float)n) : 0.f;
    if (cnt && (e % c) == 0) cnt[(size_t)s * k + m] = n;
}

// ---- scatter ------------------------------------------------------------------------------
template <int VEC>
__global__ void __launch_bounds__(1024)
scatter_kernel(const uint32_t *__restrict__ member, const int32_t *__restrict__ label,
               const int64_t *__restrict__ seg_off, int n_seg, int64_t cap, int k, int words,
               const float *__restrict__ emb, int c, float *__restrict__ out, float *__restrict__ counter,
               int pts_per_cta) {
    using V = typename VecT<VEC>::type;
    int64_t total = seg_off[n_seg];
    if (total > cap) total = 0;
    const int ch = threadIdx.x * VEC;
    const bool active = ch < c;
    const int64_t p0 = (int64_t)blockIdx.x * pts_per_cta;
    const int64_t p1 = min(total, p0 + pts_per_cta);
    if (p0 >= p1) return;
    int s = seg_of(seg_off, n_seg, p0);
    int64_t s_end = seg_off[s + 1];
    for (int64_t i = p0; i < p1; ++i) {
        while (i >= s_end) { ++s; s_end = seg_off[s + 1]; }
        const float *eb = emb + (size_t)s * k * c + ch;
        V acc; vzero(acc);
        float cntf = 0.f;
        if (label) {
            const int m = __ldg(label + i);
            if (m >= 0 && m < k) {
                if (active) vadd(acc, *reinterpret_cast<const V *>(eb + (size_t)m * c));
                cntf = 1.f;
            }
        } else {
            for (int w = 0; w < words; ++w) {
                uint32_t b = __ldg(member + i * words + w);
                while (b) {
                    const int m = w * 32 + __ffs(b) - 1;
                    b &= b - 1;
                    if (active) vadd(acc, *reinterpret_cast<const V *>(eb + (size_t)m * c));   // feat[mask] += emb
                    cntf = __fadd_rn(cntf, 1.f);                                                 // counter[mask] += 1
                }
            }
        }
        if (cntf == 0.f) cntf = 1e-5f;                        // counter[counter == 0] = 1e-5
        if (active) {
            float *a = reinterpret_cast<float *>(&acc);
#pragma unroll
            for (int j = 0; j < VEC; ++j) a[j] = __fdiv_rn(a[j], cntf);
            *reinterpret_cast<V *>(out + i * c + ch) = acc;
        }
        if (counter && threadIdx.x == 0) counter[i] = cntf;
    }
}

static int words_for(int k) { return (k + 31) / 32; }

struct PoolPlan {
    int vec, wc, threads, slices, G;
    size_t smem;
};

static bool plan_pool(int c, int k, bool aligned16, PoolPlan *pl) {
    // channels per CTA slice: the whole row when the k accumulator rows fit in shared memory and
    // one thread per VEC channels stays within POOL_MAX_THREADS, else even slices of whole warps
    int vec = (c % 4 == 0 && aligned16) ? 4 : 1;
    const int gran = 32 * vec;
    int wc = c < POOL_MAX_THREADS * vec ? c : POOL_MAX_THREADS * vec;
    while ((size_t)k * ((wc + vec - 1) / vec * vec) * 4 > (size_t)POOL_SMEM_MAX) {
        const int next = ((wc - 1) / gran) * gran;
        if (next <= 0) return false;
        wc = next;
    }
    if (wc < c) {
        const int slices = (c + wc - 1) / wc;
        const int even = ((c + slices - 1) / slices + gran - 1) / gran * gran;
        if (even <= wc) wc = even;
    }
    // few threads per CTA starve the memory pipeline: narrow the per-thread vector instead
    if (vec == 4 && wc / 4 < 128 && wc % 2 == 0) vec = 2;
    if (vec == 2 && wc / 2 < 128) vec = 1;
    const int threads = ((wc + vec - 1) / vec + 31) / 32 * 32;
    if (threads > POOL_MAX_THREADS) return false;
    pl->vec = vec; pl->wc = wc; pl->threads = threads;
    pl->slices = (c + wc - 1) / wc;
    pl->smem = (size_t)k * threads * vec * 4;
    if (pl->smem > (size_t)POOL_SMEM_MAX) return false;
    const int per_sm = (int)(POOL_SMEM_MAX / (pl->smem ? pl->smem : 1));
    const int ctas = sm_count() * (per_sm < 1 ? 1 : (per_sm > 4 ? 4 : per_sm));
    pl->G = ctas / pl->slices;
    if (pl->G < 1) pl->G = 1;
    return true;
}

template <int VEC, int WORDS, bool LABEL>
static void launch_pool(const PoolParams &P, const PoolPlan &pl, cudaStream_t stream) {
    auto kern = pool_kernel<VEC, WORDS, LABEL>;
    static bool attr_set = false;
    if (!attr_set) {
        cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, POOL_SMEM_MAX);
        attr_set = true;
    }
    kern<<<dim3(pl.G, pl.slices), pl.threads, pl.smem, stream>>>(P); count_launches(1);
}

template <int VEC>
static int dispatch_pool(const PoolParams &P, const PoolPlan &pl, cudaStream_t stream) {
    if (P.label) { launch_pool<VEC, 1, true>(P, pl, stream); return XM3D_OK; }
    switch (P.words) {
        case 1: launch_pool<VEC, 1, false>(P, pl, stream); break;
        case 2: launch_pool<VEC, 2, false>(P, pl, stream); break;
        case 3: launch_pool<VEC, 3, false>(P, pl, stream); break;
        case 4: launch_pool<VEC, 4, false>(P, pl, stream); break;
        case 5: case 6: case 7: case 8: launch_pool<VEC, 8, false>(P, pl, stream); break;
        default: return XM3D_ERR_UNSUPPORTED;
    }
    return XM3D_OK;
}

}  // namespace xm3d

using namespace xm3d;

extern "C" size_t xm3d_gather_ws_bytes(int32_t n_seg, int32_t k, int32_t h, int32_t w) {
    return align_up((size_t)n_seg * h * w * words_for(k) * 4, 256) + 256;
}

extern "C" int32_t xm3d_mask_words(int32_t k) { return words_for(k); }

extern "C" int xm3d_gather_masks_batch(const void *masks, int32_t mask_kind, int32_t thr_mode, int32_t n_seg,
                                       int32_t k, int32_t h, int32_t w, const int32_t *rowcol,
                                       const int64_t *seg_off, int64_t cap, uint32_t *member, int32_t *counts,
                                       void *ws, size_t ws_bytes, xm3d_stream_t stream_) {
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    XM3D_REQUIRE(n_seg > 0 && k > 0 && h > 0 && w > 0 && cap >= 0, "bad sizes");
    XM3D_REQUIRE(k <= 32 * MAX_WORDS, "at most 256 masks per segment");
    XM3D_REQUIRE(masks && rowcol && seg_off && member && ws, "null pointer");
    XM3D_REQUIRE(mask_kind == XM3D_MASK_U8 || mask_kind == XM3D_MASK_F32, "bad mask_kind");
    XM3D_REQUIRE(thr_mode >= 0 && thr_mode <= 2, "bad thr_mode");
    if (ws_bytes < xm3d_gather_ws_bytes(n_seg, k, h, w)) {
        set_error("xm3d_gather_masks_batch: workspace too small");
        return XM3D_ERR_WORKSPACE;
    }
    const int words = words_for(k), hw = h * w;
    uint32_t *pixbits = static_cast<uint32_t *>(ws);
    const int esz = mask_kind == XM3D_MASK_U8 ? 1 : 4;
    const int vec_ok = (hw % 4 == 0) && (reinterpret_cast<uintptr_t>(masks) % (4 * esz) == 0);
    dim3 grid((hw / 4 + 256) / 256, n_seg);
    if (mask_kind == XM3D_MASK_U8) {
        pixel_bits_kernel<unsigned char><<<grid, 256, 0, stream>>>(static_cast<const unsigned char *>(masks), thr_mode,
                                                                    k, hw, words, vec_ok, pixbits); count_launches(1); }
    else {
        pixel_bits_kernel<float><<<grid, 256, 0, stream>>>(static_cast<const float *>(masks), thr_mode, k, hw, words,
                                                           vec_ok, pixbits); count_launches(1); }
    if (counts) cudaMemsetAsync(counts, 0, sizeof(int32_t) * (size_t)n_seg * k, stream);
    if (cap > 0) {
        point_bits_kernel<<<(unsigned)((cap + 255) / 256), 256, 0, stream>>>(pixbits, rowcol, seg_off, n_seg, cap, k, h,
                                                                             w, words, member, counts); count_launches(1); }
    return check_launch("xm3d_gather_masks_batch");
}

extern "C" size_t xm3d_pool_ws_bytes(int32_t n_seg, int32_t k, int32_t c) {
    const size_t slots = (size_t)n_seg + (size_t)sm_count() * 4 + 1;
    return align_up(slots * k * c * 4, 256) + align_up(slots * k * 4, 256) + 256;
}

extern "C" int xm3d_pool_batch(const float *feat, int32_t c, const int32_t *row_index, const uint32_t *member,
                               const int32_t *label, int32_t n_seg, int32_t k, const int64_t *seg_off, int64_t cap,
                               float *sum, int32_t *cnt, float *mean, void *ws, size_t ws_bytes, int32_t *status,
                               xm3d_stream_t stream_) {
    (void)status;
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    XM3D_REQUIRE(n_seg > 0 && k > 0 && c > 0 && cap >= 0, "bad sizes");
    XM3D_REQUIRE(k <= 32 * MAX_WORDS, "at most 256 masks per segment");
    XM3D_REQUIRE(feat && seg_off && sum && ws, "null pointer");
    XM3D_REQUIRE((member != nullptr) != (label != nullptr), "exactly one of member / label");
    if (ws_bytes < xm3d_pool_ws_bytes(n_seg, k, c)) {
        set_error("xm3d_pool_batch: workspace too small");
        return XM3D_ERR_WORKSPACE;
    }
    PoolPlan pl;
    const bool aligned16 = reinterpret_cast<uintptr_t>(feat) % 16 == 0;
    if (!plan_pool(c, k, aligned16, &pl)) {
        set_error("xm3d_pool_batch: k=%d accumulator rows do not fit in shared memory", k);
        return XM3D_ERR_UNSUPPORTED;
    }
    const size_t slots = (size_t)n_seg + (size_t)sm_count() * 4 + 1;
    Carver cv(ws);
    float *part_sum = cv.take<float>(slots * k * c);
    int32_t *part_cnt = cv.take<int32_t>(slots * k);

    PoolParams P;
    P.feat = feat; P.c = c; P.row_index = row_index; P.member = member; P.label = label; P.seg_off = seg_off;
    P.n_seg = n_seg; P.k = k; P.words = words_for(k); P.cap = cap; P.part_sum = part_sum; P.part_cnt = part_cnt;
    P.wc = pl.wc;
    int rc = XM3D_OK;
    if (pl.vec == 4) rc = dispatch_pool<4>(P, pl, stream);
    else if (pl.vec == 2) rc = dispatch_pool<2>(P, pl, stream);
    else rc = dispatch_pool<1>(P, pl, stream);
    if (rc != XM3D_OK) { set_error("xm3d_pool_batch: unsupported mask count"); return rc; }
    dim3 cgrid((unsigned)(((size_t)k * c + 255) / 256), n_seg);
    pool_combine_kernel<<<cgrid, 256, 0, stream>>>(part_sum, part_cnt, seg_off, n_seg, cap, k, c, pl.G, sum, cnt, mean); count_launches(1);
    return check_launch("xm3d_pool_batch");
}

extern "C" int xm3d_scatter_batch(const uint32_t *member, const int32_t *label, int32_t n_seg, int32_t k,
                                  const int64_t *seg_off, int64_t cap, const float *emb, int32_t c, float *out,
                                  float *counter, xm3d_stream_t stream_) {
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    XM3D_REQUIRE(n_seg > 0 && k > 0 && c > 0 && cap >= 0, "bad sizes");
    XM3D_REQUIRE(k <= 32 * MAX_WORDS, "at most 256 masks per segment");
    XM3D_REQUIRE(seg_off && emb && out, "null pointer");
    XM3D_REQUIRE((member != nullptr) != (label != nullptr), "exactly one of member / label");
    if (cap == 0) return XM3D_OK;
    const bool v4 = c % 4 == 0 && c / 4 <= 1024 && reinterpret_cast<uintptr_t>(emb) % 16 == 0 &&
                    reinterpret_cast<uintptr_t>(out) % 16 == 0;
    XM3D_REQUIRE(v4 || c <= 1024, "feature width not supported");
    const int pts = 64;
    const unsigned blocks = (unsigned)((cap + pts - 1) / pts);
    if (v4) {
        const int threads = (c / 4 + 31) / 32 * 32;
        scatter_kernel<4><<<blocks, threads, 0, stream>>>(member, label, seg_off, n_seg, cap, k, words_for(k), emb, c,
                                                          out, counter, pts); count_launches(1);
    } else {
        const int threads = (c + 31) / 32 * 32;
        scatter_kernel<1><<<blocks, threads, 0, stream>>>(member, label, seg_off, n_seg, cap, k, words_for(k), emb, c,
                                                          out, counter, pts); count_launches(1);
    }
    return check_launch("xm3d_scatter_batch");
}
