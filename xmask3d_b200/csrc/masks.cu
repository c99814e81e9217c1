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
#include "vec.cuh"

namespace xm3d {


__device__ __forceinline__ bool mask_hit(float x, int thr_mode) {
    if (thr_mode == XM3D_THR_GE_HALF) return x >= 0.5f;
    // sigmoid(x) - 0.5 has the sign of x; only within 1e-6 of zero can float32 rounding of exp / add / div decide
    // otherwise (2 instructions instead of ~20 per (pixel, mask): the float32 pass becomes memory bound)
    if (fabsf(x) > 1e-6f) return x > 0.f;
    const float sg = __fdiv_rn(1.0f, __fadd_rn(1.0f, expf(-x)));       // torch.sigmoid, float32
    return thr_mode == XM3D_THR_SIGMOID_GE_HALF ? (sg >= 0.5f) : (sg > 0.5f);
}

// ---- gather, step 1: [k,h,w] masks -> per-pixel membership words ---------------------------
// One thread owns PIX = 16 / sizeof(T) consecutive pixels and walks the k mask planes with
// 16-byte loads (four planes in flight), so every plane is read exactly once, coalesced.
template <typename T>
__device__ __forceinline__ bool pix_hit(T v, int thr_mode) {
    if (sizeof(T) == 1 && thr_mode == XM3D_THR_GE_HALF) return v != 0;      // bool / uint8 masks: 1 >= 0.5
    return mask_hit((float)v, thr_mode);
}

template <typename T>
__global__ void __launch_bounds__(256)
pixel_bits_kernel(const T *__restrict__ masks, int thr_mode, int k, int hw, int words, int vec_ok,
                  uint32_t *__restrict__ pixbits) {
    constexpr int PIX = 16 / sizeof(T);
    const int s = blockIdx.y;
    const int p0 = (blockIdx.x * blockDim.x + threadIdx.x) * PIX;
    if (p0 >= hw) return;
    const T *base = masks + (size_t)s * k * hw + p0;
    uint32_t *out = pixbits + (size_t)s * words * hw + p0;      // layout [segment][word][pixel]
    const int np = min(PIX, hw - p0);
    const bool vec = vec_ok && np == PIX;
    for (int wd = 0; wd < words; ++wd) {
        uint32_t acc[PIX];
#pragma unroll
        for (int j = 0; j < PIX; ++j) acc[j] = 0u;
        const int m_end = min(k, (wd + 1) * 32);
        for (int m0 = wd * 32; m0 < m_end; m0 += 4) {
            union { uint4 q; T e[PIX]; } v[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                if (m0 + u < m_end) {
                    const T *pl = base + (size_t)(m0 + u) * hw;
                    if (vec) {
                        v[u].q = __ldg(reinterpret_cast<const uint4 *>(pl));
                    } else {
#pragma unroll
                        for (int j = 0; j < PIX; ++j) v[u].e[j] = (j < np) ? __ldg(pl + j) : T(0);
                    }
                }
            }
#pragma unroll
            for (int u = 0; u < 4; ++u)
                if (m0 + u < m_end) {
#pragma unroll
                    for (int j = 0; j < PIX; ++j)
                        if (pix_hit<T>(v[u].e[j], thr_mode)) acc[j] |= 1u << ((m0 + u) & 31);
                }
        }
        uint32_t *o = out + (size_t)wd * hw;
        if (vec) {
#pragma unroll
            for (int j = 0; j < PIX; j += 4)
                *reinterpret_cast<uint4 *>(o + j) = make_uint4(acc[j], acc[j + 1], acc[j + 2], acc[j + 3]);
        } else {
            for (int j = 0; j < np; ++j) o[j] = acc[j];
        }
    }
}

// bool / uint8 masks thresholded at >= 0.5 (the partition masks of mask_mapper): byte-SIMD path.
// One thread owns 16 consecutive pixels (4 quads of 4 bytes).  For mask m the four bytes of a quad
// become 0xFF / 0x00 with one vcmpne4, are AND-ed with bit (m & 7) replicated in every byte and
// OR-ed into the quad's accumulator of byte group (m & 31) >> 3 — three integer instructions per
// four pixels and mask; the per-pixel words are assembled once per 32 masks.
__global__ void __launch_bounds__(256)
pixel_bits_u8_kernel(const unsigned char *__restrict__ masks, int k, int hw, int words,
                     uint32_t *__restrict__ pixbits) {
    const int s = blockIdx.y;
    const int p0 = (blockIdx.x * blockDim.x + threadIdx.x) * 16;
    if (p0 >= hw) return;                                   // hw % 16 == 0 on this path
    const unsigned char *base = masks + (size_t)s * k * hw + p0;
    uint32_t *out = pixbits + (size_t)s * words * hw + p0;
    {
        const int wd = blockIdx.z;                          // one word of 32 masks per CTA: short CTAs, short tail
        uint32_t acc[4][4];                                 // [byte group][quad]
#pragma unroll
        for (int g = 0; g < 4; ++g)
#pragma unroll
            for (int q = 0; q < 4; ++q) acc[g][q] = 0u;
        const int m_beg = wd * 32, m_cnt = min(32, k - m_beg);
#pragma unroll
        for (int mm = 0; mm < 32; mm += 4) {
            if (mm < m_cnt) {
                uint4 v[4];
#pragma unroll
                for (int u = 0; u < 4; ++u)
                    v[u] = (mm + u < m_cnt) ? __ldg(reinterpret_cast<const uint4 *>(base + (size_t)(m_beg + mm + u) * hw))
                                            : make_uint4(0u, 0u, 0u, 0u);
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    const uint32_t sel = 0x01010101u << ((mm + u) & 7);
                    const int g = (mm + u) >> 3;
                    acc[g][0] |= __vcmpne4(v[u].x, 0u) & sel;
                    acc[g][1] |= __vcmpne4(v[u].y, 0u) & sel;
                    acc[g][2] |= __vcmpne4(v[u].z, 0u) & sel;
                    acc[g][3] |= __vcmpne4(v[u].w, 0u) & sel;
                }
            }
        }
        uint32_t *o = out + (size_t)wd * hw;
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            // transpose 4 byte groups x 4 pixels: word of pixel j = byte j of every group
            const uint32_t a0 = acc[0][q], a1 = acc[1][q], a2 = acc[2][q], a3 = acc[3][q];
            const uint32_t t01l = __byte_perm(a0, a1, 0x5140), t01h = __byte_perm(a0, a1, 0x7362);
            const uint32_t t23l = __byte_perm(a2, a3, 0x5140), t23h = __byte_perm(a2, a3, 0x7362);
            const uint4 w = make_uint4(__byte_perm(t01l, t23l, 0x5410), __byte_perm(t01l, t23l, 0x7632),
                                       __byte_perm(t01h, t23h, 0x5410), __byte_perm(t01h, t23h, 0x7632));
            *reinterpret_cast<uint4 *>(o + 4 * q) = w;
        }
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
    const size_t hw = (size_t)h * w;
    const uint32_t *src = pixbits + (size_t)s * words * hw + (size_t)rc.x * w + rc.y;
    const unsigned sameseg = __match_any_sync(act, s);
    for (int wd = 0; wd < words; ++wd) {
        const uint32_t b = inb ? __ldg(src + (size_t)wd * hw) : 0u;
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


}  // namespace xm3d

using namespace xm3d;

extern "C" size_t xm3d_gather_ws_bytes(int32_t n_seg, int32_t k, int32_t h, int32_t w) {
    return align_up((size_t)n_seg * h * w * words_for(k) * 4, 256) + 256;
}

extern "C" int32_t xm3d_mask_words(int32_t k) { return words_for(k); }

// [n_seg, k, h, w] masks -> per-pixel membership words [n_seg, words, h*w]
static int launch_pixel_bits(const void *masks, int mask_kind, int thr_mode, int n_seg, int k, int h, int w,
                             uint32_t *pixbits, cudaStream_t stream) {
    const int words = words_for(k), hw = h * w;
    const int esz = mask_kind == XM3D_MASK_U8 ? 1 : 4;
    const int pix = 16 / esz;
    const int vec_ok = (hw % pix == 0) && (reinterpret_cast<uintptr_t>(masks) % 16 == 0);
    dim3 grid(((hw + pix - 1) / pix + 255) / 256, n_seg);
    if (mask_kind == XM3D_MASK_U8 && thr_mode == XM3D_THR_GE_HALF && vec_ok) {
        dim3 grid8(grid.x, n_seg, words);
        pixel_bits_u8_kernel<<<grid8, 256, 0, stream>>>(static_cast<const unsigned char *>(masks), k, hw, words, pixbits);
    } else if (mask_kind == XM3D_MASK_U8) {
        pixel_bits_kernel<unsigned char><<<grid, 256, 0, stream>>>(static_cast<const unsigned char *>(masks), thr_mode,
                                                                    k, hw, words, vec_ok, pixbits);
    } else {
        pixel_bits_kernel<float><<<grid, 256, 0, stream>>>(static_cast<const float *>(masks), thr_mode, k, hw, words,
                                                           vec_ok, pixbits);
    }
    count_launches(1);
    return XM3D_OK;
}

extern "C" int xm3d_pixel_bits_batch(const void *masks, int32_t mask_kind, int32_t thr_mode, int32_t n_seg, int32_t k,
                                     int32_t h, int32_t w, uint32_t *pixbits, xm3d_stream_t stream_) {
    XM3D_REQUIRE(n_seg > 0 && k > 0 && h > 0 && w > 0, "bad sizes");
    XM3D_REQUIRE(k <= 32 * MAX_WORDS, "at most 256 masks per segment");
    XM3D_REQUIRE(masks && pixbits, "null pointer");
    XM3D_REQUIRE(mask_kind == XM3D_MASK_U8 || mask_kind == XM3D_MASK_F32, "bad mask_kind");
    XM3D_REQUIRE(thr_mode >= 0 && thr_mode <= 2, "bad thr_mode");
    launch_pixel_bits(masks, mask_kind, thr_mode, n_seg, k, h, w, pixbits, static_cast<cudaStream_t>(stream_));
    return check_launch("xm3d_pixel_bits_batch");
}

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
    const int words = words_for(k);
    uint32_t *pixbits = static_cast<uint32_t *>(ws);
    launch_pixel_bits(masks, mask_kind, thr_mode, n_seg, k, h, w, pixbits, stream);
    if (counts) cudaMemsetAsync(counts, 0, sizeof(int32_t) * (size_t)n_seg * k, stream);
    if (cap > 0) {
        point_bits_kernel<<<(unsigned)((cap + 255) / 256), 256, 0, stream>>>(pixbits, rowcol, seg_off, n_seg, cap, k, h,
                                                                             w, words, member, counts); count_launches(1); }
    return check_launch("xm3d_gather_masks_batch");
}

extern "C" int xm3d_point_bits_batch(const uint32_t *pixbits, int32_t n_seg, int32_t k, int32_t h, int32_t w,
                                     const int32_t *rowcol, const int64_t *seg_off, int64_t cap, uint32_t *member,
                                     int32_t *counts, xm3d_stream_t stream_) {
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    XM3D_REQUIRE(n_seg > 0 && k > 0 && h > 0 && w > 0 && cap >= 0, "bad sizes");
    XM3D_REQUIRE(k <= 32 * MAX_WORDS, "at most 256 masks per segment");
    XM3D_REQUIRE(pixbits && rowcol && seg_off && member, "null pointer");
    if (counts) cudaMemsetAsync(counts, 0, sizeof(int32_t) * (size_t)n_seg * k, stream);
    if (cap > 0) {
        point_bits_kernel<<<(unsigned)((cap + 255) / 256), 256, 0, stream>>>(pixbits, rowcol, seg_off, n_seg, cap, k, h,
                                                                             w, words_for(k), member, counts); count_launches(1); }
    return check_launch("xm3d_point_bits_batch");
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
