// Stage 4 — cosine-similarity logits of mask embeddings against text embeddings.
//
// Replaces XMASK3d.cal_pred_logits (reference models/xmask3d.py:129-143; twin
// models/modeling/meta_arch/odise.py:170-194) and ensemble_logits_with_labels
// (models/modeling/meta_arch/helper.py:72-97):
//   out[r, g]        = reduce_{t in group g} scale * <me_r/|me_r|, te_t/|te_t|>
//   out[r, n_groups] = scale * <me_r/|me_r|, null/|null|>
//
// The (rows x C) . (C x T) contraction runs on the tensor cores as 3xTF32 (hi*hi + lo*hi + hi*lo with the operands
// split into TF32-exact hi / lo parts and fp32 accumulation in tensor memory), which reproduces fp32 dot products to
// ~1e-6 relative so that argmax matches the reference's fp32 matmul.  ONE kernel serves the mask-level logits
// (xm3d_logits: rows = B*K mask embeddings, synonym-group reduction + null column in the epilogue) and the per-point
// logits (xm3d_point_logits): the rows are read from HBM exactly once as raw fp32 tiles by TMA and split in shared
// memory; only the (small) text side goes through a prep kernel.  Round 1 ran the mask-level logits as prep (2 launches
// writing hi / lo planes of the mask embeddings to HBM) + GEMM; that kernel is gone.
#include <cuda.h>
#include <cuda_bf16.h>

#include <stdlib.h>
#include <string.h>

#include "common.cuh"
#include "tc.cuh"

namespace xm3d {

constexpr int LG_BM = 128;          // rows of mask embeddings per CTA  (UMMA M)
constexpr int LG_BK = 32;           // fp32 elements per k-block = one 128-byte swizzle row
constexpr int LG_UMMA_K = 8;        // tf32: 32 bytes per MMA k-step
constexpr int LG_MAX_N = 256;
constexpr int LG_THREADS = 192;

// ---- prep: inverse norms + hi/lo TF32 planes ------------------------------------------------
__global__ void __launch_bounds__(256)
logits_prep_kernel(const float *__restrict__ a, const float *__restrict__ b, int64_t rows_a, int64_t rows_b,
                   int c, float *__restrict__ hi, float *__restrict__ lo, unsigned short *__restrict__ bf, int ld_bf,
                   float *__restrict__ inv_norm) {
    // one warp per row; rows [0, rows_a) come from a, [rows_a, rows_a + rows_b) from b
    const int64_t row = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (row >= rows_a + rows_b) return;
    const float *src = row < rows_a ? a + row * c : b + (row - rows_a) * c;
    const int lane = threadIdx.x & 31;
    float ss = 0.f;
    for (int j = lane; j < c; j += 32) {
        const float x = __ldg(src + j);
        ss = fmaf(x, x, ss);
        const float h = __uint_as_float(__float_as_uint(x) & 0xffffe000u);
        const float l = __uint_as_float(__float_as_uint(__fsub_rn(x, h)) & 0xffffe000u);
        hi[row * c + j] = h;
        lo[row * c + j] = l;
        bf[row * ld_bf + j] = __bfloat16_as_ushort(__float2bfloat16_rn(x));     // partner of the rows' bf16 lo tile
    }
    for (int j = c + lane; j < ld_bf; j += 32) bf[row * ld_bf + j] = 0;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) ss += __shfl_xor_sync(0xffffffffu, ss, o);
    // F.normalize: x / max(||x||_2, 1e-12)
    if (lane == 0) inv_norm[row] = __fdiv_rn(1.0f, fmaxf(sqrtf(ss), 1e-12f));
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *,
                                  const cuuint64_t *, const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn encode_fn() {
    static EncodeTiledFn fn = nullptr;
    if (fn) return fn;
    void *p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess ||
        q != cudaDriverEntryPointSuccess)
        return nullptr;
    fn = reinterpret_cast<EncodeTiledFn>(p);
    return fn;
}

bool make_map_sw128(CUtensorMap *m, const float *base, int64_t rows, int c, int box_cols, int box_rows, bool atom32) {
    EncodeTiledFn fn = encode_fn();
    if (!fn) return false;
    const cuuint64_t dims[2] = {(cuuint64_t)c, (cuuint64_t)rows};
    const cuuint64_t strides[1] = {(cuuint64_t)c * 4};
    const cuuint32_t box[2] = {(cuuint32_t)box_cols, (cuuint32_t)box_rows};
    const cuuint32_t estr[2] = {1, 1};
    return fn(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float *>(base), dims, strides, box, estr,
              CU_TENSOR_MAP_INTERLEAVE_NONE, atom32 ? CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B : CU_TENSOR_MAP_SWIZZLE_128B,
              CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}
// bf16 plane [rows, ld] with 64-byte rows per k-block: 64-byte swizzle
static bool make_map_bf16_sw64(CUtensorMap *m, const unsigned short *base, int64_t rows, int ld, int box_rows) {
    EncodeTiledFn fn = encode_fn();
    if (!fn) return false;
    const cuuint64_t dims[2] = {(cuuint64_t)ld, (cuuint64_t)rows};
    const cuuint64_t strides[1] = {(cuuint64_t)ld * 2};
    const cuuint32_t box[2] = {(cuuint32_t)LG_BK, (cuuint32_t)box_rows};
    const cuuint32_t estr[2] = {1, 1};
    return fn(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<unsigned short *>(base), dims, strides, box, estr,
              CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_64B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
              CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}
static bool make_map(CUtensorMap *m, const float *base, int64_t rows, int c, int box_rows) {
    return make_map_sw128(m, base, rows, c, LG_BK, box_rows);
}

// Plain (unswizzled) row-tile map over a row-major float32 matrix: box = box_cols x box_rows (used by the
// point-major pooling kernel in pool.cu)
bool make_row_tile_map(CUtensorMap *m, const float *base, int64_t rows, int c, int box_cols, int box_rows) {
    EncodeTiledFn fn = encode_fn();
    if (!fn) return false;
    const cuuint64_t dims[2] = {(cuuint64_t)c, (cuuint64_t)rows};
    const cuuint64_t strides[1] = {(cuuint64_t)c * 4};
    const cuuint32_t box[2] = {(cuuint32_t)box_cols, (cuuint32_t)box_rows};
    const cuuint32_t estr[2] = {1, 1};
    return fn(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float *>(base), dims, strides, box, estr,
              CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
              CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

// ============================================================================================
// Per-point logits: logit_scale * (normalize(feature) @ normalize(text).T) for
// every visible point, base / novel blending with the binary head and argmax — the pattern of
// run/infer.py:557, 606-640 (and loss_exact, models/utils/criterion.py:184-207).
//
// The [n, C] float32 features are read from HBM exactly once: TMA brings raw 128 x 32 tiles into
// shared memory; the raw tile IS the tf32 hi operand (the tensor core ignores the 13 low mantissa bits),
// four converter warps write lo = x - hi as a bf16 tile (64-byte rows, 64-byte swizzle; error 2^-20 |x|)
// and accumulate the row norms on the way; one thread issues hi x (text hi + text lo) as kind::tf32 MMAs
// and lo x text as kind::f16 MMAs into TMEM;
// four epilogue warps normalise, scale, blend and take the argmax.  The CTAs are PERSISTENT (two per SM
// when two fit) and the accumulator is double-buffered in tensor memory when there is room: the epilogue
// of a tile runs while the next tile is loaded, split and multiplied (19 classes: 1.60 -> 1.46 ms,
// 200 classes: 4.42 -> 3.29 ms, fused-stream ensemble with 19 classes: 1.96 -> 1.49 ms).
#ifndef XM3D_PL_CONV_WARPS
#define XM3D_PL_CONV_WARPS 4
#endif
constexpr int PL_CONV_WARPS = XM3D_PL_CONV_WARPS;                 // converter warps (the first four also run the epilogue)
constexpr int PL_CONV = PL_CONV_WARPS * 32;
constexpr int PL_CHUNKS = LG_BM * LG_BK * 4 / 16 / PL_CONV;    // 16-byte chunks per converter thread and k-block
constexpr int PL_EPI = 128;                                        // four epilogue warps (one per TMEM lane group)
constexpr int PL_THREADS = 64 + PL_CONV + PL_EPI;
constexpr int PL_MAX_STAGES = 8;

struct PointLogitsParams {
    int64_t rows;
    int c, n_text, bn, stages, tmem_cols;
    int nbuf, acc_cols;          // accumulator buffers in tensor memory (1 or 2) and the columns of one
    int fused;                   // 1: [B_hi; B_lo] is one N = 2 bn operand (A_hi is read once), needs 2 bn <= 256
    float scale;
    const float *inv_norm_b;     // [n_text]
    const float *binary;         // [rows] or null
    const unsigned char *is_base;// [n_text] or null (required with binary)
    // fused-stream ensemble (run/infer.py:568-600): softmax over the classes, then for a point inside final mask
    // `mask_label` the geometric mean with that mask's MaskCLIP class probabilities, base / novel ratios per class
    const int *mask_label;       // [rows] (-1 = in no mask) or null: no ensemble
    const float *mask_probs;     // [n_masks, n_text] LOG of softmax(final_pred_open_logits)
    int n_masks;
    float base_ratio, novel_ratio;
    float *out;                  // [rows, n_text] (or [rows, n_groups] when grouped) or null
    int *argmax;                 // [rows] or null
    // synonym-group reduction of cal_pred_logits (helper.py:72-97): column groups [group_off[g], group_off[g+1]) are
    // reduced (max / mean) to one output column; null: one output column per text row
    int grouped;                 // 1: goff holds the group offsets
    unsigned short goff[LG_MAX_N + 2];   // [n_groups + 1] column offsets, in the kernel parameters (no copy, no workspace)
    int n_groups, ensemble_mean;
};

// MODE bit 0: fused-stream ensemble (mask_label), bit 1: synonym-group reduction (group_off).  Compile-time so that the plain
// blend / argmax epilogue carries none of the other two (T = 200 runs one CTA per SM: nothing hides its epilogue).
template <int MODE>
__global__ void __launch_bounds__(PL_THREADS, 2)
point_logits_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_b_hi,
                    const __grid_constant__ CUtensorMap map_b_lo, const __grid_constant__ CUtensorMap map_b_bf,
                    const PointLogitsParams P) {
    extern __shared__ __align__(1024) unsigned char smem[];
    __shared__ uint64_t s_raw[PL_MAX_STAGES], s_conv[PL_MAX_STAGES], s_empty[PL_MAX_STAGES];
    __shared__ uint64_t s_done[2], s_accfree[2];       // accumulator buffer: MMAs of a tile complete / epilogue has read it
    __shared__ uint64_t s_ssfull[2], s_ssfree[2];      // row norms of a tile: converters -> epilogue
    __shared__ uint32_t s_tmem;
    __shared__ float s_invb[LG_MAX_N];
    __shared__ unsigned char s_base[LG_MAX_N];
    __shared__ float s_ss[2][LG_BM];
    __shared__ short s_gid[LG_MAX_N];           // output column of each GEMM column (grouped mode)
    __shared__ short s_glen[LG_MAX_N];          // > 0 on the last column of a group: the group's length

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int n_tiles = (int)((P.rows + LG_BM - 1) / LG_BM);       // persistent: tiles blockIdx.x, + gridDim.x, ...
    const int nkb = (P.c + LG_BK - 1) / LG_BK;
    const uint32_t a_bytes = LG_BM * LG_BK * 4, b_bytes = (uint32_t)P.bn * LG_BK * 4;
    // stage: A raw (= its tf32 hi part) | A lo as bf16 (64-byte rows, 64-byte swizzle) | B hi | B lo (tf32) | B as bf16
    const uint32_t alo_bytes = LG_BM * LG_BK * 2, bf_bytes = (uint32_t)P.bn * LG_BK * 2;
    const uint32_t stage_bytes = a_bytes + alo_bytes + 2 * b_bytes + bf_bytes;
    unsigned char *tiles = smem + ((1024u - (smem_u32(smem) & 1023u)) & 1023u);

    for (int j = tid; j < LG_MAX_N; j += PL_THREADS) {
        s_invb[j] = j < P.n_text ? P.inv_norm_b[j] : 0.f;
        s_base[j] = (j < P.n_text && P.is_base) ? P.is_base[j] : 0;
        s_gid[j] = 0;
        s_glen[j] = 0;
    }
    if (P.grouped) {
        __syncthreads();
        for (int g = tid; g < P.n_groups; g += PL_THREADS) {
            const int lo = P.goff[g], hi = P.goff[g + 1];
            for (int j = lo; j < hi && j < LG_MAX_N; ++j) s_gid[j] = (short)g;
            if (hi > lo && hi - 1 < LG_MAX_N) s_glen[hi - 1] = (short)(hi - lo);
        }
    }
    if (warp == 0 && lane == 0) {
        for (int s = 0; s < P.stages; ++s) {
            mbar_init(&s_raw[s], 1);
            mbar_init(&s_conv[s], PL_CONV);
            mbar_init(&s_empty[s], 1);
        }
        for (int b = 0; b < 2; ++b) {
            mbar_init(&s_done[b], 1);
            mbar_init(&s_accfree[b], PL_EPI);
            mbar_init(&s_ssfull[b], PL_CONV);
            mbar_init(&s_ssfree[b], PL_EPI);
        }
        mbar_fence_init();
        asm volatile("prefetch.tensormap [%0];" ::"l"(&map_a));
        asm volatile("prefetch.tensormap [%0];" ::"l"(&map_b_hi));
        asm volatile("prefetch.tensormap [%0];" ::"l"(&map_b_lo));
        asm volatile("prefetch.tensormap [%0];" ::"l"(&map_b_bf));
    }
    if (warp == 1) {
        __syncwarp();
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;"
                     ::"r"(smem_u32(&s_tmem)), "r"((uint32_t)P.tmem_cols) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_base = s_tmem;

    if (warp == 0) {
        // ===== TMA producer =====
        if (lane == 0) {
            int stage = 0;
            uint32_t phase = 0;
            for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x)
            for (int kb = 0; kb < nkb; ++kb) {
                mbar_wait(&s_empty[stage], phase ^ 1);
                unsigned char *st = tiles + (size_t)stage * stage_bytes;
                mbar_expect_tx(&s_raw[stage], a_bytes + 2 * b_bytes + bf_bytes);
                tma_load_2d(st, &map_a, kb * LG_BK, tile * LG_BM, &s_raw[stage]);
                tma_load_2d(st + a_bytes + alo_bytes, &map_b_hi, kb * LG_BK, 0, &s_raw[stage]);
                tma_load_2d(st + a_bytes + alo_bytes + b_bytes, &map_b_lo, kb * LG_BK, 0, &s_raw[stage]);
                tma_load_2d(st + a_bytes + alo_bytes + 2 * b_bytes, &map_b_bf, kb * LG_BK, 0, &s_raw[stage]);
                if (++stage == P.stages) { stage = 0; phase ^= 1; }
            }
        }
    } else if (warp == 1) {
        // ===== MMA issuer: the whole warp runs the loop (warp-uniform operands stay in uniform registers), one elected lane
        // issues; from inside `if (lane == 0)` every MMA sits in an ELECT / R2UR.BROADCAST / BRA.U.ANY loop =====
        {
            const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(P.bn >> 3) << 17) |
                                   ((uint32_t)(LG_BM >> 4) << 24);
            const uint32_t idesc2 = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)((2 * P.bn) >> 3) << 17) |
                                    ((uint32_t)(LG_BM >> 4) << 24);
            const uint32_t idesc_bf = make_idesc(LG_BM, P.bn, 1, 0, 0);       // bf16 x bf16 -> f32, both K-major
            int stage = 0, buf = 0;
            uint32_t phase = 0, bph = 0;
            for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
            mbar_wait(&s_accfree[buf], bph ^ 1);                 // the epilogue has read this accumulator buffer
            const uint32_t tmem_acc = tmem_base + (uint32_t)(buf * P.acc_cols);
            for (int kb = 0; kb < nkb; ++kb) {
                mbar_wait(&s_raw[stage], phase);                 // B tiles landed (async proxy)
                mbar_wait(&s_conv[stage], phase);                // hi / lo tiles written and fenced
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                unsigned char *st = tiles + (size_t)stage * stage_bytes;
                // A hi / B hi / B lo: tf32, 128-byte rows, 128-byte swizzle; A lo / B bf: bf16, 64-byte rows, 64-byte swizzle
                // (8-row groups 512 bytes apart)
                const uint64_t a_hi = make_sw128_desc(st);
                const uint64_t a_lo = make_sw128_desc_ex(st + a_bytes, 0, 512, 4);
                const uint64_t b_hi = make_sw128_desc(st + a_bytes + alo_bytes);
                const uint64_t b_lo = make_sw128_desc(st + a_bytes + alo_bytes + b_bytes);
                const uint64_t b_bf = make_sw128_desc_ex(st + a_bytes + alo_bytes + 2 * b_bytes, 0, 512, 4);
                if (elect_one_sync()) {
                if (P.fused) {
                    // B_hi and B_lo are adjacent 8-row-group aligned tiles: one N = 2 bn operand.  Columns
                    // [0, bn) accumulate hi*hi (+ lo*b below), columns [bn, 2 bn) hi*lo; the epilogue adds them.
#pragma unroll
                    for (int k = 0; k < LG_BK / LG_UMMA_K; ++k) {
                        const uint64_t adv = (uint64_t)((k * LG_UMMA_K * 4) >> 4);
                        umma_tf32(tmem_acc, a_hi + adv, b_hi + adv, idesc2, (kb | k) ? 1u : 0u);
                    }
                } else {
#pragma unroll
                    for (int k = 0; k < LG_BK / LG_UMMA_K; ++k) {
                        const uint64_t adv = (uint64_t)((k * LG_UMMA_K * 4) >> 4);
                        umma_tf32(tmem_acc, a_hi + adv, b_hi + adv, idesc, (kb | k) ? 1u : 0u);
                        umma_tf32(tmem_acc, a_hi + adv, b_lo + adv, idesc, 1u);
                    }
                }
                // lo (bf16, error 2^-20 |x|) x text (bf16): kind::f16, 16 channels per MMA
#pragma unroll
                for (int k = 0; k < LG_BK / 16; ++k) {
                    const uint64_t adv = (uint64_t)((k * 16 * 2) >> 4);
                    umma_f16(tmem_acc, a_lo + adv, b_bf + adv, idesc_bf, 1u);
                }
                umma_commit(&s_empty[stage]);
                if (kb == nkb - 1) umma_commit(&s_done[buf]);
                }
                __syncwarp();
                if (++stage == P.stages) { stage = 0; phase ^= 1; }
            }
            if (++buf == P.nbuf) { buf = 0; bph ^= 1; }
            }
        }
    } else {
      if (warp < 2 + PL_CONV_WARPS) {
        // ===== converter warps: thread t owns the 16-byte chunks t + PL_CONV * j =====
        const int t = tid - 64;
        int stage = 0, sb = 0;
        uint32_t phase = 0, sph = 0;
        for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        float ss[PL_CHUNKS];
#pragma unroll
        for (int j = 0; j < PL_CHUNKS; ++j) ss[j] = 0.f;
        for (int kb = 0; kb < nkb; ++kb) {
            mbar_wait(&s_raw[stage], phase);
            unsigned char *st = tiles + (size_t)stage * stage_bytes;
            const uint4 *raw = reinterpret_cast<const uint4 *>(st);  // the raw tile IS the tf32 hi operand (13 low bits ignored)
            unsigned char *lo = st + a_bytes;                       // bf16 [128 rows][32 channels], 64-byte swizzle
#pragma unroll
            for (int j = 0; j < PL_CHUNKS; ++j) {
                const int q = t + PL_CONV * j;                    // physical 16-byte chunk q of the raw tile: row q >> 3
                const uint4 v = raw[q];
                const uint32_t in[4] = {v.x, v.y, v.z, v.w};
                float l[4];
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                    const float x = __uint_as_float(in[e]);
                    l[e] = __fsub_rn(x, __uint_as_float(in[e] & 0xffffe000u));
                    ss[j] = fmaf(x, x, ss[j]);
                }
                uint2 pk;
                asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(pk.x) : "f"(l[1]), "f"(l[0]));
                asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(pk.y) : "f"(l[3]), "f"(l[2]));
                // logical chunk (channels 4 lc .. 4 lc + 3) behind the 128-byte swizzle of the raw tile, then its place in
                // the 64-byte-swizzled bf16 row: 16-byte chunk lc >> 1 (xor (row >> 1) & 3), 8-byte half lc & 1
                const int row = q >> 3, lc = (q & 7) ^ (row & 7);
                *reinterpret_cast<uint2 *>(lo + row * 64 + ((((lc >> 1) ^ ((row >> 1) & 3))) << 4) + (lc & 1) * 8) = pk;
            }
            // generic-proxy writes -> visible to the tensor core's async proxy, then signal
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            mbar_arrive(&s_conv[stage]);
            if (++stage == P.stages) { stage = 0; phase ^= 1; }
        }
        // row sums of squares: the 8 chunks of a row sit in 8 consecutive threads
        mbar_wait(&s_ssfree[sb], sph ^ 1);
#pragma unroll
        for (int j = 0; j < PL_CHUNKS; ++j) {
            float v = ss[j];
            v += __shfl_xor_sync(0xffffffffu, v, 1);
            v += __shfl_xor_sync(0xffffffffu, v, 2);
            v += __shfl_xor_sync(0xffffffffu, v, 4);
            if ((t & 7) == 0) s_ss[sb][(t >> 3) + (PL_CONV / 8) * j] = v;
        }
        mbar_arrive(&s_ssfull[sb]);
        if (++sb == P.nbuf) { sb = 0; sph ^= 1; }
        }
      } else {
        // ===== four epilogue warps (one per TMEM lane group): while they work on a tile, the producer / converters / MMAs
        // run the next one into the other accumulator buffer =====
        constexpr bool ENS = (MODE & 1) != 0, GRP = (MODE & 2) != 0;
        int buf = 0;
        uint32_t bph = 0;
        for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        const int64_t m0 = (int64_t)tile * LG_BM;
        mbar_wait(&s_ssfull[buf], bph);
        mbar_wait(&s_done[buf], bph);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const int lg = warp & 3;                                   // TMEM lane group of this warp
        const int rloc = lg * 32 + lane;
        const int64_t r = m0 + rloc;
        const bool row_ok = r < P.rows;
        const float inv_a = __fdiv_rn(1.0f, fmaxf(sqrtf(s_ss[buf][rloc]), 1e-12f));      // F.normalize
        const float b = (row_ok && P.binary) ? P.binary[r] : 0.f;
        const bool blend = P.binary != nullptr;
        const uint32_t trow = tmem_base + ((uint32_t)(lg * 32) << 16) + (uint32_t)(buf * P.acc_cols);
        // scaled cosine logit of column `col` out of the accumulator chunk(s)
        auto logit_of = [&](const uint32_t (&v)[16], const uint32_t (&v2)[16], int j, int col) {
            const float dot = P.fused ? __uint_as_float(v[j]) + __uint_as_float(v2[j]) : __uint_as_float(v[j]);
            return P.scale * ((dot * inv_a) * s_invb[col]);
        };
        float smax = -INFINITY, ssum = 0.f;
        int label = -1;
        if (ENS) {
            // logits_pred.softmax(-1): two more passes over the accumulator row (TMEM reads are cheap)
            label = row_ok ? P.mask_label[r] : -1;
            if (label >= P.n_masks) label = -1;
            for (int c0 = 0; c0 < P.bn; c0 += 16) {
                uint32_t v[16], v2[16];
                tmem_ld16(trow + (uint32_t)c0, v);
                if (P.fused) tmem_ld16(trow + (uint32_t)(P.bn + c0), v2);
#pragma unroll
                for (int j = 0; j < 16; ++j)
                    if (c0 + j < P.n_text) smax = fmaxf(smax, logit_of(v, v2, j, c0 + j));
            }
            for (int c0 = 0; c0 < P.bn; c0 += 16) {
                uint32_t v[16], v2[16];
                tmem_ld16(trow + (uint32_t)c0, v);
                if (P.fused) tmem_ld16(trow + (uint32_t)(P.bn + c0), v2);
#pragma unroll
                for (int j = 0; j < 16; ++j)
                    if (c0 + j < P.n_text) ssum += __expf(logit_of(v, v2, j, c0 + j) - smax);   // ex2.approx: ~1e-6 relative
            }
        }
        const float lsum = ENS ? logf(ssum) : 0.f;
        const bool any_unlabelled = ENS && __any_sync(0xffffffffu, label < 0);
        const float inv_ssum = ENS ? __fdiv_rn(1.0f, ssum) : 0.f;
        float best = -INFINITY;
        float gacc = P.ensemble_mean ? 0.f : -INFINITY;
        int best_i = 0;
        for (int c0 = 0; c0 < P.bn; c0 += 16) {
            uint32_t v[16], v2[16];
            tmem_ld16(trow + (uint32_t)c0, v);
            if (P.fused) tmem_ld16(trow + (uint32_t)(P.bn + c0), v2);
#pragma unroll
            for (int j = 0; j < 16; ++j) {
                const int col = c0 + j;
                if (col < P.n_text) {
                    float val = logit_of(v, v2, j, col);
                    if (ENS) {
                        // (p ** ratio * q ** (1 - ratio)).log() * overlap, base and novel halves added (:577-590), in the log
                        // domain: ratio * log p + (1 - ratio) * log q with log p = val - max - log(sum) and log q precomputed
                        // per (mask, class) — no pow / log per element.  Equal to the reference's float32 sequence up to
                        // rounding (~1e-6); where the reference's p underflows to 0 it yields NaN (-inf * 0), this form stays
                        // finite.  Branch-free per lane (rows outside every mask read mask 0's row and discard it; the class
                        // probability they need is computed only by warps that hold such a row): the 16 loads of a chunk
                        // issue back to back instead of one per divergent branch.
                        const float logp = (val - smax) - lsum;
                        const float logq = __ldg(P.mask_probs + (size_t)(label >= 0 ? label : 0) * P.n_text + col);
                        const float ratio = s_base[col] ? P.base_ratio : P.novel_ratio;
                        const float ens = fmaf(ratio, logp, (1.f - ratio) * logq);
                        if (any_unlabelled) {
                            const float prob = __expf(val - smax) * inv_ssum;                // class probability
                            val = label >= 0 ? ens : prob;
                        } else {
                            val = ens;
                        }
                    }
                    if (blend) {
                        // binary * logits_base + (1 - binary) * logits_novel, masked entries = -1e10
                        const float lb = s_base[col] ? val : -1e10f, ln = s_base[col] ? -1e10f : val;
                        val = __fadd_rn(__fmul_rn(b, lb), __fmul_rn(__fsub_rn(1.0f, b), ln));
                    }
                    if (GRP) {
                        gacc = P.ensemble_mean ? gacc + val : fmaxf(gacc, val);
                        const int glen = s_glen[col];
                        if (glen > 0) {
                            const float o = P.ensemble_mean ? __fdiv_rn(gacc, (float)glen) : gacc;
                            const int g = s_gid[col];
                            if (row_ok && P.out) P.out[r * P.n_groups + g] = o;
                            if (o > best) { best = o; best_i = g; }
                            gacc = P.ensemble_mean ? 0.f : -INFINITY;
                        }
                    } else {
                        if (row_ok && P.out) P.out[r * P.n_text + col] = val;
                        if (val > best) { best = val; best_i = col; }
                    }
                }
            }
        }
        if (row_ok && P.argmax) P.argmax[r] = best_i;
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
        mbar_arrive(&s_accfree[buf]);
        mbar_arrive(&s_ssfree[buf]);
        if (++buf == P.nbuf) { buf = 0; bph ^= 1; }
        }
      }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 1) {
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        __syncwarp();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"((uint32_t)P.tmem_cols)
                     : "memory");
    }
}

}  // namespace xm3d

using namespace xm3d;

static int launch_point_logits(const float *feat, int64_t rows, int c, const float *b_hi, const float *b_lo,
                               const unsigned short *b_bf, const float *inv_b, int n_cols, PointLogitsParams P,
                               cudaStream_t stream, const char *who);

extern "C" size_t xm3d_logits_ws_bytes(int64_t rows, int32_t n_text, int32_t c, int32_t n_groups) {
    (void)rows;
    Carver cv(nullptr);
    cv.take<float>((size_t)(n_text + 1) * c);
    cv.take<float>((size_t)(n_text + 1) * c);
    cv.take<unsigned short>((size_t)(n_text + 1) * ((c + 7) / 8 * 8));
    cv.take<float>((size_t)(n_text + 1));
    cv.take<int>((size_t)n_groups + 2);
    return cv.off + 256;
}

extern "C" int xm3d_logits(const float *mask_embed, int64_t rows, int32_t c, const float *text_embed,
                           int32_t n_text, const float *null_embed, const int32_t *group_off_host,
                           int32_t n_groups, int32_t ensemble_mean, float logit_scale, float *out,
                           int32_t *argmax, void *ws, size_t ws_bytes, xm3d_stream_t stream_) {
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    XM3D_REQUIRE(rows >= 0 && c > 0 && n_text > 0 && n_groups > 0, "bad sizes");
    XM3D_REQUIRE(c % 4 == 0, "embedding width must be a multiple of 4");
    XM3D_REQUIRE(n_text + 1 <= LG_MAX_N && n_groups + 2 <= 512, "at most 255 text embeddings");
    XM3D_REQUIRE(mask_embed && text_embed && null_embed && group_off_host && out && ws, "null pointer");
    XM3D_REQUIRE(group_off_host[0] == 0 && group_off_host[n_groups] == n_text, "group offsets must span [0, n_text]");
    XM3D_REQUIRE(reinterpret_cast<uintptr_t>(mask_embed) % 16 == 0, "mask_embed must be 16-byte aligned");
    XM3D_REQUIRE(rows < ((int64_t)1 << 31) - LG_BM, "rows exceed int32 tile coordinates");
    for (int g = 0; g < n_groups; ++g)
        XM3D_REQUIRE(group_off_host[g + 1] > group_off_host[g], "empty label group");
    if (rows == 0) return XM3D_OK;
    if (ws_bytes < xm3d_logits_ws_bytes(rows, n_text, c, n_groups)) {
        set_error("xm3d_logits: workspace too small");
        return XM3D_ERR_WORKSPACE;
    }
    const int n_cols = n_text + 1;                        // the null embedding is one more column, its own group
    Carver cv(ws);
    const int ld_bf = (c + 7) / 8 * 8;
    float *b_hi = cv.take<float>((size_t)n_cols * c);
    float *b_lo = cv.take<float>((size_t)n_cols * c);
    unsigned short *b_bf = cv.take<unsigned short>((size_t)n_cols * ld_bf);
    float *inv_b = cv.take<float>((size_t)n_cols);
    cv.take<int>((size_t)n_groups + 2);                   // (kept in the workspace layout; the offsets travel as kernel parameters)
    logits_prep_kernel<<<(unsigned)((n_cols + 7) / 8), 256, 0, stream>>>(text_embed, null_embed, n_text, 1, c, b_hi, b_lo, b_bf, ld_bf, inv_b);
    count_launches(1);
    PointLogitsParams P;
    memset(&P, 0, sizeof(P));
    P.scale = logit_scale; P.out = out; P.argmax = argmax;
    P.grouped = 1; P.n_groups = n_groups + 1; P.ensemble_mean = ensemble_mean;
    for (int g = 0; g <= n_groups; ++g) P.goff[g] = (unsigned short)group_off_host[g];
    P.goff[n_groups + 1] = (unsigned short)n_cols;         // the null embedding: one more column, its own group
    return launch_point_logits(mask_embed, rows, c, b_hi, b_lo, b_bf, inv_b, n_cols, P, stream, "xm3d_logits");
}

extern "C" size_t xm3d_point_logits_ws_bytes(int32_t n_text, int32_t c) {
    Carver cv(nullptr);
    cv.take<float>((size_t)n_text * c);
    cv.take<float>((size_t)n_text * c);
    cv.take<unsigned short>((size_t)n_text * ((c + 7) / 8 * 8));
    cv.take<float>((size_t)n_text);
    return cv.off + 256;
}

extern "C" int xm3d_point_logits(const float *feat, int64_t rows, int32_t c, const float *text_embed, int32_t n_text,
                                 float logit_scale, const float *binary, const uint8_t *is_base,
                                 const int32_t *mask_label, const float *mask_probs, int32_t n_masks, float base_ratio,
                                 float novel_ratio, float *out, int32_t *argmax, void *ws, size_t ws_bytes,
                                 xm3d_stream_t stream_) {
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    XM3D_REQUIRE(rows >= 0 && c > 0 && n_text > 0, "bad sizes");
    XM3D_REQUIRE(c % 4 == 0, "feature width must be a multiple of 4");
    XM3D_REQUIRE(n_text <= LG_MAX_N, "at most 256 classes");
    XM3D_REQUIRE(feat && text_embed && ws && (out || argmax), "null pointer");
    XM3D_REQUIRE(!binary || is_base, "is_base is required with binary");
    XM3D_REQUIRE(!mask_label || (mask_probs && is_base && n_masks >= 0), "the ensemble needs mask_probs and is_base");
    XM3D_REQUIRE(rows < ((int64_t)1 << 31) - LG_BM, "rows exceed int32 tile coordinates");
    XM3D_REQUIRE(reinterpret_cast<uintptr_t>(feat) % 16 == 0, "feat must be 16-byte aligned");
    if (rows == 0) return XM3D_OK;
    if (ws_bytes < xm3d_point_logits_ws_bytes(n_text, c)) {
        set_error("xm3d_point_logits: workspace too small");
        return XM3D_ERR_WORKSPACE;
    }
    Carver cv(ws);
    const int ld_bf = (c + 7) / 8 * 8;
    float *b_hi = cv.take<float>((size_t)n_text * c);
    float *b_lo = cv.take<float>((size_t)n_text * c);
    unsigned short *b_bf = cv.take<unsigned short>((size_t)n_text * ld_bf);
    float *inv_b = cv.take<float>((size_t)n_text);
    logits_prep_kernel<<<(unsigned)((n_text + 7) / 8), 256, 0, stream>>>(text_embed, nullptr, n_text, 0, c, b_hi, b_lo, b_bf, ld_bf, inv_b);
    count_launches(1);

    PointLogitsParams P;
    memset(&P, 0, sizeof(P));
    P.scale = logit_scale; P.binary = binary; P.is_base = is_base; P.out = out; P.argmax = argmax;
    P.mask_label = mask_label; P.mask_probs = mask_probs; P.n_masks = n_masks; P.base_ratio = base_ratio; P.novel_ratio = novel_ratio;
    return launch_point_logits(feat, rows, c, b_hi, b_lo, b_bf, inv_b, n_text, P, stream, "xm3d_point_logits");
}

static int launch_point_logits(const float *feat, int64_t rows, int c, const float *b_hi, const float *b_lo,
                               const unsigned short *b_bf, const float *inv_b, int n_text, PointLogitsParams P,
                               cudaStream_t stream, const char *who) {
    P.rows = rows; P.c = c; P.n_text = n_text; P.bn = (n_text + 15) / 16 * 16; P.inv_norm_b = inv_b;
    P.fused = (2 * P.bn <= 256) ? 1 : 0;
    int tc = 32;
    while (tc < (P.fused ? 2 * P.bn : P.bn)) tc <<= 1;
    P.acc_cols = tc;
    // A raw + A lo (bf16) + B hi + B lo (tf32) + B (bf16)
    const size_t stage_bytes = (size_t)LG_BM * LG_BK * 6 + (size_t)P.bn * LG_BK * 10;
    // two CTAs per SM when two stages of each fit (the second CTA's TMA / MMA hides the first one's
    // conversion pass), else one CTA with as many stages as fit
    int stages = (int)((108 * 1024) / stage_bytes);
    if (stages < 2) stages = (int)((220 * 1024) / stage_bytes);
    if (stages > 4) stages = 4;
    if (stages < 1) { set_error("%s: tile does not fit shared memory", who); return XM3D_ERR_UNSUPPORTED; }
    P.stages = stages;
    // persistent CTAs (two per SM when two fit): the epilogue of a tile overlaps the next tile's loads and MMAs through a
    // second accumulator buffer when tensor memory has room for it (512 columns per SM, shared by the co-resident CTAs)
    const int ctas_per_sm = (stage_bytes * (size_t)stages + 1024 <= 110 * 1024) ? 2 : 1;
    P.nbuf = (ctas_per_sm * 2 * tc <= 512) ? 2 : 1;
    P.tmem_cols = P.nbuf * tc;
    CUtensorMap ma, mbh, mbl, mbf;
    if (!make_map(&ma, feat, rows, c, LG_BM) || !make_map(&mbh, b_hi, n_text, c, P.bn) ||
        !make_map(&mbl, b_lo, n_text, c, P.bn) || !make_map_bf16_sw64(&mbf, b_bf, n_text, (c + 7) / 8 * 8, P.bn)) {
        set_error("%s: cuTensorMapEncodeTiled failed", who);
        return XM3D_ERR_CUDA;
    }
    static std::atomic<uint64_t> attr_set{0};
    if (first_use_on_device(&attr_set)) {
        cudaFuncSetAttribute(point_logits_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, 221 * 1024);   // + 4 KB static + 1 KB reserved <= 227 KB
        cudaFuncSetAttribute(point_logits_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 221 * 1024);   // + 4 KB static + 1 KB reserved <= 227 KB
        cudaFuncSetAttribute(point_logits_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, 221 * 1024);   // + 4 KB static + 1 KB reserved <= 227 KB
        cudaFuncSetAttribute(point_logits_kernel<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, 221 * 1024);   // + 4 KB static + 1 KB reserved <= 227 KB
    }
    const int64_t n_tiles = (rows + LG_BM - 1) / LG_BM;
    const int64_t slots = (int64_t)sm_count() * ctas_per_sm;
    const unsigned grid = (unsigned)(n_tiles < slots ? n_tiles : slots);
    const size_t smem = stage_bytes * stages + 1024;
    switch ((P.mask_label ? 1 : 0) | (P.grouped ? 2 : 0)) {
        case 0: point_logits_kernel<0><<<grid, PL_THREADS, smem, stream>>>(ma, mbh, mbl, mbf, P); break;
        case 1: point_logits_kernel<1><<<grid, PL_THREADS, smem, stream>>>(ma, mbh, mbl, mbf, P); break;
        case 2: point_logits_kernel<2><<<grid, PL_THREADS, smem, stream>>>(ma, mbh, mbl, mbf, P); break;
        default: point_logits_kernel<3><<<grid, PL_THREADS, smem, stream>>>(ma, mbh, mbl, mbf, P); break;
    }
    count_launches(1);
    return check_launch(who);
}

