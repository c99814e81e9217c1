// Stage 3, overlapping masks — segmented pooling on the tensor cores.
//
// Replaces the per-mask `feature[mask_3d[k]]` -> mean(0) loop of the reference for its RAW thresholded
// predictions (models/utils/criterion.py:83-85, 148-157), where a point belongs to several masks (7.7 on
// average in the synthetic overlap variant).  The pooled sums are a contraction over the points,
//
//     sum[k, c] = sum_p member[k, p] * F[p, c]          member in {0, 1},
//
// and run here as tcgen05.mma with the accumulator in tensor memory:
//   D[M = masks (64 / 128), N = 128 channels] += A[masks, K points] * B[K points, 128 channels]
//   B = the feature tile exactly as it lies in HBM (row = point, 128 bytes = 32 channels): an N-MAJOR operand, brought in
//       by TMA with the 128-byte swizzle of 32-byte atoms — the one layout tcgen05 accepts for N-major 32-bit operands
//       (four 32-channel column blocks x 64 points = 32 KB per tile).  The tensor core ignores the 13 low mantissa bits of
//       a tf32 operand, so the raw tile IS its hi part (kind::tf32, 8 points per MMA); converter warps write
//       lo = x - hi as a bf16 N-major tile (kind::f16, 16 points per MMA; its error 2^-19 |x| is far inside the 1e-5 bar).
//   A = the membership bits of the tile's 64 points as 0.0f / 1.0f (tf32) and bf16 pairs IN TENSOR MEMORY (tcgen05.mma
//       with A from TMEM): no shared-memory write and no operand read for the 0/1 matrix.
// The accumulator is flushed every P2_GROUP tiles into float32 registers of the epilogue warps with round-to-nearest
// adds: a whole segment accumulated inside the tensor core loses precision (its fp32 accumulation truncates; measured
// 1.1e-4 at 70 k points), and the result stays deterministic.
// A work item is (segment, 128-channel slice), handed out by DESCENDING segment size through an atomic counter to one
// persistent CTA per SM; every feature row is read from HBM exactly once, independent of the number of memberships.
// A NaN / Inf feature would reach, multiplied by a 0 membership, masks of its tile it does not belong to: the converters
// detect it (XM3D_FLAG_NONFINITE) and the caller pools that batch with the CUDA-core kernel.
//
// Warp roles (640 threads): 0 TMA producer + work-item ring | 1 MMA issue | 2-5 converters | 6-9 builders (one per TMEM
// lane group) | 10-17 epilogue | 18-19 preparation (membership words -> transposed rows in shared memory).
// History, measurements and what was rejected: DESIGN.md section 4.4.  (The first version — membership operand as a K-major
// B tile in shared memory, 208 KB of shared-memory traffic per 32 KB tile, 2.08 ms — was removed in round 2.)
#include <cuda.h>
#include <string.h>
#include <type_traits>

#include "common.cuh"
#include "tc.cuh"
#include "vec.cuh"

namespace xm3d {

constexpr int PM_TP = 64;                            // points per tile = 8 k-steps of 8 (tf32)
constexpr int PM_SLICE = 128;                        // channels per work item = UMMA M
constexpr int PM_RAW = PM_TP * PM_SLICE * 4;         // bytes of a raw / lo tile (32 KB)
constexpr int PM_CB = PM_TP * 128;                   // bytes of one 32-channel column block (8 KB)
constexpr int PM_CONV = 128, PM_BUILD = 128;         // converter / builder threads
constexpr int PM_MAX_STAGES = 4;

struct PoolMmaParams {
    const uint32_t *member;      // [cap, words]
    const int64_t *seg_off;      // [n_seg + 1]
    int words, k, n_seg, c;
    int64_t cap;
    float *sum, *mean;           // [n_seg, k, c]
    int32_t *cnt;                // [n_seg, k] or null
    int *work;                   // item counter (zeroed by the host)
    const int *order;            // [n_seg] segments by descending size, or null (identity)
    int raw_stages, conv_stages;
    int dbg;                     // experiments only: 1 = skip the lo MMAs, 2 = skip the hi MMAs (results are then wrong)
    int dbg2;                    // experiments only: bit 0 = no TMA loads, bit 1 = converters idle (results are then wrong)
    int m64;                     // version 2: k <= 64 -> MMAs with M = 64 (half the accumulator read-modify-write per MMA)
    int32_t *status;             // XM3D_FLAG_NONFINITE is raised here (may be null)
};

// Work items are handed out through an atomic counter; with segments of very different sizes (0 .. 60 k points per
// view) the CTA that draws a large item last finishes late: measured 1 990 tiles on the busiest CTA against a mean of
// 1 485.  Handing the segments out by DESCENDING size (longest-processing-time first) removes that tail.
__global__ void __launch_bounds__(256)
pool_order_kernel(const int64_t *__restrict__ seg_off, int n_seg, int *__restrict__ order) {
    for (int s = blockIdx.x * blockDim.x + threadIdx.x; s < n_seg; s += gridDim.x * blockDim.x) {
        const int64_t ns = seg_off[s + 1] - seg_off[s];
        int rank = 0;
        for (int t = 0; t < n_seg; ++t) {
            const int64_t nt = seg_off[t + 1] - seg_off[t];
            rank += (nt > ns) || (nt == ns && t < s);
        }
        order[rank] = s;
    }
}

// Instrumented build (-DXM3D_PM_TIMING, scripts/exp_pool_mma.py): per-role wait / work cycles of every CTA
#ifdef XM3D_PM_TIMING
__device__ long long *g_pm_dbg = nullptr;
// experiment switches (results are then wrong, the timing is what is measured): dbg 1 / 2 = skip the lo / hi MMAs;
// dbg2 bit 0 = no TMA loads, bit 1 = idle converters
#define PM_DBG(P) ((P).dbg)
#define PM_DBG2(P) ((P).dbg2)
#define PM_CLK() clock64()
#define PM_ACC(var, t0) do { var += clock64() - (t0); } while (0)
#define PM_OUT(role, a, b, c, d) do { if (g_pm_dbg) { long long *o_ = g_pm_dbg + ((size_t)blockIdx.x * 8 + (role)) * 4; \
    o_[0] += (a); o_[1] += (b); o_[2] += (c); o_[3] += (d); } } while (0)
#else
#define PM_DBG(P) 0
#define PM_DBG2(P) 0
#define PM_CLK() 0ll
#define PM_ACC(var, t0) do { (void)(t0); } while (0)
#define PM_OUT(role, a, b, c, d) do { } while (0)
#endif

// TMEM map (512 columns): [0, 128) the accumulator; then P2_CS membership stages of 64 points as tf32 (64 columns each) and the
// same as bf16 pairs (32 columns each).  Shared-memory traffic per 32 KB tile: 32 (TMA) + 32 + 16 (converter) + 32 + 16 (MMA
// operand reads) = 128 KB.
constexpr int P2_LO = PM_TP * PM_SLICE * 2;          // bytes of a bf16 lo tile (16 KB)
constexpr int P2_CS = 3;                             // membership (tensor memory) / lo-tile (shared memory) stages
constexpr int P2_DB = 1;                             // accumulator buffers
constexpr int P2_D = 0, P2_A32 = P2_DB * PM_SLICE, P2_A16 = P2_A32 + P2_CS * 64;   // TMEM column bases
static_assert(P2_A16 + P2_CS * 32 <= 512, "tensor memory columns");
constexpr int P2_MAX_DYN_SMEM = PM_MAX_STAGES * PM_RAW + P2_CS * PM_TP * PM_SLICE * 2 + 1024;   // + ~10 KB static < 227 KB
constexpr int P2_IQ = 4;                                   // work items published ahead of their consumers
constexpr int P2_MW = 8;                                   // tiles of membership words in flight (asynchronous copies)
constexpr int P2_GROUP = 16;                               // tiles (of 64 points) per accumulator flush
constexpr int P2_TW = 4;                                   // transposed-word stages between the preparation warps and the builders
constexpr int P2_EPI_WARPS = 8, P2_THREADS = 32 * (10 + P2_EPI_WARPS + 2);   // + 2 preparation warps

struct P2Item { int s, sl, n; int64_t a; };

// Producer side of the item ring (whole warp calls, lane 0 works): draw the next item from the global counter, publish it.
__device__ __forceinline__ P2Item p2_publish_item(const PoolMmaParams &P, int4 *s_iq, long long *s_iq_a, uint64_t *s_full,
                                                  uint64_t *s_free, int &iq, uint32_t &iqph, int &next_item, int n_items, int nsl,
                                                  bool over, int lane) {
    P2Item it{0, 0, 0, 0};
    if (lane == 0) {
        mbar_wait(&s_free[iq], iqph ^ 1);
        // drawn only now (all loads of the previous item are issued): drawing one item ahead takes the last items away
        // from the CTAs that would be free for them (measured: 1 638 instead of 1 517 tiles on the busiest CTA)
        const int item = next_item ? n_items : atomicAdd(P.work, 1);
        if (item < n_items) {
            it.s = P.order ? P.order[item / nsl] : item / nsl;
            it.sl = item % nsl;
            it.a = P.seg_off[it.s];
            it.n = over ? 0 : (int)(P.seg_off[it.s + 1] - it.a);
        } else {
            it.n = -1;
            next_item = 1;                             // the counter ran out: do not touch it again
        }
        s_iq[iq] = make_int4(it.s, it.sl, it.n, 0);
        s_iq_a[iq] = it.a;
        mbar_arrive(&s_full[iq]);
    }
    it.s = __shfl_sync(0xffffffffu, it.s, 0); it.sl = __shfl_sync(0xffffffffu, it.sl, 0);
    it.n = __shfl_sync(0xffffffffu, it.n, 0); it.a = __shfl_sync(0xffffffffu, it.a, 0);
    if (++iq == P2_IQ) { iq = 0; iqph ^= 1; }
    return it;
}

// Consumer side (whole warp calls): lane 0 reads the record and frees the slot for this warp.
__device__ __forceinline__ P2Item p2_take_item(const int4 *s_iq, const long long *s_iq_a, uint64_t *s_full, uint64_t *s_free,
                                               int &iq, uint32_t &iqph, int lane) {
    P2Item it{0, 0, 0, 0};
    if (lane == 0) {
        mbar_wait(&s_full[iq], iqph);
        const int4 r = s_iq[iq];
        it.a = s_iq_a[iq];
        mbar_arrive(&s_free[iq]);
        it.s = r.x; it.sl = r.y; it.n = r.z;
    }
    it.s = __shfl_sync(0xffffffffu, it.s, 0); it.sl = __shfl_sync(0xffffffffu, it.sl, 0);
    it.n = __shfl_sync(0xffffffffu, it.n, 0); it.a = __shfl_sync(0xffffffffu, it.a, 0);
    if (++iq == P2_IQ) { iq = 0; iqph ^= 1; }
    return it;
}

template <int UNR>                                    // converter loads in flight per thread (shared-memory latency under load)
__global__ void __launch_bounds__(P2_THREADS, 1)   // 18 warps = 5 on one scheduler: 96 registers per thread (112 do not fit its 16 K)
pool_mma2_kernel(const __grid_constant__ CUtensorMap map, const PoolMmaParams P) {
    constexpr int EPI_THREADS = P2_EPI_WARPS * 32;
    extern __shared__ __align__(1024) unsigned char pm_smem[];
    __shared__ uint64_t s_raw_full[PM_MAX_STAGES], s_raw_empty[PM_MAX_STAGES];
    __shared__ uint64_t s_conv_full[P2_CS], s_conv_empty[P2_CS];
    __shared__ uint64_t s_tile_done[P2_DB], s_tmem_free[P2_DB];
    __shared__ uint32_t s_tmem;
    // Work items flow through the roles WITHOUT a CTA-wide barrier (draining and refilling the TMA -> converter -> MMA ->
    // epilogue pipeline at every item cost ~7 % of the kernel): the producer thread draws the item from the global counter
    // and publishes (segment, slice, first row, rows) in a ring; lane 0 of every other warp reads it in order.
    __shared__ uint64_t s_iq_full[P2_IQ], s_iq_free[P2_IQ];
    __shared__ int4 s_iq[P2_IQ];                     // x = segment, y = slice, z = rows (< 0: no more items), w = unused
    __shared__ long long s_iq_a[P2_IQ];              // first row of the segment
    __shared__ uint64_t s_cnt_full[2], s_cnt_free[2];
    __shared__ int s_cnt[2][128];                    // per-mask counts of the builders' item -> epilogue (two items in flight)
    __shared__ uint32_t s_mw[P2_MW][4][PM_TP];       // membership words of the next P2_MW tiles (preparation warps)
    __shared__ uint32_t s_tw[P2_TW][4][2][32];       // transposed words [stage][word][points 0..31 / 32..63][mask row]
    __shared__ uint64_t s_tw_full[P2_TW], s_tw_empty[P2_TW];

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    unsigned char *base = pm_smem + ((1024u - (smem_u32(pm_smem) & 1023u)) & 1023u);
    unsigned char *raw_base = base;
    unsigned char *lo_base = raw_base + (size_t)P.raw_stages * PM_RAW;

    if (tid == 0) {
        for (int s = 0; s < PM_MAX_STAGES; ++s) { mbar_init(&s_raw_full[s], 1); mbar_init(&s_raw_empty[s], 1); }
        for (int b = 0; b < P2_CS; ++b) {
            mbar_init(&s_conv_full[b], PM_CONV + PM_BUILD);
            mbar_init(&s_conv_empty[b], 1);
        }
        for (int b = 0; b < P2_DB; ++b) {
            mbar_init(&s_tile_done[b], 1);
            mbar_init(&s_tmem_free[b], EPI_THREADS);
        }
        for (int q = 0; q < P2_TW; ++q) { mbar_init(&s_tw_full[q], 64); mbar_init(&s_tw_empty[q], PM_BUILD); }
        for (int q = 0; q < P2_IQ; ++q) { mbar_init(&s_iq_full[q], 1); mbar_init(&s_iq_free[q], P2_THREADS / 32 - 1); }
        for (int b = 0; b < 2; ++b) { mbar_init(&s_cnt_full[b], PM_BUILD); mbar_init(&s_cnt_free[b], EPI_THREADS); }
        mbar_fence_init();
        asm volatile("prefetch.tensormap [%0];" ::"l"(&map));
    }
    if (warp == 1) {
        __syncwarp();
        tmem_alloc(&s_tmem, 512);
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = s_tmem;

    const int nsl = P.c / PM_SLICE;
    const int n_items = P.n_seg * nsl;
    const int64_t total_rows = P.seg_off[P.n_seg];
    const bool over = total_rows > P.cap;
    int rs = 0, cs = 0, buf = 0;
    uint32_t rph = 0, cph = 0, bph = 0;
    {
        if (warp == 0) {
            int next_item = 0;                             // becomes 1 when the counter ran out
            int iq = 0; uint32_t iqph = 0;
            for (;;) {
                const P2Item it = p2_publish_item(P, s_iq, s_iq_a, s_iq_full, s_iq_free, iq, iqph, next_item, n_items, nsl, over, lane);
                if (it.n < 0) break;
                const int s = it.s, sl = it.sl, n = it.n; const int64_t a = it.a;
                const int ntile = (n + PM_TP - 1) / PM_TP;
                (void)s; (void)sl; (void)a;
                // ===== TMA producer =====
                if (lane == 0) {
                    long long tw0 = 0, tall = PM_CLK();
                    for (int t = 0; t < ntile; ++t) {
                        long long c0_ = PM_CLK();
                        mbar_wait_sleep(&s_raw_empty[rs], rph ^ 1, 200);
                        PM_ACC(tw0, c0_);
                        unsigned char *dst = raw_base + (size_t)rs * PM_RAW;
                        if (PM_DBG2(P) & 1) {
                            mbar_arrive(&s_raw_full[rs]);
                        } else {
                            mbar_expect_tx(&s_raw_full[rs], PM_RAW);
                            const int y = (int)(a + (int64_t)t * PM_TP);
#pragma unroll
                            for (int cb = 0; cb < 4; ++cb)
                                tma_load_2d(dst + cb * PM_CB, &map, sl * PM_SLICE + cb * 32, y, &s_raw_full[rs]);
                        }
                        if (++rs == P.raw_stages) { rs = 0; rph ^= 1; }
                    }
                    PM_OUT(0, tw0, 0, PM_CLK() - tall, ntile);
                }
            }
        } else if (warp == 1) {
            int iq = 0; uint32_t iqph = 0;
            for (;;) {
                const P2Item it = p2_take_item(s_iq, s_iq_a, s_iq_full, s_iq_free, iq, iqph, lane);
                if (it.n < 0) break;
                const int s = it.s, sl = it.sl, n = it.n; const int64_t a = it.a;
                const int ntile = (n + PM_TP - 1) / PM_TP;
                (void)s; (void)sl; (void)a;
                // ===== MMA issuer: 8 tf32 MMAs (hi, 8 points each) + 4 bf16 MMAs (lo, 16 points each) per tile =====
                // The WHOLE warp runs this loop and one elected lane issues: every operand is then warp-uniform for the compiler
                // and lives in uniform registers.  Issued from inside `if (lane == 0)` each MMA was wrapped in an
                // ELECT / 3 x R2UR.BROADCAST / BRA.U.ANY loop — ~55 cycles of serial issue per MMA, 650 per tile, with the
                // tensor pipe idle in between (the pipe itself needs ~72 cycles per MMA).
                {
                    // M = 64 when all masks fit 64 rows (same MMA time, half the tensor work)
                    const int mrows = P.m64 ? 64 : 128;
                    const uint32_t id32 = make_idesc(mrows, PM_SLICE, 2, /*A: TMEM*/ 0, /*B N-major*/ 1);
                    const uint32_t id16 = make_idesc(mrows, PM_SLICE, 1, 0, 1);
                    // tf32 N-major: 32-byte-atom swizzle, 4 column blocks 8 KB apart; bf16 N-major: 128-byte swizzle, 2 blocks of
                    // 64 channels 8 KB apart
                    const uint64_t d_hi0 = make_sw128_desc_ex(raw_base, PM_CB, 512, 1);
                    const uint64_t d_lo0 = make_sw128_desc_ex(lo_base, PM_TP * 128, 1024, 2);
                    long long tw0 = 0, tw1 = 0, tw2 = 0, tall = PM_CLK();
                    for (int t = 0; t < ntile; ++t) {
                        const bool first = t % P2_GROUP == 0, last = (t % P2_GROUP == P2_GROUP - 1) || t == ntile - 1;
                        long long c0_ = PM_CLK();
                        // conv_full implies raw_full: every converter thread has seen the tile's TMA barrier complete before it
                        // arrives here (mbarrier operations are cumulative), so the issuer waits for ONE barrier per tile
                        PM_ACC(tw0, c0_); c0_ = PM_CLK();
                        mbar_wait(&s_conv_full[cs], cph);
                        PM_ACC(tw1, c0_); c0_ = PM_CLK();
                        if (first) mbar_wait(&s_tmem_free[buf], bph ^ 1);      // a fresh accumulator every P2_GROUP tiles
                        PM_ACC(tw2, c0_);
                        tc_fence_after();
                        const uint64_t b_hi = d_hi0 + (uint64_t)((rs * PM_RAW) >> 4);      // (start address field: bytes >> 4)
                        const uint64_t b_lo = d_lo0 + (uint64_t)((cs * P2_LO) >> 4);
                        const uint32_t d = tmem_base + (uint32_t)(P2_D + buf * PM_SLICE);
                        const uint32_t a32 = tmem_base + (uint32_t)(P2_A32 + cs * 64), a16 = tmem_base + (uint32_t)(P2_A16 + cs * 32);
                        if (elect_one_sync()) {
                            if (PM_DBG(P) != 2) {
#pragma unroll
                                for (int ks = 0; ks < PM_TP / 8; ++ks)
                                    umma_tf32_ts(d, a32 + ks * 8, b_hi + (uint64_t)((ks * 1024) >> 4), id32, (ks || !first) ? 1u : 0u);
                            }
                            if (PM_DBG(P) != 1) {
#pragma unroll
                                for (int ks = 0; ks < PM_TP / 16; ++ks)
                                    umma_f16_ts(d, a16 + ks * 8, b_lo + (uint64_t)((ks * 2048) >> 4), id16,
                                                (PM_DBG(P) == 2 && ks == 0 && first) ? 0u : 1u);
                            }
                            umma_commit(&s_raw_empty[rs]);
                            umma_commit(&s_conv_empty[cs]);
                            if (last) umma_commit(&s_tile_done[buf]);
                        }
                        __syncwarp();
                        if (++rs == P.raw_stages) { rs = 0; rph ^= 1; }
                        if (++cs == P2_CS) { cs = 0; cph ^= 1; }
                        if (last && ++buf == P2_DB) { buf = 0; bph ^= 1; }
                    }
                    if (lane == 0) PM_OUT(1, tw0, tw1, tw2, PM_CLK() - tall);
                }
            }
        } else if (warp < 6) {
            int iq = 0; uint32_t iqph = 0;
            for (;;) {
                const P2Item it = p2_take_item(s_iq, s_iq_a, s_iq_full, s_iq_free, iq, iqph, lane);
                if (it.n < 0) break;
                const int s = it.s, sl = it.sl, n = it.n; const int64_t a = it.a;
                const int ntile = (n + PM_TP - 1) / PM_TP;
                (void)s; (void)sl; (void)a;
                // ===== converters: lo = bf16(x - hi) into a 128-byte-swizzled N-major tile =====
                const int t0 = tid - 64;
                const int r0 = t0 >> 3, p8 = t0 & 7;
                // logical 32-byte chunk of this thread's physical chunk (32-byte-atom swizzle of the raw tile), its 16-byte
                // chunk inside the 64-channel block of the lo tile for even / odd 32-channel column blocks, swizzled by the row
                const int l32 = (p8 >> 1) ^ (r0 & 3);
                const int lo_even = r0 * 128 + ((l32 ^ (r0 & 7)) << 4) + (p8 & 1) * 8;
                const int lo_odd = r0 * 128 + (((4 + l32) ^ (r0 & 7)) << 4) + (p8 & 1) * 8;
                long long tw0 = 0, tw1 = 0, tw2 = 0, tall = PM_CLK();
                // A NaN / Inf feature would reach, multiplied by a 0 membership, masks it does not belong to: lo of such an
                // element is NaN (Inf - Inf), and NaN * 0 stays NaN in these four chains -> XM3D_FLAG_NONFINITE.
                float nf0 = 0.f, nf1 = 0.f, nf2 = 0.f, nf3 = 0.f;
                for (int t = 0; t < ntile; ++t) {
                    long long c0_ = PM_CLK();
                    mbar_wait(&s_raw_full[rs], rph);
                    PM_ACC(tw0, c0_); c0_ = PM_CLK();
                    mbar_wait(&s_conv_empty[cs], cph ^ 1);
                    PM_ACC(tw1, c0_);
                    uint4 *raw = reinterpret_cast<uint4 *>(raw_base + (size_t)rs * PM_RAW);
                    unsigned char *lo = lo_base + (size_t)cs * P2_LO;
                    const int rows_valid = min(PM_TP, n - t * PM_TP);
                    // Chunk j of this thread is physical chunk q = t0 + 128 j of the raw tile ([column block][row][8 chunks]):
                    // row = r0 + 16 (j & 3), column block = j >> 2, chunk-in-row p8 — so every address below is a per-thread
                    // base plus a compile-time offset (the loop is fully unrolled).  UNR loads are issued back to back before
                    // the first one is used (the shared-memory pipe is shared with the TMA writes and the MMA operand reads).
                    // full tiles (all but the last of a segment) take the branch-free copy of the loop
                    auto convert = [&](auto full_tag) {
                        constexpr bool FULL = decltype(full_tag)::value;
                        for (int j0 = 0; j0 < ((PM_DBG2(P) & 2) ? 0 : PM_RAW / 16 / PM_CONV); j0 += UNR) {
                            uint4 v[UNR];
#pragma unroll
                            for (int u = 0; u < UNR; ++u) v[u] = raw[t0 + PM_CONV * (j0 + u)];
#pragma unroll
                            for (int u = 0; u < UNR; ++u) {
                                const int j = j0 + u;
                                uint2 l = make_uint2(0u, 0u);
                                if (FULL || r0 + 16 * (j & 3) < rows_valid) {
                                    const float l0 = __fsub_rn(__uint_as_float(v[u].x), __uint_as_float(v[u].x & 0xffffe000u));
                                    const float l1 = __fsub_rn(__uint_as_float(v[u].y), __uint_as_float(v[u].y & 0xffffe000u));
                                    const float l2 = __fsub_rn(__uint_as_float(v[u].z), __uint_as_float(v[u].z & 0xffffe000u));
                                    const float l3 = __fsub_rn(__uint_as_float(v[u].w), __uint_as_float(v[u].w & 0xffffe000u));
                                    asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(l.x) : "f"(l1), "f"(l0));
                                    asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(l.y) : "f"(l3), "f"(l2));
                                    nf0 = __fmaf_rn(l0, 0.f, nf0); nf1 = __fmaf_rn(l1, 0.f, nf1);
                                    nf2 = __fmaf_rn(l2, 0.f, nf2); nf3 = __fmaf_rn(l3, 0.f, nf3);
                                } else {
                                    raw[t0 + PM_CONV * j] = make_uint4(0u, 0u, 0u, 0u);   // rows past the segment never reach the tensor core
                                }
                                // lo tile: [64-channel block j >> 3][row][16-byte chunk ^ (row & 7)][8-byte half]
                                *reinterpret_cast<uint2 *>(lo + (((j >> 2) & 1) ? lo_odd : lo_even) + (j >> 3) * (PM_TP * 128) +
                                                           (j & 3) * 2048) = l;
                            }
                        }
                    };
                    if (rows_valid == PM_TP) convert(std::true_type{}); else convert(std::false_type{});
                    c0_ = PM_CLK();
                    fence_proxy_async();
                    PM_ACC(tw2, c0_);
                    mbar_arrive(&s_conv_full[cs]);
                    if (++rs == P.raw_stages) { rs = 0; rph ^= 1; }
                    if (++cs == P2_CS) { cs = 0; cph ^= 1; }
                }
                if (!((nf0 + nf1) + (nf2 + nf3) == 0.f) && P.status) atomicOr(P.status, XM3D_FLAG_NONFINITE);
                if (t0 == 0) PM_OUT(2, tw0, tw1, PM_CLK() - tall, tw2);
            }
        } else if (warp < 10) {
            int iq = 0; uint32_t iqph = 0; int ib = 0; uint32_t ibph = 0; int ts = 0; uint32_t tph = 0;
            for (;;) {
                const P2Item it = p2_take_item(s_iq, s_iq_a, s_iq_full, s_iq_free, iq, iqph, lane);
                if (it.n < 0) break;
                const int s = it.s, sl = it.sl, n = it.n; const int64_t a = it.a;
                const int ntile = (n + PM_TP - 1) / PM_TP;
                (void)s; (void)sl; (void)a;
                // ===== builders: lane = mask.  32 x 32 bit transposes (ballots) of the tile's membership words, expanded in
                // registers to 0.0f / 1.0f (tf32) and to bf16 pairs, stored into tensor memory =====
                // M = 128: TMEM lane = mask, lane group lg holds masks 32 lg .. 32 lg + 31 = membership word lg.
                // M = 64:  rows 16 q .. 16 q + 15 live in lanes 0..15 of lane group q: this warp's masks are 16 lg .. 16 lg + 15,
                //          bits (lg & 1) * 16 .. of membership word lg >> 1.
                const int lg = warp & 3;
                const int wsel = P.m64 ? (lg >> 1) : lg;
                const int tail = P.k & 31;
                const bool word_ok = wsel < P.words && wsel * 32 < P.k;
                int cnt = 0, cnt_b = 0;                                // (M = 64: rows t/4 and t/4 + 8 of this warp)
                long long tw0 = 0, tw1 = 0, tw2 = 0, tw3 = 0, tall = PM_CLK();
                for (int t = 0; t < ntile; ++t) {
                    // The transposed membership words (row = mask, bit = point) come from the two preparation warps through a
                    // shared-memory ring: loading and transposing them here cost 600 of this warp's serial ~1 500 cycles per
                    // tile, and this warp's chain (rows -> expansion -> tensor-memory stores) paces the MMAs.
                    long long c1_ = PM_CLK();
                    mbar_wait(&s_tw_full[ts], tph);
                    PM_ACC(tw3, c1_);
                    uint32_t m0 = 0u, m1 = 0u, ra0 = 0u, ra1 = 0u, rb0 = 0u, rb1 = 0u;
                    if (P.m64) {
                        // M = 64: the 16 mask rows of this warp go to lanes 0..15 of its lane group with the 16-lane store
                        // shape (thread t: rows t/4 and t/4 + 8, columns 8 i + 2 (t % 4) + {0, 1}) — half the tensor-memory
                        // stores and expansion work of the 32-lane shape, which would write 16 unused lanes.
                        if (word_ok) {
                            const int row = (lg & 1) * 16 + (lane >> 2);
                            ra0 = s_tw[ts][wsel][0][row]; rb0 = s_tw[ts][wsel][0][row + 8];
                            ra1 = s_tw[ts][wsel][1][row]; rb1 = s_tw[ts][wsel][1][row + 8];
                        }
                        cnt += __popc(ra0) + __popc(ra1);
                        cnt_b += __popc(rb0) + __popc(rb1);
                    } else {
                        if (word_ok) { m0 = s_tw[ts][wsel][0][lane]; m1 = s_tw[ts][wsel][1][lane]; }
                        cnt += __popc(m0) + __popc(m1);
                    }
                    mbar_arrive(&s_tw_empty[ts]);                     // (release: the reads above are ordered before it)
                    if (++ts == P2_TW) { ts = 0; tph ^= 1; }
                    long long c0_ = PM_CLK();
                    mbar_wait(&s_conv_empty[cs], cph ^ 1);
                    PM_ACC(tw0, c0_);
                    c0_ = PM_CLK();
                    tc_fence_after();
                    const uint32_t lane_addr = (uint32_t)(lg * 32) << 16;
                    const uint32_t a32 = tmem_base + lane_addr + (uint32_t)(P2_A32 + cs * 64);
                    const uint32_t a16 = tmem_base + lane_addr + (uint32_t)(P2_A16 + cs * 32);
                    if (P.m64) {
                        // one AND + one multiply per value: (w & (1 << b)) * (ONE >> b) is ONE or 0 exactly (b <= 23 for
                        // 0x3f800000 = 0x7f << 23, b <= 7 for 0x3f80 = 0x7f << 7; higher bits come from a shifted copy)
                        {   // tf32 0.0f / 1.0f: column = point
                            const int sh = 2 * (lane & 3);
                            const uint32_t a_lo = ra0 >> sh, a_hi = ra1 >> sh, b_lo = rb0 >> sh, b_hi = rb1 >> sh;
                            const uint32_t a_lo8 = a_lo >> 8, a_hi8 = a_hi >> 8, b_lo8 = b_lo >> 8, b_hi8 = b_hi >> 8;
                            uint32_t v[32];
#pragma unroll
                            for (int i = 0; i < 8; ++i) {
                                const bool up = (i & 3) == 3;                       // bits 24, 25 -> bits 16, 17 of the >> 8 copy
                                const uint32_t wa = i < 4 ? (up ? a_lo8 : a_lo) : (up ? a_hi8 : a_hi);
                                const uint32_t wb = i < 4 ? (up ? b_lo8 : b_lo) : (up ? b_hi8 : b_hi);
                                const int bit = up ? 16 : 8 * (i & 3);
                                v[4 * i + 0] = (wa & (1u << bit)) * (0x3f800000u >> bit);
                                v[4 * i + 1] = (wa & (2u << bit)) * (0x3f800000u >> (bit + 1));
                                v[4 * i + 2] = (wb & (1u << bit)) * (0x3f800000u >> bit);
                                v[4 * i + 3] = (wb & (2u << bit)) * (0x3f800000u >> (bit + 1));
                            }
                            tmem_st_16x256b_x8(a32, v);
                        }
                        {   // bf16 pairs: column c = points 2 c (low half), 2 c + 1
                            const int sh = 4 * (lane & 3);
                            const uint32_t a_lo = ra0 >> sh, a_hi = ra1 >> sh, b_lo = rb0 >> sh, b_hi = rb1 >> sh;
                            const uint32_t a_lo16 = a_lo >> 16, a_hi16 = a_hi >> 16, b_lo16 = b_lo >> 16, b_hi16 = b_hi >> 16;
                            uint32_t v[16];
#pragma unroll
                            for (int i = 0; i < 4; ++i) {
                                const uint32_t wa = i < 2 ? ((i & 1) ? a_lo16 : a_lo) : ((i & 1) ? a_hi16 : a_hi);
                                const uint32_t wb = i < 2 ? ((i & 1) ? b_lo16 : b_lo) : ((i & 1) ? b_hi16 : b_hi);
#pragma unroll
                                for (int e = 0; e < 2; ++e) {
                                    const int bit = 2 * e;
                                    v[4 * i + e] = (wa & (1u << bit)) * (0x3f80u >> bit) + (wa & (2u << bit)) * (0x3f800000u >> (bit + 1));
                                    v[4 * i + 2 + e] = (wb & (1u << bit)) * (0x3f80u >> bit) + (wb & (2u << bit)) * (0x3f800000u >> (bit + 1));
                                }
                            }
                            tmem_st_16x256b_x4(a16, v);
                        }
                    } else {
                        // M = 128: lane = mask row, 32-lane store shape; same AND + multiply expansion
                        const uint32_t m0_8 = m0 >> 8, m1_8 = m1 >> 8;
#pragma unroll
                        for (int h = 0; h < 4; ++h) {
                            uint32_t v[16];
#pragma unroll
                            for (int j = 0; j < 16; ++j) {
                                const int b = (h & 1) * 16 + j;                      // bit of m0 / m1
                                const uint32_t w = b < 24 ? (h < 2 ? m0 : m1) : (h < 2 ? m0_8 : m1_8);
                                const int bb = b < 24 ? b : b - 8;
                                v[j] = (w & (1u << bb)) * (0x3f800000u >> bb);
                            }
                            tmem_st16(a32 + h * 16, v);
                        }
#pragma unroll
                        for (int h = 0; h < 2; ++h) {
                            const uint32_t bits = h ? m1 : m0;
                            uint32_t v[16];
#pragma unroll
                            for (int j = 0; j < 16; ++j) {
                                const uint32_t w = bits >> (8 * (j >> 2));           // points 2 j, 2 j + 1 = bits r, r + 1 of this copy
                                const int r = 2 * (j & 3);
                                v[j] = (w & (1u << r)) * (0x3f80u >> r) + (w & (2u << r)) * (0x3f800000u >> (r + 1));
                            }
                            tmem_st16(a16 + h * 16, v);
                        }
                    }
                    PM_ACC(tw1, c0_); c0_ = PM_CLK();
                    tmem_st_wait();
                    tc_fence_before();
                    PM_ACC(tw2, c0_);
                    mbar_arrive(&s_conv_full[cs]);
                    if (++cs == P2_CS) { cs = 0; cph ^= 1; }
                }
                mbar_wait(&s_cnt_free[ib], ibph ^ 1);                  // the epilogue has read the counts of two items ago
                if (!P.m64) s_cnt[ib][lg * 32 + lane] = cnt;          // exactly one thread per mask
                else if ((lane & 3) == 0) { s_cnt[ib][lg * 16 + (lane >> 2)] = cnt; s_cnt[ib][lg * 16 + (lane >> 2) + 8] = cnt_b; }
                mbar_arrive(&s_cnt_full[ib]);
                if (tid == 192) { PM_OUT(3, tw0, tw1, PM_CLK() - tall, tw2); PM_OUT(5, tw3, 0, 0, 0); }
                if (++ib == 2) { ib = 0; ibph ^= 1; }
            }
        } else if (warp < 10 + P2_EPI_WARPS) {
            int iq = 0; uint32_t iqph = 0; int ib = 0; uint32_t ibph = 0;
            for (;;) {
                const P2Item it = p2_take_item(s_iq, s_iq_a, s_iq_full, s_iq_free, iq, iqph, lane);
                if (it.n < 0) break;
                const int s = it.s, sl = it.sl, n = it.n; const int64_t a = it.a;
                const int ntile = (n + PM_TP - 1) / PM_TP;
                (void)s; (void)sl; (void)a;
                // ===== epilogue: lane = mask, 64 channels per warp =====
                const int lg = warp & 3, half = (warp - 10) >> 2;
                float acc[64];
#pragma unroll
                for (int j = 0; j < 64; ++j) acc[j] = 0.f;
                long long tw0 = 0, tall = PM_CLK();
                for (int g = 0; g < (ntile + P2_GROUP - 1) / P2_GROUP; ++g) {
                    long long c0_ = PM_CLK();
                    mbar_wait_sleep(&s_tile_done[buf], bph, 20);
                    PM_ACC(tw0, c0_);
                    tc_fence_after();
                    const uint32_t taddr = tmem_base + ((uint32_t)(lg * 32) << 16) + (uint32_t)(P2_D + buf * PM_SLICE + half * 64);
#pragma unroll
                    for (int c0 = 0; c0 < 64; c0 += 32) {
                        uint32_t v0[16], v1[16];
                        tmem_ld16_nowait(taddr + c0, v0);
                        tmem_ld16_nowait(taddr + c0 + 16, v1);
                        tmem_ld_wait();
#pragma unroll
                        for (int j = 0; j < 16; ++j) {
                            acc[c0 + j] = __fadd_rn(acc[c0 + j], __uint_as_float(v0[j]));
                            acc[c0 + 16 + j] = __fadd_rn(acc[c0 + 16 + j], __uint_as_float(v1[j]));
                        }
                    }
                    tc_fence_before();
                    mbar_arrive(&s_tmem_free[buf]);
                    if (++buf == P2_DB) { buf = 0; bph ^= 1; }
                }
                if (tid == 320) PM_OUT(4, tw0, 0, PM_CLK() - tall, 0);
                mbar_wait(&s_cnt_full[ib], ibph);                      // the builders' counts of this item
                const int m = P.m64 ? (lane < 16 ? lg * 16 + lane : P.k) : lg * 32 + lane;
                const int nm = m < P.k ? s_cnt[ib][m] : 0;
                mbar_arrive(&s_cnt_free[ib]);
                if (m < P.k) {
                    const size_t o = ((size_t)s * P.k + m) * P.c + (size_t)sl * PM_SLICE + half * 64;
                    float4 *so = reinterpret_cast<float4 *>(P.sum + o);
                    float4 *mo = P.mean ? reinterpret_cast<float4 *>(P.mean + o) : nullptr;
                    const float d = (float)nm;
#pragma unroll
                    for (int j = 0; j < 16; ++j) {
                        const float4 v = make_float4(acc[4 * j], acc[4 * j + 1], acc[4 * j + 2], acc[4 * j + 3]);
                        so[j] = v;
                        if (mo) mo[j] = nm > 0 ? make_float4(__fdiv_rn(v.x, d), __fdiv_rn(v.y, d), __fdiv_rn(v.z, d), __fdiv_rn(v.w, d))
                                               : make_float4(0.f, 0.f, 0.f, 0.f);
                    }
                    if (P.cnt && sl == 0 && half == 0) P.cnt[s * P.k + m] = nm;
                }
                if (++ib == 2) { ib = 0; ibph ^= 1; }
            }
        } else {
            // ===== preparation warps: membership words -> transposed rows (row = mask, bit = point) in shared memory =====
            // Warp w handles membership words w and w + 2.  The words of a tile arrive through an asynchronous copy ring P2_MW
            // tiles ahead (a register prefetch that is rotated at the end of the iteration makes the rotation wait for the
            // load: the pipeline then runs at one loaded global-memory latency per tile).
            const int pw = warp - (10 + P2_EPI_WARPS);
            const bool second_word = pw + 2 < P.words && (pw + 2) * 32 < P.k;      // K > 64: this warp transposes two words
            const int tail = P.k & 31;
            int iq = 0; uint32_t iqph = 0; int ts = 0; uint32_t tph = 0;
            for (;;) {
                const P2Item it = p2_take_item(s_iq, s_iq_a, s_iq_full, s_iq_free, iq, iqph, lane);
                if (it.n < 0) break;
                const int n = it.n; const int64_t a = it.a;
                const int ntile = (n + PM_TP - 1) / PM_TP;
                auto issue_words = [&](int t) {
#pragma unroll
                    for (int wq = 0; wq < 2; ++wq) {
                        const int w = pw + 2 * wq;
                        const bool word_ok = w < P.words && w * 32 < P.k;
                        uint32_t *dst = &s_mw[t & (P2_MW - 1)][w][lane];
#pragma unroll
                        for (int h = 0; h < 2; ++h) {
                            const int pt = t * PM_TP + 32 * h + lane;
                            const bool ok = word_ok && t < ntile && pt < n;
                            const uint32_t *src = ok ? P.member + (size_t)(a + pt) * P.words + w : P.member;
                            if (word_ok)
                                asm volatile("cp.async.ca.shared.global [%0], [%1], 4, %2;" ::"r"(smem_u32(dst + 32 * h)), "l"(src),
                                             "r"(ok ? 4 : 0) : "memory");
                        }
                    }
                    asm volatile("cp.async.commit_group;" ::: "memory");
                };
#pragma unroll
                for (int u = 0; u < P2_MW; ++u) issue_words(u);
                long long pw0 = 0, pw1 = 0, pw2 = 0, pw3 = 0;
                for (int t = 0; t < ntile; ++t) {
                    long long c0_ = PM_CLK();
                    asm volatile("cp.async.wait_group %0;" ::"n"(P2_MW - 1) : "memory");
                    PM_ACC(pw0, c0_); c0_ = PM_CLK();
                    // the (up to) four 32 x 32 bit transposes of this warp's two words run round by round together
                    uint32_t tr[4];
#pragma unroll
                    for (int wq = 0; wq < 2; ++wq) {
                        const int w = pw + 2 * wq;
                        const bool word_ok = w < P.words && w * 32 < P.k;
                        const uint32_t tail_mask = (tail && w == (P.k >> 5)) ? (1u << tail) - 1u : 0xffffffffu;
                        tr[2 * wq] = word_ok ? s_mw[t & (P2_MW - 1)][w][lane] & tail_mask : 0u;
                        tr[2 * wq + 1] = word_ok ? s_mw[t & (P2_MW - 1)][w][32 + lane] & tail_mask : 0u;
                    }
                    if (second_word) {
                        warp_transpose32_n<4>(tr);
                    } else {
                        uint32_t t2[2] = {tr[0], tr[1]};
                        warp_transpose32_n<2>(t2);
                        tr[0] = t2[0]; tr[1] = t2[1];
                    }
                    PM_ACC(pw1, c0_); c0_ = PM_CLK();
                    issue_words(t + P2_MW);                           // the slot is free: its words are in registers
                    PM_ACC(pw2, c0_); c0_ = PM_CLK();
                    mbar_wait(&s_tw_empty[ts], tph ^ 1);
                    PM_ACC(pw3, c0_);
#pragma unroll
                    for (int wq = 0; wq < 2; ++wq) {
                        s_tw[ts][pw + 2 * wq][0][lane] = tr[2 * wq];
                        s_tw[ts][pw + 2 * wq][1][lane] = tr[2 * wq + 1];
                    }
                    mbar_arrive(&s_tw_full[ts]);
                    if (++ts == P2_TW) { ts = 0; tph ^= 1; }
                }
                if (tid == 576) PM_OUT(6, pw0, pw1, pw2, pw3);
                asm volatile("cp.async.wait_group 0;" ::: "memory");   // the copy ring is reused by the next item
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        __syncwarp();
        tmem_dealloc(tmem_base, 512);
    }
}

#ifdef XM3D_PM_TIMING
extern "C" __attribute__((visibility("default"))) int xm3d_pool_mma_debug(void *dev_buf) {
    return (int)cudaMemcpyToSymbol(g_pm_dbg, &dev_buf, sizeof(void *));
}
#endif

// Host side -----------------------------------------------------------------------------------------
bool pool_mma_eligible(const float *feat, int c, const int32_t *row_index, const uint32_t *member, int k, int64_t cap,
                       const float *sum, const float *mean) {
    return member && !row_index && c % PM_SLICE == 0 && k >= 1 && k <= 128 && cap > 0 &&
           cap < ((int64_t)1 << 31) - PM_TP && reinterpret_cast<uintptr_t>(feat) % 16 == 0 && sum &&
           reinterpret_cast<uintptr_t>(sum) % 16 == 0 && (!mean || reinterpret_cast<uintptr_t>(mean) % 16 == 0);
}

int launch_pool_mma(const float *feat, int c, const uint32_t *member, int words, int n_seg, int k, const int64_t *seg_off,
                    int64_t cap, float *sum, int32_t *cnt, float *mean, int *work, int *order, int tune, int32_t *status,
                    cudaStream_t stream) {
    CUtensorMap map;
    memset(&map, 0, sizeof(map));
    // rows past `cap` are zero-filled by TMA; rows past a segment are zeroed by the converters
    if (!make_map_sw128(&map, feat, cap, c, 32, PM_TP, /*atom32=*/true)) {
        set_error("xm3d_pool_batch: cuTensorMapEncodeTiled failed");
        return XM3D_ERR_CUDA;
    }
    PoolMmaParams P;
    P.member = member; P.seg_off = seg_off; P.words = words; P.k = k; P.n_seg = n_seg; P.c = c; P.cap = cap;
    P.sum = sum; P.mean = mean; P.cnt = cnt; P.work = work; P.dbg = (tune >> 14) & 3; P.dbg2 = (tune >> 16) & 7; P.status = status;
    P.m64 = (k <= 64 && !((tune >> 10) & 1)) ? 1 : 0;      // (bit 10: experiments with M = 128 for every k)
    const int n_items = n_seg * (c / PM_SLICE);
    const unsigned grid = (unsigned)(n_items < sm_count() ? n_items : sm_count());
    cudaMemsetAsync(work, 0, sizeof(int), stream);
    P.order = nullptr;
    if (order && n_seg > 1 && n_seg <= 4096 && !((tune >> 9) & 1)) {      // (bit 9: experiments without the ordering)
        pool_order_kernel<<<(unsigned)((n_seg + 255) / 256), 256, 0, stream>>>(seg_off, n_seg, order);
        count_launches(1);
        P.order = order;
    }
    static std::atomic<uint64_t> attr_set{0};
    if (first_use_on_device(&attr_set))
        cudaFuncSetAttribute(pool_mma2_kernel<16>, cudaFuncAttributeMaxDynamicSharedMemorySize, P2_MAX_DYN_SMEM);
    const int t_rs = tune & 15;                            // experiments: depth of the raw ring (default 4)
    P.raw_stages = (t_rs >= 1 && t_rs <= PM_MAX_STAGES) ? t_rs : 4;
    P.conv_stages = P2_CS;
    const size_t smem2 = (size_t)P.raw_stages * PM_RAW + P2_CS * (size_t)P2_LO + 1024;
    pool_mma2_kernel<16><<<grid, P2_THREADS, smem2, stream>>>(map, P);      // 16 converter loads in flight (8: +2 %, 4: +7 %)
    count_launches(1);
    return check_launch("xm3d_pool_batch (tensor-core path)");
}

}  // namespace xm3d
