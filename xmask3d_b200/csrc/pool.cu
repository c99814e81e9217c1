// Stage 3 — segmented mean pooling of per-point features under each mask.
//
// Replaces the inline pooling of the reference: models/utils/criterion.py:148-157
// (`feature_3d[mask_3d[k]]` -> mean(0)), the scalar form models/xmask3d.py:362-367 (c = 1) and the
// per-scene global mean models/xmask3d.py:239-258 (one all-ones mask).
//
//   sum[s,m,:] = sum of feat[row(i),:] over points i of segment s inside mask m ;  cnt[s,m]
//
// This is the dominant HBM stream of the whole path (n x C float32, read once).  Design:
//   count   one CTA per tile of 1024 points: members of every mask in the tile (warp ballots)
//   prefix  per (segment, mask): exclusive prefix of the tile counts, unit totals cnt[s][m]
//   scan    one CTA: exclusive prefixes -> pair offsets and chunk offsets of every (segment, mask)
//   fill    one CTA per tile: STABLE placement (ballot ranks inside a warp, byte counters across
//           warps, tile prefixes across tiles) -> perm[] = feature row of every (point, mask)
//           pair, masks ascending, points ascending inside a mask
//   sum     one CTA per chunk of <= 256 pairs of one (segment, mask): each thread owns VEC channels
//           and accumulates the chunk's rows in REGISTERS — no atomics, no shared-memory traffic;
//           rows are fetched 8 at a time with 16-byte L1-bypassing loads and many CTAs are
//           resident per SM, so the memory pipeline is kept full by thread-level parallelism
//   combine fixed-order sum of the chunk partials of each (segment, mask) -> sum, cnt, mean
// The result is deterministic (no float atomics; every order is fixed by the point order).
// A first version kept all k accumulator rows of a persistent CTA in shared memory and streamed
// points in storage order; it reached only 16 % of the HBM roofline on B200 (one CTA per SM, and
// register-ring prefetch deeper than the six scoreboard slots does not overlap) — see DESIGN.md.
#include <cuda.h>
#include <stdlib.h>
#include <string.h>

#include "common.cuh"
#include "vec.cuh"

namespace xm3d {

#ifndef XM3D_POOL_CH
#define XM3D_POOL_CH 256
#endif
constexpr int POOL_CH = XM3D_POOL_CH;        // pairs per chunk (one partial row each)
constexpr int POOL_UNROLL = 8;      // rows in flight per thread in the sum kernel
constexpr int FILL_THREADS = 1024;

struct PoolIdx {
    const uint32_t *member;   // [cap, words] or null
    const int32_t *label;     // [cap] or null
    const int32_t *row_index; // [cap] or null
    const int64_t *seg_off;
    int n_seg, k, words;
    int64_t cap;
};

// membership words of point i (label mode: a single bit, or none)
template <int W>
__device__ __forceinline__ void load_bits(const PoolIdx &P, int64_t i, bool valid, uint32_t (&b)[W]) {
#pragma unroll
    for (int w = 0; w < W; ++w) b[w] = 0u;
    if (!valid) return;
    if (P.label) {
        const int m = __ldg(P.label + i);
        if (m >= 0 && m < P.k) {
#pragma unroll
            for (int w = 0; w < W; ++w)
                if ((m >> 5) == w) b[w] = 1u << (m & 31);
        }
    } else {
#pragma unroll
        for (int w = 0; w < W; ++w)
            if (w < P.words) b[w] = __ldg(P.member + i * P.words + w);
        // ignore bits of masks >= k
        const int tail = P.k & 31;
#pragma unroll
        for (int w = 0; w < W; ++w)
            if (tail && w == P.words - 1) b[w] &= (1u << tail) - 1u;
    }
}

// tiles: every segment is cut into tiles of FILL_THREADS points; tile_off[s] = first tile of segment s
__global__ void __launch_bounds__(1024, 1)
pool_tileplan_kernel(const int64_t *__restrict__ seg_off, int n_seg, int64_t cap, int32_t *__restrict__ tile_off) {
    __shared__ int s_w[32];
    __shared__ int s_carry;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const bool over = seg_off[n_seg] > cap;
    if (tid == 0) s_carry = 0;
    __syncthreads();
    for (int base = 0; base < n_seg; base += 1024) {
        const int s = base + tid;
        const int val = (s < n_seg && !over) ? (int)((seg_off[s + 1] - seg_off[s] + FILL_THREADS - 1) / FILL_THREADS) : 0;
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
        if (s < n_seg) tile_off[s] = excl;
        __syncthreads();
        if (tid == 1023) s_carry = excl + val;
        __syncthreads();
    }
    if (tid == 0) tile_off[n_seg] = s_carry;
}

// Per-warp member counts of one tile: s_wtot[warp][mask] (bytes, <= 32 each).  All threads call it.
// `single`: every lane of the warp is in at most one mask (partition masks / labels) — warp-uniform.
// Then one match.any groups the lanes by mask and replaces the per-mask ballot loop.
template <int W>
__device__ __forceinline__ bool warp_single_membership(const uint32_t (&b)[W]) {
    int pc = 0;
#pragma unroll
    for (int w = 0; w < W; ++w) pc += __popc(b[w]);
    return __all_sync(0xffffffffu, pc <= 1);
}
template <int W>
__device__ __forceinline__ int single_mask_id(const uint32_t (&b)[W]) {      // -1: in no mask
    int m = -1;
#pragma unroll
    for (int w = 0; w < W; ++w)
        if (b[w]) m = w * 32 + __ffs(b[w]) - 1;
    return m;
}

template <int W>
__device__ __forceinline__ void tile_warp_counts(const uint32_t (&b)[W], uint32_t (&uni)[W],
                                                 unsigned char (*s_wtot)[32 * W], int warp, int lane) {
    if (warp_single_membership<W>(b)) {
        const int m = single_mask_id<W>(b);
        const unsigned grp = __match_any_sync(0xffffffffu, m);
        if (m >= 0 && lane == __ffs(grp) - 1) s_wtot[warp][m] = (unsigned char)__popc(grp);
#pragma unroll
        for (int w = 0; w < W; ++w) uni[w] = 0u;            // unused on this path
        return;
    }
#pragma unroll
    for (int w = 0; w < W; ++w) {
        uni[w] = __reduce_or_sync(0xffffffffu, b[w]);
        uint32_t u = uni[w];
        while (u) {
            const int bit = __ffs(u) - 1;
            u &= u - 1;
            const unsigned vote = __ballot_sync(0xffffffffu, (b[w] >> bit) & 1u);
            if (lane == 0) s_wtot[warp][w * 32 + bit] = (unsigned char)__popc(vote);
        }
    }
}

__device__ __forceinline__ int tile_segment(const int32_t *__restrict__ tile_off, int n_seg, int tile) {
    int lo = 0, hi = n_seg;        // largest s with tile_off[s] <= tile
    while (hi - lo > 1) {
        const int mid = (lo + hi) >> 1;
        if (tile_off[mid] <= tile) lo = mid; else hi = mid;
    }
    return lo;
}

// one CTA per tile: members of every mask inside the tile -> tile_cnt[tile][k]
template <int W>
__global__ void __launch_bounds__(FILL_THREADS, 1)
pool_tilecount_kernel(const PoolIdx P, const int32_t *__restrict__ tile_off, int32_t *__restrict__ tile_cnt) {
    __shared__ unsigned char s_wtot[32][32 * W];
    __shared__ int s_seg;
    const int tile = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (tile >= tile_off[P.n_seg]) return;
    if (tid == 0) s_seg = tile_segment(tile_off, P.n_seg, tile);
    for (int j = tid; j < 32 * 32 * W / 4; j += FILL_THREADS) reinterpret_cast<uint32_t *>(&s_wtot[0][0])[j] = 0u;
    __syncthreads();
    const int s = s_seg;
    const int64_t e = P.seg_off[s + 1];
    const int64_t i = P.seg_off[s] + (int64_t)(tile - tile_off[s]) * FILL_THREADS + tid;
    uint32_t b[W], uni[W];
    load_bits<W>(P, i, i < e, b);
    tile_warp_counts<W>(b, uni, s_wtot, warp, lane);
    __syncthreads();
    if (tid < P.k) {
        int run = 0;
#pragma unroll 8
        for (int w = 0; w < 32; ++w) run += s_wtot[w][tid];
        tile_cnt[(size_t)tile * P.k + tid] = run;
    }
}

// one CTA per segment, one thread per mask: exclusive prefix of the tile counts along the
// segment's tiles (tile_cnt is overwritten with it) and the unit totals cnt[s][m]
__global__ void __launch_bounds__(256)
pool_unitprefix_kernel(const int32_t *__restrict__ tile_off, int k, int32_t *__restrict__ tile_cnt,
                       int32_t *__restrict__ cnt) {
    const int s = blockIdx.x;
    const int t0 = tile_off[s], t1 = tile_off[s + 1];
    for (int m = threadIdx.x; m < k; m += blockDim.x) {
        int run = 0;
        for (int t = t0; t < t1; ++t) {
            const int v = tile_cnt[(size_t)t * k + m];
            tile_cnt[(size_t)t * k + m] = run;
            run += v;
        }
        cnt[(size_t)s * k + m] = run;
    }
}

// exclusive prefixes over the (segment, mask) units: pair offsets and chunk offsets
__global__ void __launch_bounds__(1024, 1)
pool_scan_kernel(const int32_t *__restrict__ cnt, int n_units, int64_t cap_pairs, int64_t *__restrict__ pair_off,
                 int32_t *__restrict__ chunk_off, int32_t *status) {
    __shared__ int64_t s_wp[32];
    __shared__ int s_wc[32];
    __shared__ int64_t s_cp;
    __shared__ int s_cc;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (tid == 0) { s_cp = 0; s_cc = 0; }
    __syncthreads();
    for (int base = 0; base < n_units; base += 1024) {
        const int u = base + tid;
        const int64_t vp = u < n_units ? cnt[u] : 0;
        const int vc = (int)((vp + POOL_CH - 1) / POOL_CH);
        int64_t ip = vp;
        int ic = vc;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int64_t tp = __shfl_up_sync(0xffffffffu, ip, o);
            const int tc = __shfl_up_sync(0xffffffffu, ic, o);
            if (lane >= o) { ip += tp; ic += tc; }
        }
        if (lane == 31) { s_wp[warp] = ip; s_wc[warp] = ic; }
        __syncthreads();
        if (warp == 0) {
            const int64_t wp = s_wp[lane];
            const int wc = s_wc[lane];
            int64_t xp = wp;
            int xc = wc;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const int64_t tp = __shfl_up_sync(0xffffffffu, xp, o);
                const int tc = __shfl_up_sync(0xffffffffu, xc, o);
                if (lane >= o) { xp += tp; xc += tc; }
            }
            s_wp[lane] = xp - wp;
            s_wc[lane] = xc - wc;
        }
        __syncthreads();
        const int64_t ep = s_cp + s_wp[warp] + ip - vp;
        const int ec = s_cc + s_wc[warp] + ic - vc;
        if (u < n_units) { pair_off[u] = ep; chunk_off[u] = ec; }
        __syncthreads();
        if (tid == 1023) { s_cp = ep + vp; s_cc = ec + vc; }
        __syncthreads();
    }
    if (tid == 0) {
        pair_off[n_units] = s_cp;
        chunk_off[n_units] = s_cc;
        if (s_cp > cap_pairs && status) atomicOr(status, XM3D_FLAG_PAIR_OVERFLOW);
    }
}

// one CTA per tile: STABLE placement of the tile's (point, mask) pairs:
// perm[pair_off[s][m] + (members of m in earlier tiles) + (members in earlier warps / lanes)] = row
template <int W>
__global__ void __launch_bounds__(FILL_THREADS, 1)
pool_fill_kernel(const PoolIdx P, const int32_t *__restrict__ tile_off, const int32_t *__restrict__ tile_pre,
                 const int64_t *__restrict__ pair_off, int64_t cap_pairs, int32_t *__restrict__ perm) {
    __shared__ unsigned char s_wtot[32][32 * W];     // per warp, per mask: members in this tile
    __shared__ unsigned short s_wpre[32][32 * W];    // exclusive prefix over warps
    __shared__ int64_t s_base[32 * W];               // first pair slot of this tile, per mask
    __shared__ int s_seg;
    const int tile = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int k = P.k;
    if (tile >= tile_off[P.n_seg] || pair_off[(size_t)P.n_seg * k] > cap_pairs) return;     // overflow: flagged by the scan
    if (tid == 0) s_seg = tile_segment(tile_off, P.n_seg, tile);
    for (int j = tid; j < 32 * 32 * W / 4; j += FILL_THREADS) reinterpret_cast<uint32_t *>(&s_wtot[0][0])[j] = 0u;
    __syncthreads();
    const int s = s_seg;
    const int64_t e = P.seg_off[s + 1];
    const int64_t i = P.seg_off[s] + (int64_t)(tile - tile_off[s]) * FILL_THREADS + tid;
    uint32_t b[W], uni[W];
    load_bits<W>(P, i, i < e, b);
    tile_warp_counts<W>(b, uni, s_wtot, warp, lane);
    __syncthreads();
    if (tid < k) {
        int run = 0;
#pragma unroll 8
        for (int w = 0; w < 32; ++w) {
            s_wpre[w][tid] = (unsigned short)run;
            run += s_wtot[w][tid];
        }
        s_base[tid] = pair_off[(size_t)s * k + tid] + tile_pre[(size_t)tile * k + tid];
    }
    __syncthreads();
    const int row = (i < e) ? (P.row_index ? __ldg(P.row_index + i) : (int)i) : 0;
    if (warp_single_membership<W>(b)) {
        const int m = single_mask_id<W>(b);
        const unsigned grp = __match_any_sync(0xffffffffu, m);
        if (m >= 0) perm[s_base[m] + s_wpre[warp][m] + __popc(grp & ((1u << lane) - 1u))] = row;
        return;
    }
#pragma unroll
    for (int w = 0; w < W; ++w) {
        uint32_t u = uni[w];
        while (u) {
            const int bit = __ffs(u) - 1;
            u &= u - 1;
            const unsigned vote = __ballot_sync(0xffffffffu, (b[w] >> bit) & 1u);
            if ((b[w] >> bit) & 1u) {
                const int m = w * 32 + bit;
                perm[s_base[m] + s_wpre[warp][m] + __popc(vote & ((1u << lane) - 1u))] = row;
            }
        }
    }
}

struct SumParams {
    const float *feat;
    int c;
    const int32_t *perm;
    const int64_t *pair_off;
    const int32_t *chunk_off;
    int n_units;
    int64_t cap_pairs;
    float *partial;            // [chunks, c]
    // units made of exactly one chunk are finished here (no partial row, no combine pass)
    const int32_t *cnt_in;     // [n_units] pairs per unit
    float *sum, *mean;         // [n_units, c]; mean may be null
    int32_t *cnt;              // [n_units] or null
    int direct;                // 1: sum / mean are 16-byte aligned for the vector store
};

template <int VEC>
__global__ void __launch_bounds__(1024) pool_sum_kernel(const SumParams P) {
    using V = typename VecT<VEC>::type;
    __shared__ int s_unit;
    const int chunk = blockIdx.x, tid = threadIdx.x, lane = tid & 31;
    if (chunk >= P.chunk_off[P.n_units] || P.pair_off[P.n_units] > P.cap_pairs) return;
    if (tid == 0) {
        int lo = 0, hi = P.n_units;        // largest u with chunk_off[u] <= chunk
        while (hi - lo > 1) {
            const int mid = (lo + hi) >> 1;
            if (P.chunk_off[mid] <= chunk) lo = mid; else hi = mid;
        }
        s_unit = lo;
    }
    __syncthreads();
    const int u = s_unit;
    const int64_t begin = P.pair_off[u] + (int64_t)(chunk - P.chunk_off[u]) * POOL_CH;
    const int64_t uend = P.pair_off[u + 1];
    const int n = (int)((uend - begin) < POOL_CH ? (uend - begin) : POOL_CH);
    const int ch = tid * VEC;
    const bool active = ch < P.c;
    const float *fb = P.feat + ch;
    V acc; vzero(acc);
    for (int r0 = 0; r0 < n; r0 += 32) {
        const int myrow = (r0 + lane < n) ? __ldg(P.perm + begin + r0 + lane) : 0;     // coalesced, then broadcast
        const int nb = (n - r0) < 32 ? (n - r0) : 32;
        for (int j0 = 0; j0 < nb; j0 += POOL_UNROLL) {
            V buf[POOL_UNROLL];
#pragma unroll
            for (int j = 0; j < POOL_UNROLL; ++j) {
                const int row = __shfl_sync(0xffffffffu, myrow, (j0 + j) & 31);
                if (active && j0 + j < nb) buf[j] = ld_stream<VEC>(fb + (size_t)row * P.c);
            }
#pragma unroll
            for (int j = 0; j < POOL_UNROLL; ++j)
                if (active && j0 + j < nb) vadd(acc, buf[j]);          // fixed row order
        }
    }
    if (P.direct && P.chunk_off[u + 1] - P.chunk_off[u] == 1) {
        const size_t o = (size_t)u * P.c + ch;
        if (active) {
            *reinterpret_cast<V *>(P.sum + o) = acc;
            if (P.mean) {
                float *a = reinterpret_cast<float *>(&acc);
#pragma unroll
                for (int j = 0; j < VEC; ++j) a[j] = __fdiv_rn(a[j], (float)n);      // n >= 1 here
                *reinterpret_cast<V *>(P.mean + o) = acc;
            }
        }
        if (P.cnt && tid == 0) P.cnt[u] = n;
        return;
    }
    if (active) *reinterpret_cast<V *>(P.partial + (size_t)chunk * P.c + ch) = acc;
}

// VEC consecutive channels per thread (16-byte loads / stores when c % 4 == 0)
template <int VEC>
__global__ void __launch_bounds__(256)
pool_combine_kernel(const float *__restrict__ partial, const int32_t *__restrict__ cnt_in,
                    const int64_t *__restrict__ pair_off, const int32_t *__restrict__ chunk_off, int n_units,
                    int64_t cap_pairs, int k, int c, float *__restrict__ sum, int32_t *__restrict__ cnt,
                    float *__restrict__ mean, int direct) {
    using V = typename VecT<VEC>::type;
    const int s = blockIdx.y;
    const int cv = c / VEC;                                    // vectors per row
    const int e = blockIdx.x * blockDim.x + threadIdx.x;       // vector element of [k, c / VEC]
    if (e >= k * cv) return;
    const int m = e / cv, ch = (e - m * cv) * VEC;
    const int u = s * k + m;
    const bool ok = pair_off[n_units] <= cap_pairs;
    if (ok && direct && chunk_off[u + 1] - chunk_off[u] == 1) return;      // finished by the sum kernel
    const int n = ok ? cnt_in[u] : 0;
    V acc; vzero(acc);
    if (ok)
        for (int q = chunk_off[u]; q < chunk_off[u + 1]; ++q)      // fixed order
            vadd(acc, *reinterpret_cast<const V *>(partial + (size_t)q * c + ch));
    const size_t o = ((size_t)s * k + m) * c + ch;
    *reinterpret_cast<V *>(sum + o) = acc;
    if (mean) {
        V mv = acc;
        float *a = reinterpret_cast<float *>(&mv);
#pragma unroll
        for (int j = 0; j < VEC; ++j) a[j] = n > 0 ? __fdiv_rn(a[j], (float)n) : 0.f;
        *reinterpret_cast<V *>(mean + o) = mv;
    }
    if (cnt && ch == 0) cnt[u] = n;
}


// ---- point-major variant for overlapping masks ------------------------------------------------
// With overlapping masks the pair lists above read every feature row once PER MEMBERSHIP (7.7 x 3 KB per
// point in the reference's raw thresholded predictions).  Here every row is read from HBM exactly once:
// a work item is (segment, 128-channel slice); a persistent CTA streams the item's rows through two
// 80 KB shared-memory buffers (one 512-byte bulk copy per row, mbarrier-tracked, the next tile in
// flight while the current one is consumed — row_index indirection included), transposes the tile's
// membership words into per-mask bit rows with warp ballots, and every warp sums the member rows of
// its masks out of shared memory in point order (one float4 per lane), adding the tile's partial to the
// mask's accumulator row in shared memory.  No pair list, no partial rows in HBM, no combine pass;
// the order of every sum is fixed (tiles ascending, points ascending), so the result is deterministic.
bool make_row_tile_map(CUtensorMap *m, const float *base, int64_t rows, int c, int box_cols, int box_rows);   // logits.cu
// tensor-core variant (pool_mma.cu)
bool pool_mma_eligible(const float *feat, int c, const int32_t *row_index, const uint32_t *member, int k, int64_t cap,
                       const float *sum, const float *mean);
int launch_pool_mma(const float *feat, int c, const uint32_t *member, int words, int n_seg, int k, const int64_t *seg_off,
                    int64_t cap, float *sum, int32_t *cnt, float *mean, int *work, int *order, int tune, int32_t *status,
                    cudaStream_t stream);

constexpr int PR_THREADS = 512;
constexpr int PR_WARPS = PR_THREADS / 32;
constexpr int PR_SLICE = 128;                 // channels per work item
constexpr int PR_TP = 160;                    // points per tile
constexpr int PR_GROUPS = PR_TP / 32;
constexpr int PR_KMAX = 96;                   // accumulator rows that fit next to the two buffers
static size_t pool_rows_smem(int k) {
    return 2 * (size_t)PR_TP * PR_SLICE * 4 + 2 * (size_t)PR_TP * 4 * 4 + (size_t)k * PR_SLICE * 4 + (size_t)((k + 3) & ~3) * 4 +
           (size_t)PR_WARPS * PR_TP + 128;
}

template <int W>
__global__ void __launch_bounds__(PR_THREADS, 1)
pool_rows_kernel(const __grid_constant__ CUtensorMap map, const int use_map, const PoolIdx P,
                 const float *__restrict__ feat, int c, float *__restrict__ sum,
                 float *__restrict__ mean, int32_t *__restrict__ cnt, int *__restrict__ work) {
    extern __shared__ __align__(128) unsigned char pr_smem[];
    float *buf0 = reinterpret_cast<float *>(pr_smem);
    float *buf1 = buf0 + PR_TP * PR_SLICE;
    uint32_t *bits_all = reinterpret_cast<uint32_t *>(buf1 + PR_TP * PR_SLICE);      // [2][PR_TP][W] membership words of two tiles
    float *acc = reinterpret_cast<float *>(bits_all + 2 * PR_TP * W);                // [k][PR_SLICE], 16-byte aligned
    int *cntm = reinterpret_cast<int *>(acc + P.k * PR_SLICE);                       // [k]
    unsigned char *lst = reinterpret_cast<unsigned char *>(cntm + ((P.k + 3) & ~3)) + (threadIdx.x >> 5) * PR_TP;   // per warp
    __shared__ uint64_t s_full[2];
    __shared__ int s_item, s_next;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int nsl = c / PR_SLICE;
    const int n_items = P.n_seg * nsl;
    if (tid == 0) { mbar_init(&s_full[0], 1); mbar_init(&s_full[1], 1); mbar_fence_init(); }
    __syncthreads();
    uint32_t ph0 = 0, ph1 = 0;
    const int64_t total_rows = P.seg_off[P.n_seg];
    const bool over = total_rows > P.cap;
    for (;;) {
        if (tid == 0) s_item = atomicAdd(work, 1);
        __syncthreads();
        const int item = s_item;
        if (item >= n_items) break;
        const int s = item / nsl, sl = item - s * nsl;
        const int64_t a = P.seg_off[s];
        const int n = over ? 0 : (int)(P.seg_off[s + 1] - a);
        for (int j = tid; j < P.k * PR_SLICE; j += PR_THREADS) acc[j] = 0.f;
        for (int j = tid; j < P.k; j += PR_THREADS) cntm[j] = 0;
        const int ntile = (n + PR_TP - 1) / PR_TP;
        const float *fsl = feat + (size_t)sl * PR_SLICE;
        auto issue = [&](int t) {                         // warp 0: the rows of tile t -> buffer t & 1
            const int rows = min(PR_TP, n - t * PR_TP);
            uint64_t *bar = &s_full[t & 1];
            float *dst = (t & 1) ? buf1 : buf0;
            if (use_map && a + (int64_t)(t + 1) * PR_TP <= total_rows) {
                // contiguous rows: ONE tensor copy per tile (160 bulk copies of 512 bytes cost ~10 us per tile);
                // rows past the segment (they belong to the next one) are loaded but never referenced; a tile that
                // would reach past the last point of the batch takes the row-by-row path below
                if (lane == 0) {
                    mbar_expect_tx(bar, (uint32_t)PR_TP * PR_SLICE * 4);
                    asm volatile(
                        "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
                        ::"r"(smem_u32(dst)), "l"(&map), "r"(smem_u32(bar)), "r"(sl * PR_SLICE), "r"((int)(a + (int64_t)t * PR_TP))
                        : "memory");
                }
                return;
            }
            if (lane == 0) mbar_expect_tx(bar, (uint32_t)rows * PR_SLICE * 4);
            __syncwarp();
            for (int r = lane; r < rows; r += 32) {
                const int64_t i = a + (int64_t)t * PR_TP + r;
                const int64_t row = P.row_index ? (int64_t)__ldg(P.row_index + i) : i;
                bulk_g2s(dst + r * PR_SLICE, fsl + (size_t)row * c, PR_SLICE * 4, bar);
            }
        };
        if (warp == 0 && ntile > 0) issue(0);
        if (warp < PR_GROUPS && ntile > 0) {              // membership words of tile 0
            const int pnt = warp * 32 + lane;
            uint32_t b0[W];
            load_bits<W>(P, a + pnt, pnt < n, b0);
#pragma unroll
            for (int w = 0; w < W; ++w) bits_all[(warp * 32 + lane) * W + w] = b0[w];
        }
        for (int t = 0; t < ntile; ++t) {
            // the other buffer is free: everybody passed the barrier that ends iteration t - 1
            if (warp == 0 && t + 1 < ntile) issue(t + 1);
            if (tid == PR_THREADS - 1) s_next = 0;
            __syncthreads();                              // the words of tile t are in wbuf (stored one iteration ago)
            const uint32_t *bits = bits_all + (t & 1) * (PR_TP * W);
            uint32_t nb[W];
            const bool pre = warp < PR_GROUPS && t + 1 < ntile;
            if (pre) {                                    // membership words of the NEXT tile: in flight during the sums
                const int pnt = (t + 1) * PR_TP + warp * 32 + lane;
                load_bits<W>(P, a + pnt, pnt < n, nb);
            }
            if (t & 1) { mbar_wait(&s_full[1], ph1); ph1 ^= 1; } else { mbar_wait(&s_full[0], ph0); ph0 ^= 1; }
            const float4 *rows4 = reinterpret_cast<const float4 *>((t & 1) ? buf1 : buf0);
            for (;;) {
                // masks are handed out dynamically: their sizes differ by an order of magnitude (every mask is
                // still summed by ONE warp per tile, in point order, so the result does not depend on who takes it)
                int m = 0;
                if (lane == 0) m = atomicAdd(&s_next, 1);
                m = __shfl_sync(0xffffffffu, m, 0);
                if (m >= P.k) break;
                // bit rows of mask m (one ballot per group of 32 points), kept in registers
                uint32_t v[PR_GROUPS];
                int members = 0;
#pragma unroll
                for (int g = 0; g < PR_GROUPS; ++g) {
                    v[g] = __ballot_sync(0xffffffffu, (bits[(g * 32 + lane) * W + (m >> 5)] >> (m & 31)) & 1u);
                    members += __popc(v[g]);
                }
                if (!members) continue;
                // the members' row numbers as a dense per-warp list (ballot ranks), so that the sum is a loop with a
                // uniform trip count: 8 instead of 22 instructions per row (ncu: the kernel is issue bound)
                {
                    int basec = 0;
#pragma unroll
                    for (int g = 0; g < PR_GROUPS; ++g) {
                        if ((v[g] >> lane) & 1u) lst[basec + __popc(v[g] & ((1u << lane) - 1u))] = (unsigned char)(g * 32 + lane);
                        basec += __popc(v[g]);
                    }
                }
                __syncwarp();
                float4 part = make_float4(0.f, 0.f, 0.f, 0.f);
                const float4 *rl = rows4 + lane;
                int j = 0;
                for (; j + 4 <= members; j += 4) {            // four rows in flight, added in point order
                    const uchar4 id = *reinterpret_cast<const uchar4 *>(lst + j);
                    const float4 x0 = rl[id.x * (PR_SLICE / 4)], x1 = rl[id.y * (PR_SLICE / 4)];
                    const float4 x2 = rl[id.z * (PR_SLICE / 4)], x3 = rl[id.w * (PR_SLICE / 4)];
                    vadd(part, x0); vadd(part, x1); vadd(part, x2); vadd(part, x3);
                }
                for (; j < members; ++j) vadd(part, rl[lst[j] * (PR_SLICE / 4)]);
                __syncwarp();                                 // the list is rewritten for the warp's next mask
                float4 *A = reinterpret_cast<float4 *>(acc + m * PR_SLICE) + lane;
                float4 cur = *A;
                vadd(cur, part);
                *A = cur;
                if (lane == 0) cntm[m] += members;
            }
            if (pre) {
#pragma unroll
                for (int w = 0; w < W; ++w) bits_all[((t + 1) & 1) * (PR_TP * W) + (warp * 32 + lane) * W + w] = nb[w];
            }
            __syncthreads();
        }
        if (ntile == 0) __syncthreads();                  // the zeroing above before the read below
        for (int m = warp; m < P.k; m += PR_WARPS) {
            const float4 v = *(reinterpret_cast<const float4 *>(acc + m * PR_SLICE) + lane);
            const int nm = cntm[m];
            const size_t o = ((size_t)s * P.k + m) * c + (size_t)sl * PR_SLICE + lane * 4;
            *reinterpret_cast<float4 *>(sum + o) = v;
            if (mean) {
                float4 q = make_float4(0.f, 0.f, 0.f, 0.f);
                if (nm > 0) {
                    const float d = (float)nm;
                    q = make_float4(__fdiv_rn(v.x, d), __fdiv_rn(v.y, d), __fdiv_rn(v.z, d), __fdiv_rn(v.w, d));
                }
                *reinterpret_cast<float4 *>(mean + o) = q;
            }
            if (cnt && sl == 0 && lane == 0) cnt[s * P.k + m] = nm;
        }
        __syncthreads();
    }
}

struct PoolWs {
    int32_t *cnt, *chunk_off, *perm, *tile_off, *tile_cnt;
    int64_t max_tiles;
    int64_t *pair_off;
    float *partial;
    int64_t max_chunks;
    int *work;                 // work-item counter of the point-major / tensor-core kernels
    int *order;                // [n_seg] segments by descending size (longest-processing-time hand-out)
};

static PoolWs carve_pool(void *ws, int n_seg, int k, int c, int64_t cap, int64_t cap_pairs, size_t *bytes) {
    Carver cv(ws);
    PoolWs w;
    const size_t units = (size_t)n_seg * k;
    w.max_chunks = cap_pairs / POOL_CH + (int64_t)units + 1;
    w.cnt = cv.take<int32_t>(units);
    w.chunk_off = cv.take<int32_t>(units + 1);
    w.pair_off = cv.take<int64_t>(units + 1);
    w.perm = cv.take<int32_t>((size_t)cap_pairs + 32);
    w.partial = cv.take<float>((size_t)w.max_chunks * c);
    w.max_tiles = cap / FILL_THREADS + n_seg + 1;
    w.tile_off = cv.take<int32_t>((size_t)n_seg + 1);
    w.tile_cnt = cv.take<int32_t>((size_t)w.max_tiles * k);
    w.work = cv.take<int>(64);
    w.order = cv.take<int>((size_t)n_seg);
    *bytes = cv.off + 256;
    return w;
}

}  // namespace xm3d

using namespace xm3d;

extern "C" size_t xm3d_pool_ws_bytes(int32_t n_seg, int32_t k, int32_t c, int64_t cap, int64_t cap_pairs) {
    size_t b = 0;
    carve_pool(nullptr, n_seg, k, c, cap, cap_pairs, &b);
    return b;
}

extern "C" int xm3d_pool_batch(const float *feat, int32_t c, const int32_t *row_index, const uint32_t *member,
                               const int32_t *label, int32_t n_seg, int32_t k, const int64_t *seg_off, int64_t cap,
                               int64_t cap_pairs, int32_t path, float *sum, int32_t *cnt, float *mean, void *ws,
                               size_t ws_bytes, int32_t *status, xm3d_stream_t stream_) {
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    XM3D_REQUIRE(n_seg > 0 && k > 0 && c > 0 && cap >= 0 && cap_pairs >= 0, "bad sizes");
    XM3D_REQUIRE(k <= 32 * MAX_WORDS, "at most 256 masks per segment");
    XM3D_REQUIRE(feat && seg_off && sum && ws, "null pointer");
    XM3D_REQUIRE((member != nullptr) != (label != nullptr), "exactly one of member / label");
    XM3D_REQUIRE(cap_pairs < ((int64_t)1 << 31) && (int64_t)n_seg * k < ((int64_t)1 << 30), "sizes exceed int32");
    size_t need = 0;
    PoolWs w = carve_pool(ws, n_seg, k, c, cap, cap_pairs, &need);
    if (ws_bytes < need) {
        set_error("xm3d_pool_batch: workspace too small (%zu < %zu)", ws_bytes, need);
        return XM3D_ERR_WORKSPACE;
    }
    const int vec = (c % 4 == 0 && c / 4 <= 1024 && reinterpret_cast<uintptr_t>(feat) % 16 == 0) ? 4 : 1;
    XM3D_REQUIRE(vec == 4 || c <= 1024, "feature width not supported (c % 4 != 0 and c > 1024)");
    PoolIdx I;
    I.member = member; I.label = label; I.row_index = row_index; I.seg_off = seg_off; I.n_seg = n_seg; I.k = k;
    I.words = words_for(k); I.cap = cap;
    const int n_units = n_seg * k;
    // Overlapping masks (the caller's bound allows more memberships than points): every row is read once, by the
    // tensor-core kernel (pool_mma.cu) or the point-major CUDA-core kernel.  cap_pairs is not a limit on these
    // paths (there is no pair list to overrun).
    const int tune = path >> 8;          // experiments: ring depths of the tensor-core kernel (0 = defaults)
    path &= 0xff;
    XM3D_REQUIRE(path >= XM3D_POOL_AUTO && path <= XM3D_POOL_MMA, "unknown pooling path");
    const bool out16 = reinterpret_cast<uintptr_t>(sum) % 16 == 0 && (!mean || reinterpret_cast<uintptr_t>(mean) % 16 == 0);
    const bool overlap = member && cap_pairs > cap + 1;
    const bool mma_ok = pool_mma_eligible(feat, c, row_index, member, k, cap, sum, mean);
    const bool rows_ok = member && vec == 4 && c % PR_SLICE == 0 && k <= PR_KMAX && out16;
    if (path == XM3D_POOL_MMA && !mma_ok) { set_error("xm3d_pool_batch: tensor-core path not eligible"); return XM3D_ERR_UNSUPPORTED; }
    if (path == XM3D_POOL_ROWS && !rows_ok) { set_error("xm3d_pool_batch: point-major path not eligible"); return XM3D_ERR_UNSUPPORTED; }
    if (path == XM3D_POOL_MMA || (path == XM3D_POOL_AUTO && overlap && mma_ok)) {
        if (g_pool_ev[0]) cudaEventRecord(g_pool_ev[0], stream);
        const int rc = launch_pool_mma(feat, c, member, I.words, n_seg, k, seg_off, cap, sum, cnt, mean, w.work, w.order, tune, status, stream);
        if (g_pool_ev[1]) cudaEventRecord(g_pool_ev[1], stream);
        return rc;
    }
    if (path == XM3D_POOL_ROWS || (path == XM3D_POOL_AUTO && overlap && rows_ok)) {
        const size_t smem = pool_rows_smem(k);
        static std::atomic<uint64_t> attr_set{0};
        if (first_use_on_device(&attr_set)) {
            cudaFuncSetAttribute(pool_rows_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 224 * 1024);
            cudaFuncSetAttribute(pool_rows_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, 224 * 1024);
            cudaFuncSetAttribute(pool_rows_kernel<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, 224 * 1024);
        }
        cudaMemsetAsync(w.work, 0, sizeof(int), stream);
        const int n_items = n_seg * (c / PR_SLICE);
        const unsigned grid = (unsigned)(n_items < sm_count() ? n_items : sm_count());
        CUtensorMap map;
        memset(&map, 0, sizeof(map));
        // (the map's row count only bounds the coordinates: the kernel never asks for rows past the last point)
        const int use_map = (!row_index && cap > 0 && cap < ((int64_t)1 << 31) - PR_TP &&
                             make_row_tile_map(&map, feat, cap, c, PR_SLICE, PR_TP)) ? 1 : 0;
        if (g_pool_ev[0]) cudaEventRecord(g_pool_ev[0], stream);
        if (I.words <= 1) pool_rows_kernel<1><<<grid, PR_THREADS, smem, stream>>>(map, use_map, I, feat, c, sum, mean, cnt, w.work);
        else if (I.words <= 2) pool_rows_kernel<2><<<grid, PR_THREADS, smem, stream>>>(map, use_map, I, feat, c, sum, mean, cnt, w.work);
        else pool_rows_kernel<4><<<grid, PR_THREADS, smem, stream>>>(map, use_map, I, feat, c, sum, mean, cnt, w.work);
        if (g_pool_ev[1]) cudaEventRecord(g_pool_ev[1], stream);
        count_launches(1);
        return check_launch("xm3d_pool_batch");
    }
    pool_tileplan_kernel<<<1, 1024, 0, stream>>>(seg_off, n_seg, cap, w.tile_off); count_launches(1);
    const int wt = I.words <= 1 ? 1 : (I.words <= 2 ? 2 : (I.words <= 4 ? 4 : 8));
    const unsigned tgrid = (unsigned)w.max_tiles;
    switch (wt) {
        case 1: pool_tilecount_kernel<1><<<tgrid, FILL_THREADS, 0, stream>>>(I, w.tile_off, w.tile_cnt); break;
        case 2: pool_tilecount_kernel<2><<<tgrid, FILL_THREADS, 0, stream>>>(I, w.tile_off, w.tile_cnt); break;
        case 4: pool_tilecount_kernel<4><<<tgrid, FILL_THREADS, 0, stream>>>(I, w.tile_off, w.tile_cnt); break;
        default: pool_tilecount_kernel<8><<<tgrid, FILL_THREADS, 0, stream>>>(I, w.tile_off, w.tile_cnt); break;
    }
    count_launches(1);
    pool_unitprefix_kernel<<<n_seg, 256, 0, stream>>>(w.tile_off, k, w.tile_cnt, w.cnt); count_launches(1);
    pool_scan_kernel<<<1, 1024, 0, stream>>>(w.cnt, n_units, cap_pairs, w.pair_off, w.chunk_off, status); count_launches(1);
    switch (wt) {
        case 1: pool_fill_kernel<1><<<tgrid, FILL_THREADS, 0, stream>>>(I, w.tile_off, w.tile_cnt, w.pair_off, cap_pairs, w.perm); break;
        case 2: pool_fill_kernel<2><<<tgrid, FILL_THREADS, 0, stream>>>(I, w.tile_off, w.tile_cnt, w.pair_off, cap_pairs, w.perm); break;
        case 4: pool_fill_kernel<4><<<tgrid, FILL_THREADS, 0, stream>>>(I, w.tile_off, w.tile_cnt, w.pair_off, cap_pairs, w.perm); break;
        default: pool_fill_kernel<8><<<tgrid, FILL_THREADS, 0, stream>>>(I, w.tile_off, w.tile_cnt, w.pair_off, cap_pairs, w.perm); break;
    }
    count_launches(1);
    SumParams S;
    S.feat = feat; S.c = c; S.perm = w.perm; S.pair_off = w.pair_off; S.chunk_off = w.chunk_off; S.n_units = n_units;
    S.cap_pairs = cap_pairs; S.partial = w.partial;
    const bool cv4 = vec == 4 && reinterpret_cast<uintptr_t>(sum) % 16 == 0 && (!mean || reinterpret_cast<uintptr_t>(mean) % 16 == 0);
    S.cnt_in = w.cnt; S.sum = sum; S.mean = mean; S.cnt = cnt; S.direct = (cv4 || vec == 1) ? 1 : 0;
    const int threads = ((c + vec - 1) / vec + 31) / 32 * 32;
    if (g_pool_ev[0]) cudaEventRecord(g_pool_ev[0], stream);
    if (vec == 4) { pool_sum_kernel<4><<<(unsigned)w.max_chunks, threads, 0, stream>>>(S); count_launches(1); }
    else { pool_sum_kernel<1><<<(unsigned)w.max_chunks, threads, 0, stream>>>(S); count_launches(1); }
    if (g_pool_ev[1]) cudaEventRecord(g_pool_ev[1], stream);
    if (cv4) {
        dim3 cgrid((unsigned)(((size_t)k * (c / 4) + 255) / 256), n_seg);
        pool_combine_kernel<4><<<cgrid, 256, 0, stream>>>(w.partial, w.cnt, w.pair_off, w.chunk_off, n_units, cap_pairs, k, c,
                                                          sum, cnt, mean, S.direct);
    } else {
        dim3 cgrid((unsigned)(((size_t)k * c + 255) / 256), n_seg);
        pool_combine_kernel<1><<<cgrid, 256, 0, stream>>>(w.partial, w.cnt, w.pair_off, w.chunk_off, n_units, cap_pairs, k, c,
                                                          sum, cnt, mean, S.direct);
    }
    count_launches(1);
    return check_launch("xm3d_pool_batch");
}
