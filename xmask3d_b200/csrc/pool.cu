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

struct PoolWs {
    int32_t *cnt, *chunk_off, *perm, *tile_off, *tile_cnt;
    int64_t max_tiles;
    int64_t *pair_off;
    float *partial;
    int64_t max_chunks;
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
                               int64_t cap_pairs, float *sum, int32_t *cnt, float *mean, void *ws, size_t ws_bytes,
                               int32_t *status, xm3d_stream_t stream_) {
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
