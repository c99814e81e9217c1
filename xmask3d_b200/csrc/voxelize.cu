// Stage 1 — voxelization: grid quantisation, FNV-1 keys, unique / first-index / inverse maps.
//
// Replaces Voxelizer.voxelize after the matrix draw (reference dataset/voxelizer.py:110-122),
// fnv_hash_vec / ravel_hash_vec (dataset/voxelization_utils.py:6-35) and the np.unique call of
// sparse_quantize (:86, :95), batched over segments (one segment = one (scene, view)).
// np.unique orders its output by ascending KEY (the 64-bit FNV value), returns the index of the
// first occurrence and the rank of every element; all three are reproduced bit-exactly.
//
//   plan      per-segment table / splitter offsets, counters zeroed                      (1 CTA)
//   clear     hash-table slots <- empty
//   min       grid = floor(FMA-chain transform); per-segment column minima
//   sample    one CTA per segment: S sample keys, bitonic-sorted in shared memory -> splitters
//   insert    key = FNV-1(grid - min).  Warp-cooperative open addressing: lanes holding the same
//             key elect the lowest lane (= lowest point index) with match.any; only that lane
//             probes the segment's table (linear probing, 16-byte slots {key, -, rank}).  The
//             winner of an empty slot appends (slot, key) to the segment's unique list (one
//             warp-aggregated atomicAdd), finds the key's splitter bucket and takes a ticket in it
//   bscan     one CTA per segment: exclusive scan of the bucket counts; prefix of unique counts
//   scatter   unique keys -> bucket order
//   rank      one thread per unique key: rank = bucket start + #smaller keys inside its bucket (~10
//             keys, staged in shared memory); the rank is written into the key's table slot
//   inverse   inverse[i] = rank stored in the slot point i resolved to; first[rank] = atomicMin(i)
//   coords    voxel coordinates of every unique key from its first occurrence, in unique order
//
// Fast path (vox_fast_kernel, taken whenever every segment fits): after `min` has written the grid
// coordinates of every point, ONE persistent 1024-thread CTA per unit does insert / rank / inverse /
// coords entirely in shared memory (16384-slot table of 64-bit keys, unique list from a scan of the
// table, bucket ranking as above) — no dependent global-memory round trips, so it does not starve next
// to the pooling stream.  A segment of n points is split by KEY RANGE into ceil(n / unit_pts) units
// (quantiles of a 256 / 1024-key sample every unit of the segment computes identically); the owner
// of a point tags its stored grid record in place.  Unique counts of the units are exchanged through
// a decoupled look-back, which yields both the rank base inside the segment and uniq_off.  A unit
// that overflows its table raises ctl[1] and the multi-kernel path below (gated on ctl) recomputes
// the whole batch.
//
// Everything that decides an integer is exact: the transform is the fp64 FMA chain numpy's dgemm
// performs, floor() is exact, keys are 64-bit so FNV collisions merge voxels exactly as the
// reference's np.unique does.  The sample sort is comparison based, so the skewed high bits of
// FNV-1 over small integers (only ~50 distinct top-12-bit patterns per scene) cost nothing.
#include "common.cuh"

namespace xm3d {

constexpr unsigned long long KEY_EMPTY = 0xFFFFFFFFFFFFFFFFull;
constexpr unsigned long long FNV_OFFSET = 14695981039346656037ull;
constexpr unsigned long long FNV_PRIME = 1099511628211ull;
constexpr int VOX_THREADS = 256;
constexpr int SPL_MAX = 4096;             // splitters (= buckets) per segment, at most
constexpr int SPL_MIN = 32;
constexpr int GRID_LIMIT = 1 << 30;
// shared-memory fast path
#ifndef XM3D_FV_THREADS
#define XM3D_FV_THREADS 1024
#endif
// (measured: 512- / 256-thread units, which leave room for the pooling stream's CTAs on the same SM, make
// the step slower — 1.82 / 1.95 ms against 1.80 ms — the SM time of the two streams simply adds up)
constexpr int FV_THREADS = XM3D_FV_THREADS;
constexpr int FV_TABLE = 16384;           // 64-bit key slots per unit (128 KB)
constexpr int FV_MU = 8704;               // unique keys per unit, at most (load factor <= 0.53)
constexpr int FV_UNIT_PTS = 7000;         // points per unit the plan aims at (24 % headroom to FV_MU)
constexpr int FV_UNIT_PTS_MIN = 64;       // smallest unit size a call may ask for (sizes the unit tables)
constexpr int FV_PMAX = 32;               // units per segment, at most (224 k points)
constexpr int FV_OUT_BATCH = 6;             // gathers of voxel coordinates in flight per thread (one round for M <= 6144)
constexpr int FV_PROBE_MAX = 2048;        // a longer probe sequence means the table is (nearly) full: give up
constexpr int FV_BATCH = 4;               // loads a thread keeps in flight in the point passes
constexpr int FV_NS = 1024;               // sample keys for the key-range split / rank buckets
constexpr size_t FV_SMEM = (size_t)FV_TABLE * 8 + (size_t)FV_MU * 4 + (size_t)FV_MU * 2 + (size_t)FV_MU * 4 +
                           (size_t)FV_NS * 8 + (size_t)(FV_NS + 32) * 4;


struct __align__(16) Slot {
    unsigned long long key;
    unsigned int pad;
    unsigned int rank;    // rank of the key among the segment's unique keys
};

// slots per segment table: load factor <= 0.67, an even number (slots are probed in aligned pairs)
__host__ __device__ inline int64_t table_size(int64_t n) { return (n + n / 2 + 32) & ~(int64_t)1; }
// buckets of a segment of n elements: power of two, about one per 8..16 elements
__host__ __device__ inline int bucket_count(int64_t n) {
    int s = SPL_MIN;
    while (s < SPL_MAX && (int64_t)s * 16 < n) s <<= 1;
    return s;
}

__device__ __forceinline__ unsigned long long mix64(unsigned long long h) {
    h ^= h >> 33; h *= 0xff51afd7ed558ccdull;
    h ^= h >> 33; h *= 0xc4ceb9fe1a85ec53ull;
    h ^= h >> 33;
    return h;
}

__device__ __forceinline__ unsigned long long fnv3(unsigned long long a, unsigned long long b,
                                                   unsigned long long c) {
    unsigned long long h = FNV_OFFSET;      // voxelization_utils.py:13-17: multiply, then xor the word
    h *= FNV_PRIME; h ^= a;
    h *= FNV_PRIME; h ^= b;
    h *= FNV_PRIME; h ^= c;
    return h;
}

// numpy's float64 -> uint64 cast on x86-64 (cvttsd2si below 2^63, wraps negatives)
__device__ __forceinline__ unsigned long long f64_to_u64_numpy(double d) {
    if (d < 9223372036854775808.0) return (unsigned long long)(long long)d;
    return (unsigned long long)d;
}

// Point coordinates as the caller holds them: float32 (ScanNet .pth) or float64 (what the reference's
// ElasticDistortion returns, dataset/augmentation.py:171 — `homo_coords` then stays float64, voxelizer.py:110-112).
struct Xyz {
    const float *f;
    const double *d;
    __host__ __device__ explicit operator bool() const { return f != nullptr || d != nullptr; }
};

// floor([x y z 1] @ RT.T[:, :3]) for one point: voxelizer.py:110-113
__device__ __forceinline__ void grid_of(const Xyz src, int64_t i, const double *__restrict__ rt, double out[3]) {
    double x, y, z;
    if (src.d) { x = src.d[i * 3]; y = src.d[i * 3 + 1]; z = src.d[i * 3 + 2]; }
    else { x = (double)src.f[i * 3]; y = (double)src.f[i * 3 + 1]; z = (double)src.f[i * 3 + 2]; }
#pragma unroll
    for (int j = 0; j < 3; ++j) {
        const double *r = rt + 4 * j;
        double s = __dmul_rn(x, r[0]);
        s = __fma_rn(y, r[1], s);
        s = __fma_rn(z, r[2], s);
        s = __fma_rn(1.0, r[3], s);
        out[j] = floor(s);
    }
}

// FNV key of point i of segment s (voxelizer.py:115-121: floor(grid - min) -> uint64 -> FNV-1)
__device__ __forceinline__ unsigned long long point_key(const Xyz xyz, int64_t i, int s,
                                                        const double *__restrict__ rt,
                                                        const int *__restrict__ grid_min) {
    double f[3];
    grid_of(xyz, i, rt + 12 * s, f);
    unsigned long long w[3];
#pragma unroll
    for (int j = 0; j < 3; ++j) {
        if (!(f[j] > -(double)GRID_LIMIT && f[j] < (double)GRID_LIMIT)) f[j] = 0.0;
        // floor(grid - min) of two integers is their exact difference
        w[j] = (unsigned long long)(long long)((int)f[j] - grid_min[3 * s + j]);
    }
    return fnv3(w[0], w[1], w[2]);
}

__device__ __forceinline__ unsigned long long clean_key(unsigned long long key, int *status) {
    if (key == KEY_EMPTY) {              // 2^-64 event: the sentinel value cannot be stored
        if (status) atomicOr(status, XM3D_FLAG_KEY_SENTINEL);
        return KEY_EMPTY - 1;
    }
    return key;
}

// bucket of a key: number of splitters < key, splitter j (0..S-2) = spl[j+1]
__device__ __forceinline__ int bucket_of(const unsigned long long *__restrict__ spl, int S, unsigned long long key) {
    int lo = 0, hi = S - 1;
    while (lo < hi) {
        const int mid = (lo + hi) >> 1;
        if (spl[mid + 1] < key) lo = mid + 1; else hi = mid;
    }
    return lo;
}

// exclusive prefix of `val` over a 1024-thread block, carried across calls through *s_carry
__device__ __forceinline__ int64_t block_excl_scan_1024(int64_t val, int64_t *s_warp, int64_t *s_carry) {
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
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
    const int64_t excl = *s_carry + s_warp[warp] + incl - val;
    __syncthreads();
    if (tid == 1023) *s_carry = excl + val;
    __syncthreads();
    return excl;
}

// Segment of every thread of a block that walks consecutive elements: almost every block lies inside
// one segment, so thread 0 resolves the block's first and last element and the others reuse it.
struct BlockSeg { int s; bool uniform; };
__device__ __forceinline__ BlockSeg block_segment(const int64_t *__restrict__ seg_off, int n_seg, int64_t total,
                                                  int64_t i, bool valid, int *s_pair /*[2] shared*/, int items = 1) {
    if (threadIdx.x < 32) {
        const int64_t span = (int64_t)blockDim.x * items;
        const int64_t b0 = (int64_t)blockIdx.x * span;
        const int64_t last = (b0 + span - 1 < total) ? b0 + span - 1 : total - 1;
        if (n_seg <= 2048) {
            // seg_of(x) = #{ j in [1, n_seg) : seg_off[j] <= x }: the warp counts both targets with all loads in flight at
            // once (one memory round trip) instead of two binary searches of dependent loads by one thread (~3 us per
            // block, which was most of vox_min's 44 us)
            int c0 = 0, c1 = 0;
            for (int j = 1 + (int)threadIdx.x; j < n_seg; j += 32) {
                const int64_t v = seg_off[j];
                c0 += v <= b0;
                c1 += v <= last;
            }
            c0 = __reduce_add_sync(0xffffffffu, c0);
            c1 = __reduce_add_sync(0xffffffffu, c1);
            if (threadIdx.x == 0) { s_pair[0] = c0; s_pair[1] = c1; }
        } else if (threadIdx.x == 0) {
            s_pair[0] = seg_of(seg_off, n_seg, b0);
            s_pair[1] = seg_of(seg_off, n_seg, last);
        }
    }
    __syncthreads();
    BlockSeg r;
    r.uniform = s_pair[0] == s_pair[1];
    r.s = r.uniform ? s_pair[0] : (valid ? seg_of(seg_off, n_seg, i) : s_pair[0]);
    return r;
}

// ---- plan ---------------------------------------------------------------------------------
// ctl[0] = 1: the fast path is not eligible (decided here); ctl[1] = 1: a fast unit overflowed.
__global__ void __launch_bounds__(1024, 1)
vox_plan_kernel(const int64_t *__restrict__ seg_off, int n_seg, int64_t cap, int64_t *__restrict__ tbl_off,
                int64_t *__restrict__ spl_off, int64_t *__restrict__ total_eff, int *__restrict__ m,
                int *__restrict__ grid_min, int *status, int unit_pts, int units_max, int force_slow,
                int *__restrict__ unit_off, int *__restrict__ unit_seg, int *__restrict__ unit_m,
                int *__restrict__ ctl) {
    __shared__ int64_t s_warp[32];
    __shared__ int64_t s_carry[3];
    __shared__ int s_slow;
    const int tid = threadIdx.x;
    // more elements than the caller's capacity: flag it and process nothing (never overrun)
    const bool over = seg_off[n_seg] > cap;
    if (tid == 0) {
        s_carry[0] = s_carry[1] = s_carry[2] = 0;
        s_slow = force_slow & 1;
        *total_eff = over ? 0 : seg_off[n_seg];
        if (over && status) atomicOr(status, XM3D_FLAG_VIS_OVERFLOW);
    }
    for (int t = tid; t < units_max; t += 1024) unit_m[t] = -1;
    __syncthreads();
    for (int base = 0; base < n_seg; base += 1024) {
        const int s = base + tid;
        int64_t vt = 0, vs = 0, vu = 0;
        if (s < n_seg) {
            const int64_t n = over ? 0 : seg_off[s + 1] - seg_off[s];
            vt = table_size(n);
            vs = bucket_count(n);
            vu = n <= unit_pts ? 1 : (n + unit_pts - 1) / unit_pts;
            if (vu > FV_PMAX) { s_slow = 1; vu = 1; }
            m[s] = 0;
            if (grid_min) grid_min[3 * s] = grid_min[3 * s + 1] = grid_min[3 * s + 2] = 0x7fffffff;
        }
        const int64_t et = block_excl_scan_1024(vt, s_warp, &s_carry[0]);
        const int64_t es = block_excl_scan_1024(vs, s_warp, &s_carry[1]);
        const int64_t eu = block_excl_scan_1024(vu, s_warp, &s_carry[2]);
        if (s < n_seg) {
            tbl_off[s] = et; spl_off[s] = es; unit_off[s] = (int)eu;
            if (eu + vu <= units_max)
                for (int p = 0; p < (int)vu; ++p) unit_seg[eu + p] = s;
        }
    }
    __syncthreads();
    if (tid == 0) {
        tbl_off[n_seg] = s_carry[0]; spl_off[n_seg] = s_carry[1]; unit_off[n_seg] = (int)s_carry[2];
        ctl[0] = (s_slow || s_carry[2] > units_max) ? 1 : 0;
        ctl[1] = 0;
        ctl[2] = 0;                               // unit ticket counter of vox_fast_kernel
        ctl[3] = (force_slow >> 1) & 1;           // XM3D_VOX_FAST_ONLY: the multi-kernel path is not launched
        if (ctl[3] && ctl[0] && status) atomicOr(status, XM3D_FLAG_VOX_FALLBACK);   // ... so "not eligible" is an error
    }
}

__global__ void __launch_bounds__(256)
vox_clear_kernel(Slot *__restrict__ tbl, const int64_t *__restrict__ tbl_off, int n_seg,
                 const int64_t *__restrict__ total_eff, int *__restrict__ first, int *__restrict__ m,
                 const int *__restrict__ ctl) {
    if (!(ctl[0] | ctl[1])) return;              // the fast path produced the result
    // (a fast unit that overflowed may already have added its count)
    for (int s = blockIdx.x * blockDim.x + threadIdx.x; s < n_seg; s += gridDim.x * blockDim.x) m[s] = 0;
    const int64_t used = tbl_off[n_seg];
    const uint4 e = make_uint4(0xffffffffu, 0xffffffffu, 0xffffffffu, 0xffffffffu);
    uint4 *t = reinterpret_cast<uint4 *>(tbl);
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < used; i += stride) t[i] = e;
    const int64_t total = *total_eff;            // first[] receives atomicMin in the inverse pass
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += stride) first[i] = 0x7fffffff;
}

// ---- min pass -----------------------------------------------------------------------------
// (a 4-points-per-thread variant of this kernel was measured 2x slower: 97 vs 47 us)
__global__ void __launch_bounds__(VOX_THREADS)
vox_min_kernel(const Xyz xyz, const int64_t *__restrict__ seg_off, int n_seg,
               const int64_t *__restrict__ total_eff, const double *__restrict__ rt, int *__restrict__ grid_min,
               int4 *__restrict__ pgrid, int *status) {
    __shared__ int s_min[VOX_THREADS / 32][3];
    __shared__ int s_pair[2];
    const int64_t total = *total_eff;
    const int64_t b0 = (int64_t)blockIdx.x * blockDim.x;
    if (b0 >= total) return;
    const int64_t i = b0 + threadIdx.x;
    const bool valid = i < total;
    const BlockSeg bs = block_segment(seg_off, n_seg, total, i, valid, s_pair);
    const int s = bs.s;
    int g[3] = {0x7fffffff, 0x7fffffff, 0x7fffffff};
    if (valid) {
        double f[3];
        grid_of(xyz, i, rt + 12 * s, f);
#pragma unroll
        for (int j = 0; j < 3; ++j) {
            if (!(f[j] > -(double)GRID_LIMIT && f[j] < (double)GRID_LIMIT)) {
                if (status) atomicOr(status, XM3D_FLAG_GRID_RANGE);
                f[j] = 0.0;
            }
            g[j] = (int)f[j];
        }
        if (pgrid) pgrid[i] = make_int4(g[0], g[1], g[2], 0);     // read by the fast path
    }
    if (bs.uniform) {                  // warp-shuffle min, then one atomic per block and column
#pragma unroll
        for (int j = 0; j < 3; ++j)
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) g[j] = min(g[j], __shfl_xor_sync(0xffffffffu, g[j], o));
        if (lane_id() == 0) {
            s_min[threadIdx.x >> 5][0] = g[0]; s_min[threadIdx.x >> 5][1] = g[1]; s_min[threadIdx.x >> 5][2] = g[2];
        }
        __syncthreads();
        if (threadIdx.x < 3) {
            int v = 0x7fffffff;
            for (int w = 0; w < VOX_THREADS / 32; ++w) v = min(v, s_min[w][threadIdx.x]);
            atomicMin(&grid_min[3 * s + threadIdx.x], v);
        }
    } else if (valid) {
        atomicMin(&grid_min[3 * s + 0], g[0]);
        atomicMin(&grid_min[3 * s + 1], g[1]);
        atomicMin(&grid_min[3 * s + 2], g[2]);
    }
}

// ---- sample pass --------------------------------------------------------------------------
// Bitonic sort of n keys in shared memory (n a power of two >= 32, blockDim.x a multiple of 32).
// Strides >= 32 are shared-memory steps with a block barrier each; all strides < 32 of a merge
// size run in registers (element e sits in lane e & 31, partners are exchanged with shuffles), so
// a 1024-key sort needs 21 barriers instead of 55.
template <typename K>
__device__ __forceinline__ void bitonic_smem(K *k, int n /*pow2*/) {
    const int lane = threadIdx.x & 31;
    for (int size = 2; size <= n; size <<= 1) {
        int stride = size >> 1;
        for (; stride >= 32; stride >>= 1) {
            for (int t = threadIdx.x; t < (n >> 1); t += blockDim.x) {
                const int lo = 2 * t - (t & (stride - 1));     // insert a 0 at bit log2(stride)
                const int hi = lo + stride;
                const bool up = (lo & size) == 0;
                const K a = k[lo], b = k[hi];
                if ((a > b) == up) { k[lo] = b; k[hi] = a; }
            }
            __syncthreads();
        }
        // register pass: merge sizes 2..32 in one go the first time, then strides 16..1 of `size`
        const int top = size < 32 ? 32 : size;
        for (int e = threadIdx.x; e < n; e += blockDim.x) {
            K v = k[e];
            for (int sz = size; sz <= top; sz <<= 1) {
                const bool up = (e & sz) == 0;
                for (int st = (sz < 32 ? sz : 32) >> 1; st > 0; st >>= 1) {
                    const K o = __shfl_xor_sync(0xffffffffu, v, st);
                    const bool lower = (lane & st) == 0;
                    const bool take_min = lower == up;
                    v = ((o < v) == take_min) ? o : v;
                }
            }
            k[e] = v;
        }
        if (size < 32) size = 32;
        __syncthreads();
    }
}

// KEY_SRC 0: keys from xyz through the transform (voxelize); 1: keys given (unique_batch)
template <int KEY_SRC>
__global__ void __launch_bounds__(1024)
vox_sample_kernel(const Xyz xyz, const unsigned long long *__restrict__ keys_in,
                  const int64_t *__restrict__ seg_off, const int64_t *__restrict__ total_eff,
                  const double *__restrict__ rt, const int *__restrict__ grid_min,
                  const int64_t *__restrict__ spl_off, unsigned long long *__restrict__ spl, int *__restrict__ hist, const int *__restrict__ ctl) {
    if (!(ctl[0] | ctl[1])) return;              // the fast path produced the result
    __shared__ unsigned long long s_key[SPL_MAX];
    const int s = blockIdx.x, tid = threadIdx.x;
    const int64_t a = seg_off[s];
    const int64_t n = (*total_eff > 0) ? seg_off[s + 1] - a : 0;
    const int S = (int)(spl_off[s + 1] - spl_off[s]);
    for (int j = tid; j < S; j += blockDim.x) {
        unsigned long long key = KEY_EMPTY;
        if (n > 0) {
            const int64_t i = a + (int64_t)j * n / S;          // j < 4096, n < 2^31
            key = (KEY_SRC == 0) ? point_key(xyz, i, s, rt, grid_min) : keys_in[i];
            if (key == KEY_EMPTY) key = KEY_EMPTY - 1;
        }
        s_key[j] = key;
        hist[spl_off[s] + j] = 0;
    }
    __syncthreads();
    bitonic_smem(s_key, S);
    for (int j = tid; j < S; j += blockDim.x) spl[spl_off[s] + j] = s_key[j];
}

// ---- insert pass --------------------------------------------------------------------------
template <int KEY_SRC>
__global__ void __launch_bounds__(VOX_THREADS)
vox_insert_kernel(const Xyz xyz, const unsigned long long *__restrict__ keys_in,
                  const int64_t *__restrict__ seg_off, int n_seg, const int64_t *__restrict__ total_eff,
                  const double *__restrict__ rt, const int *__restrict__ grid_min, Slot *__restrict__ tbl,
                  const int64_t *__restrict__ tbl_off, const int64_t *__restrict__ spl_off,
                  const unsigned long long *__restrict__ spl, int *__restrict__ hist,
                  unsigned int *__restrict__ pslot, unsigned int *__restrict__ uniq,
                  unsigned long long *__restrict__ ukey, unsigned int *__restrict__ ubkt,
                  unsigned int *__restrict__ upos, int *__restrict__ m, int *status, const int *__restrict__ ctl) {
    if (!(ctl[0] | ctl[1])) return;              // the fast path produced the result
    __shared__ int s_pair[2];
    const int64_t total = *total_eff;
    if ((int64_t)blockIdx.x * blockDim.x >= total) return;
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const bool valid = i < total;
    const int s = block_segment(seg_off, n_seg, total, i, valid, s_pair).s;
    const unsigned act = __ballot_sync(0xffffffffu, valid);
    if (!valid) return;
    const int lane = lane_id();
    const int64_t base = seg_off[s];
    unsigned long long key = (KEY_SRC == 0) ? point_key(xyz, i, s, rt, grid_min) : keys_in[i];
    key = clean_key(key, status);
    // warp-cooperative de-duplication: one prober per distinct (segment, key) in the warp
    const unsigned peers = __match_any_sync(act, key) & __match_any_sync(act, s);
    const int leader = __ffs(peers) - 1;
    unsigned int slot = 0;
    bool is_new = false;
    if (lane == leader) {
        Slot *tb = tbl + tbl_off[s];
        const unsigned int npairs = (unsigned int)((tbl_off[s + 1] - tbl_off[s]) >> 1);
        // Probe aligned PAIRS of slots: both keys sit in one 32-byte sector, so one round trip
        // examines two slots.  Invariant: a key goes into the second slot of a pair only while the
        // first is occupied, and slots never empty again — an empty first slot ends the search.
        unsigned int pair = (unsigned int)(((mix64(key) >> 32) * (unsigned long long)npairs) >> 32);
        for (unsigned int probe = 0; probe <= npairs; ++probe) {
            Slot *p0 = tb + 2 * pair;
            unsigned long long k0 = *reinterpret_cast<volatile unsigned long long *>(&p0[0].key);
            unsigned long long k1 = *reinterpret_cast<volatile unsigned long long *>(&p0[1].key);
            if (k0 == key) { slot = 2 * pair; break; }
            if (k0 == KEY_EMPTY) {
                k0 = atomicCAS(&p0[0].key, KEY_EMPTY, key);
                if (k0 == KEY_EMPTY) { is_new = true; slot = 2 * pair; break; }
                if (k0 == key) { slot = 2 * pair; break; }
                k1 = *reinterpret_cast<volatile unsigned long long *>(&p0[1].key);   // lost the race: look again
            }
            if (k1 == key) { slot = 2 * pair + 1; break; }
            if (k1 == KEY_EMPTY) {
                k1 = atomicCAS(&p0[1].key, KEY_EMPTY, key);
                if (k1 == KEY_EMPTY) { is_new = true; slot = 2 * pair + 1; break; }
                if (k1 == key) { slot = 2 * pair + 1; break; }
            }
            pair = (pair + 1 == npairs) ? 0u : pair + 1;
        }
    }
    slot = __shfl_sync(act, slot, leader);
    pslot[i] = slot;
    // new unique keys: bucket ticket + append to the unique list (one atomicAdd per warp when the
    // warp sits inside one segment)
    unsigned int bkt = 0, pos = 0;
    if (is_new) {
        const int64_t so = spl_off[s];
        // (the splitters of a segment stay hot in L1: staging them in shared memory was measured
        // 90 us slower — the 32 KB per CTA shrink the L1 that serves the table and xyz reads)
        bkt = (unsigned int)bucket_of(spl + so, (int)(spl_off[s + 1] - so), key);
        pos = (unsigned int)atomicAdd(&hist[so + bkt], 1);
    }
    const unsigned newm = __ballot_sync(act, is_new);
    const unsigned sameseg = __match_any_sync(act, s);
    int64_t up = -1;
    if (sameseg == act) {
        if (newm) {
            int start = 0;
            const int first_lane = __ffs(act) - 1;
            if (lane == first_lane) start = atomicAdd(&m[s], __popc(newm));
            start = __shfl_sync(act, start, first_lane);
            if (is_new) up = base + start + __popc(newm & ((1u << lane) - 1u));
        }
    } else if (is_new) {
        up = base + atomicAdd(&m[s], 1);
    }
    if (up >= 0) {
        uniq[up] = slot;
        ukey[up] = key;
        ubkt[up] = bkt;
        upos[up] = pos;
    }
}

// ---- bucket scan: one CTA per segment -----------------------------------------------------
__global__ void __launch_bounds__(256)
vox_bscan_kernel(const int64_t *__restrict__ spl_off, int n_seg, int *__restrict__ hist, const int *__restrict__ m,
                 int64_t *__restrict__ uniq_off, const int *__restrict__ ctl) {
    if (!(ctl[0] | ctl[1])) return;              // the fast path produced the result
    __shared__ int s_w[8];
    __shared__ int64_t s_red[8];
    __shared__ int s_carry;
    const int s = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    // exclusive prefix of the unique counts of the segments before this one
    int64_t part = 0;
    for (int t = tid; t < s; t += 256) part += m[t];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) part += __shfl_xor_sync(0xffffffffu, part, o);
    if (lane == 0) s_red[warp] = part;
    if (tid == 0) s_carry = 0;
    __syncthreads();
    if (tid == 0) {
        int64_t ub = 0;
        for (int w = 0; w < 8; ++w) ub += s_red[w];
        uniq_off[s] = ub;
        if (s == n_seg - 1) uniq_off[n_seg] = ub + m[s];
    }
    int *h = hist + spl_off[s];
    const int S = (int)(spl_off[s + 1] - spl_off[s]);
    for (int base = 0; base < S; base += 256) {
        const int val = (base + tid < S) ? h[base + tid] : 0;
        int incl = val;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int t = __shfl_up_sync(0xffffffffu, incl, o);
            if (lane >= o) incl += t;
        }
        if (lane == 31) s_w[warp] = incl;
        __syncthreads();
        if (warp == 0) {
            const int w = lane < 8 ? s_w[lane] : 0;
            int wi = w;
#pragma unroll
            for (int o = 1; o < 8; o <<= 1) {
                const int t = __shfl_up_sync(0xffffffffu, wi, o);
                if (lane >= o) wi += t;
            }
            if (lane < 8) s_w[lane] = wi - w;
        }
        __syncthreads();
        const int excl = s_carry + s_w[warp] + incl - val;
        if (base + tid < S) h[base + tid] = excl;
        __syncthreads();
        if (tid == 255) s_carry = excl + val;
        __syncthreads();
    }
}

// ---- scatter unique keys into bucket order ------------------------------------------------
__global__ void __launch_bounds__(VOX_THREADS)
vox_scatter_kernel(const int64_t *__restrict__ seg_off, int n_seg, const int64_t *__restrict__ total_eff,
                   const int *__restrict__ m, const int64_t *__restrict__ spl_off, const int *__restrict__ hist,
                   const unsigned int *__restrict__ uniq, const unsigned long long *__restrict__ ukey,
                   const unsigned int *__restrict__ ubkt, const unsigned int *__restrict__ upos,
                   unsigned long long *__restrict__ bk_key, unsigned long long *__restrict__ bk_sb, const int *__restrict__ ctl) {
    if (!(ctl[0] | ctl[1])) return;              // the fast path produced the result
    __shared__ int s_pair[2];
    const int64_t total = *total_eff;
    if ((int64_t)blockIdx.x * blockDim.x >= total) return;
    const int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const bool valid = p < total;
    const int s = block_segment(seg_off, n_seg, total, p, valid, s_pair).s;
    if (!valid) return;
    const int64_t base = seg_off[s];
    if (p - base >= m[s]) return;                       // the unique list of a segment has m[s] entries
    const unsigned int b = ubkt[p];
    const int64_t d = base + hist[spl_off[s] + b] + upos[p];
    bk_key[d] = ukey[p];
    bk_sb[d] = ((unsigned long long)b << 32) | uniq[p];      // (bucket, table slot)
}

// ---- rank + emit: one thread per unique key, in bucket order --------------------------------
// A block's 256 keys and their buckets sit in a contiguous window of bk_key, staged once in shared
// memory, so the "count the smaller keys of my bucket" loop makes no global memory requests.
constexpr int RANK_PAD = 192;
constexpr int RANK_WIN = VOX_THREADS + 2 * RANK_PAD;

__global__ void __launch_bounds__(VOX_THREADS)
vox_rank_kernel(const int64_t *__restrict__ seg_off, int n_seg, const int64_t *__restrict__ total_eff,
                const int *__restrict__ m, const int64_t *__restrict__ spl_off, const int *__restrict__ hist,
                const unsigned long long *__restrict__ bk_key, const unsigned long long *__restrict__ bk_sb,
                Slot *__restrict__ tbl, const int64_t *__restrict__ tbl_off, const int *__restrict__ ctl) {
    if (!(ctl[0] | ctl[1])) return;              // the fast path produced the result
    __shared__ int s_pair[2];
    __shared__ unsigned long long s_win[RANK_WIN];
    const int64_t total = *total_eff;
    const int64_t p0 = (int64_t)blockIdx.x * blockDim.x;
    if (p0 >= total) return;
    const int64_t p = p0 + threadIdx.x;
    const bool valid = p < total;
    const BlockSeg bs = block_segment(seg_off, n_seg, total, p, valid, s_pair);
    const int s = bs.s;
    const int64_t base = seg_off[s];
    const int M = m[s];
    // window of bucket-ordered keys around the block (only when the block lies in one segment)
    int64_t wlo = 0, whi = 0;
    if (bs.uniform) {
        wlo = (p0 - RANK_PAD > base) ? p0 - RANK_PAD : base;
        whi = (p0 + VOX_THREADS + RANK_PAD < base + M) ? p0 + VOX_THREADS + RANK_PAD : base + M;
        for (int64_t j = wlo + threadIdx.x; j < whi; j += VOX_THREADS) s_win[j - wlo] = bk_key[j];
    }
    __syncthreads();
    if (!valid || p - base >= M) return;
    const unsigned long long key = bk_key[p];
    const unsigned long long sb = bk_sb[p];
    const int b = (int)(sb >> 32);
    const int S = (int)(spl_off[s + 1] - spl_off[s]);
    const int *h = hist + spl_off[s];
    const int lo = h[b], hi = (b + 1 < S) ? h[b + 1] : M;
    int smaller = 0;
    for (int q = lo; q < hi; ++q) {
        const int64_t g = base + q;
        const unsigned long long kq = (g >= wlo && g < whi) ? s_win[g - wlo] : bk_key[g];
        smaller += (kq < key) ? 1 : 0;
    }
    // read back by the inverse pass, which also resolves the first-occurrence index
    tbl[tbl_off[s] + (unsigned int)sb].rank = (unsigned int)(lo + smaller);
}

// ---- inverse pass -------------------------------------------------------------------------
// inverse[i] = rank of point i's key; first[rank] = min over the points of that key of the index
// inside the segment (np.unique's return_index), resolved here with one reduction per point.
__global__ void __launch_bounds__(VOX_THREADS)
vox_inverse_kernel(const int64_t *__restrict__ seg_off, int n_seg, const int64_t *__restrict__ total_eff,
                   const Slot *__restrict__ tbl, const int64_t *__restrict__ tbl_off,
                   const unsigned int *__restrict__ pslot, const int64_t *__restrict__ uniq_off, int collate,
                   int *__restrict__ inverse, int *__restrict__ first, int *__restrict__ counts, const int *__restrict__ ctl) {
    if (!(ctl[0] | ctl[1])) return;              // the fast path produced the result
    __shared__ int s_pair[2];
    const int64_t total = *total_eff;
    if ((int64_t)blockIdx.x * blockDim.x >= total) return;
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const bool valid = i < total;
    const int s = block_segment(seg_off, n_seg, total, i, valid, s_pair).s;
    if (!valid) return;
    const int rank = (int)tbl[tbl_off[s] + pslot[i]].rank;
    const int64_t o = uniq_off[s] + rank;
    if (inverse) inverse[i] = rank + (collate ? (int)uniq_off[s] : 0);
    atomicMin(&first[o], (int)(i - seg_off[s]));
    if (counts) atomicAdd(&counts[o], 1);
}

// ---- voxel coordinates in unique order: grid(first occurrence) - min ------------------------
__global__ void __launch_bounds__(VOX_THREADS)
vox_coords_kernel(const int64_t *__restrict__ seg_off, int n_seg, const int64_t *__restrict__ uniq_off,
                  const int *__restrict__ first, const Xyz xyz, const double *__restrict__ rt,
                  const int *__restrict__ grid_min, int *__restrict__ voxel_xyz, const int *__restrict__ ctl) {
    if (!(ctl[0] | ctl[1])) return;              // the fast path produced the result
    __shared__ int s_pair[2];
    const int64_t total = uniq_off[n_seg];
    if ((int64_t)blockIdx.x * blockDim.x >= total) return;
    const int64_t o = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const bool valid = o < total;
    const int s = block_segment(uniq_off, n_seg, total, o, valid, s_pair).s;
    if (!valid) return;
    double g[3];
    grid_of(xyz, seg_off[s] + first[o], rt + 12 * s, g);
#pragma unroll
    for (int j = 0; j < 3; ++j) {
        if (!(g[j] > -(double)GRID_LIMIT && g[j] < (double)GRID_LIMIT)) g[j] = 0.0;
        voxel_xyz[o * 3 + j] = (int)g[j] - grid_min[3 * s + j];
    }
}

// ---- fast path: one persistent CTA per unit, everything in shared memory ----------------------
static_assert(FV_TABLE == 1 << 14, "fv_hash yields 14 bits");
static_assert(FV_TABLE / FV_THREADS <= 32 && FV_TABLE % FV_THREADS == 0, "table scan: one bit per slot");
__device__ __forceinline__ unsigned int fv_hash(unsigned long long key) {
    const unsigned int x = (unsigned int)key ^ (unsigned int)(key >> 32);
    return (x * 0x9E3779B1u) >> 18;
}

// key of an element: FNV-1 of (grid - min) from the grid coordinates the min pass stored, or the given key.
// fv_load / fv_make are separate so that a thread can put several loads in flight before it hashes.
// (pgrid is read with plain loads: the owning unit tags .w in place during the insert pass.)
struct FvRaw { int4 g; unsigned long long k; };
template <int KEY_SRC>
__device__ __forceinline__ FvRaw fv_load(const int4 *pgrid, const unsigned long long *__restrict__ keys_in, int64_t i) {
    FvRaw r;
    if (KEY_SRC == 0) { r.g = pgrid[i]; r.k = 0; }
    else { r.g = make_int4(0, 0, 0, 0); r.k = keys_in[i]; }
    return r;
}
template <int KEY_SRC>
__device__ __forceinline__ unsigned long long fv_make(const FvRaw &r, int g0, int g1, int g2, int *status) {
    unsigned long long key;
    if (KEY_SRC == 0)
        key = fnv3((unsigned long long)(long long)(r.g.x - g0), (unsigned long long)(long long)(r.g.y - g1),
                   (unsigned long long)(long long)(r.g.z - g2));
    else
        key = r.k;
    return clean_key(key, status);
}
template <int KEY_SRC>
__device__ __forceinline__ unsigned long long fv_key(const int4 *pgrid, const unsigned long long *__restrict__ keys_in,
                                                     int64_t i, int g0, int g1, int g2, int *status) {
    return fv_make<KEY_SRC>(fv_load<KEY_SRC>(pgrid, keys_in, i), g0, g1, g2, status);
}

#ifdef XM3D_FV_TIMING
__device__ long long *g_fv_dbg = nullptr;
#define FV_T(i) do { if (threadIdx.x == 0 && g_fv_dbg) { g_fv_dbg[blockIdx.x * 32 + (i)] = clock64(); \
    if ((i) == 0 || (i) == 8) { unsigned long long gt; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(gt)); \
                                g_fv_dbg[blockIdx.x * 32 + ((i) == 0 ? 28 : 29)] = (long long)gt; } } } while (0)
#else
#define FV_T(i) do { } while (0)
#endif

template <int KEY_SRC>
__global__ void __launch_bounds__(FV_THREADS, 65536 / (FV_THREADS * 64))
vox_fast_kernel(int4 *pgrid, const unsigned long long *__restrict__ keys_in,
                const int64_t *__restrict__ seg_off, int n_seg, const int64_t *__restrict__ total_eff,
                const int *__restrict__ unit_off, const int *__restrict__ unit_seg, int *unit_m, int *ctl,
                const int *__restrict__ grid_min, unsigned int *__restrict__ pslot, int *__restrict__ m,
                int64_t *__restrict__ uniq_off, int *__restrict__ first, int *__restrict__ inverse, int collate,
                int *__restrict__ voxel_xyz, int *status) {
    extern __shared__ __align__(16) unsigned char fv_smem[];
    unsigned long long *s_tab = reinterpret_cast<unsigned long long *>(fv_smem);     // keys, later ranks
    unsigned int *s_ub = reinterpret_cast<unsigned int *>(s_tab + FV_TABLE);          // bucket << 16 | ticket; later
                                                                                      // ranks (u16), first index (int)
    unsigned short *s_uslot = reinterpret_cast<unsigned short *>(s_ub + FV_MU);       // slot of unique u (insertion order)
    unsigned int *s_bk = reinterpret_cast<unsigned int *>(s_uslot + FV_MU);           // bucket order: bucket << 16 | slot
    unsigned long long *s_spl = reinterpret_cast<unsigned long long *>(s_bk + FV_MU); // sample keys / splitters
    int *s_hist = reinterpret_cast<int *>(s_spl + FV_NS);                             // [S + 1]
    __shared__ int s_cnt, s_ovf;
    __shared__ long long s_red[2][32];
    __shared__ int s_wsum[32];

    if (ctl[0]) return;                           // not eligible: the multi-kernel path runs
    // The unit id is a TICKET, not blockIdx.x: the look-back below waits for the units t < u, and CUDA does not
    // promise to dispatch blocks in index order.  With a ticket a waiting CTA only ever depends on CTAs that
    // have already started (CUB's dynamic tile id).
    __shared__ int s_unit;
    if (threadIdx.x == 0) s_unit = atomicAdd(&ctl[2], 1);
    __syncthreads();
    const int u = s_unit;
    const int U = unit_off[n_seg];
    if (u >= U) return;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int s = unit_seg[u];
    const int u0 = unit_off[s];
    const int p = u - u0, P = unit_off[s + 1] - u0;
    const int64_t a = seg_off[s];
    const int n = (*total_eff > 0) ? (int)(seg_off[s + 1] - a) : 0;
    FV_T(0);
    int g0 = 0, g1 = 0, g2 = 0;
    if (KEY_SRC == 0) { g0 = grid_min[3 * s]; g1 = grid_min[3 * s + 1]; g2 = grid_min[3 * s + 2]; }

    // key range of this unit: [klo, khi) between the P-quantiles of a 1024-key sample (every unit
    // of the segment computes the same sample); KEY_EMPTY is never a key, so it closes the last range
    unsigned long long klo = 0, khi = KEY_EMPTY;
    if (P > 1) {
        const int ns = P <= 8 ? FV_NS / 4 : FV_NS;          // >= 32 sample keys per unit
        for (int j = tid; j < ns; j += FV_THREADS)
            s_spl[j] = fv_key<KEY_SRC>(pgrid, keys_in, a + (int64_t)j * n / ns, g0, g1, g2, nullptr);
        __syncthreads();
        bitonic_smem(s_spl, ns);
        if (p > 0) klo = s_spl[(int)((int64_t)p * ns / P)];
        if (p < P - 1) khi = s_spl[(int)((int64_t)(p + 1) * ns / P)];
        __syncthreads();
    }
    FV_T(1);
    for (int j = tid; j < FV_TABLE / 2; j += FV_THREADS)
        reinterpret_cast<uint4 *>(s_tab)[j] = make_uint4(0xffffffffu, 0xffffffffu, 0xffffffffu, 0xffffffffu);
    if (tid == 0) { s_cnt = 0; s_ovf = 0; }
    __syncthreads();

    FV_T(2);
    // ---- insert: open addressing in shared memory (the unique list is read off the table afterwards)
    volatile unsigned long long *vtab = s_tab;
    const unsigned int tag = (unsigned int)(p + 1) << 16;          // owner mark in pgrid[i].w (slot in the low half)
    for (int base = 0; base < n; base += FV_BATCH * FV_THREADS) {
        if (*reinterpret_cast<volatile int *>(&s_ovf)) break;      // overflow: the batch is recomputed anyway
        FvRaw raw[FV_BATCH];
#pragma unroll
        for (int k = 0; k < FV_BATCH; ++k) {
            const int j = base + k * FV_THREADS + tid;
            if (j < n) raw[k] = fv_load<KEY_SRC>(pgrid, keys_in, a + j);
        }
#pragma unroll
        for (int k = 0; k < FV_BATCH; ++k) {
            const int j = base + k * FV_THREADS + tid;
            if (j >= n) continue;
            const unsigned long long key = fv_make<KEY_SRC>(raw[k], g0, g1, g2, status);
            if (!(key >= klo && key < khi)) continue;
            unsigned int h = fv_hash(key);
            int probe = 0;
            for (; probe < FV_PROBE_MAX; ++probe) {
                unsigned long long cur = vtab[h];
                if (cur == key) break;
                if (cur == KEY_EMPTY) {
                    cur = atomicCAS(&s_tab[h], KEY_EMPTY, key);
                    if (cur == KEY_EMPTY || cur == key) break;
                }
                h = (h + 1) & (FV_TABLE - 1);
            }
            if (probe == FV_PROBE_MAX) s_ovf = 1;
            if (KEY_SRC == 0) reinterpret_cast<unsigned int *>(pgrid + a + j)[3] = tag | h;
            else pslot[a + j] = h;
        }
    }
    __syncthreads();
    FV_T(3);
    // ---- unique list: the occupied slots, found by scanning the table (no shared counter in the insert loop)
    {
        constexpr int SPT = FV_TABLE / FV_THREADS;                 // slots per thread
        unsigned int occ = 0;
#pragma unroll
        for (int i = 0; i < SPT; ++i) occ |= (s_tab[i * FV_THREADS + tid] != KEY_EMPTY ? 1u : 0u) << i;   // conflict-free
        const int val = __popc(occ);
        int incl = val;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int t = __shfl_up_sync(0xffffffffu, incl, o);
            if (lane >= o) incl += t;
        }
        if (lane == 31) s_wsum[warp] = incl;
        __syncthreads();
        if (warp == 0) {
            const int w = lane < FV_THREADS / 32 ? s_wsum[lane] : 0;
            int wi = w;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const int t = __shfl_up_sync(0xffffffffu, wi, o);
                if (lane >= o) wi += t;
            }
            s_wsum[lane] = wi - w;
            if (lane == 31) s_cnt = wi;
        }
        __syncthreads();
        int ui = s_wsum[warp] + incl - val;
        if (s_cnt <= FV_MU) {
            while (occ) {
                const int i = __ffs(occ) - 1;
                occ &= occ - 1;
                s_uslot[ui++] = (unsigned short)(i * FV_THREADS + tid);
            }
        }
    }
    __syncthreads();
    FV_T(9);
    const int M = s_cnt;
    if (M > FV_MU || s_ovf) {                     // does not fit: publish (nobody may wait for ever) and fall back
        if (tid == 0) {
            atomicExch(&ctl[1], 1);
            atomicExch(&unit_m[u], 0);
            if (ctl[3] && status) atomicOr(status, XM3D_FLAG_VOX_FALLBACK);   // nobody will recompute the batch
        }
        return;
    }
    if (tid == 0) atomicExch(&unit_m[u], M);

    // ---- rank of every unique key: sample -> splitters -> bucket tickets -> count inside the bucket
    if (M > 0) {
        int S = 32;
        while (S < FV_NS && S < FV_THREADS && S * 8 < M) S <<= 1;
        for (int j = tid; j < S; j += FV_THREADS) s_spl[j] = s_tab[s_uslot[(int)((int64_t)j * M / S)]];
        for (int j = tid; j <= S; j += FV_THREADS) s_hist[j] = 0;
        __syncthreads();
        bitonic_smem(s_spl, S);
        FV_T(10);
        int logS = 5;
        while ((1 << logS) < S) ++logS;
        // four independent binary searches per thread (the chain of dependent shared-memory loads is
        // what bounds this phase); S is a power of two, so every search takes exactly log2 S steps
        for (int q0 = tid; q0 < M; q0 += FV_BATCH * FV_THREADS) {
            unsigned long long key[FV_BATCH];
            int lo[FV_BATCH], hi[FV_BATCH];
#pragma unroll
            for (int k = 0; k < FV_BATCH; ++k) {
                const int q = q0 + k * FV_THREADS;
                key[k] = q < M ? s_tab[s_uslot[q]] : 0ull;
                lo[k] = 0; hi[k] = S - 1;
            }
            for (int step = 0; step < logS; ++step) {
#pragma unroll
                for (int k = 0; k < FV_BATCH; ++k) {
                    const int mid = (lo[k] + hi[k]) >> 1;
                    if (s_spl[mid + 1] < key[k]) lo[k] = mid + 1; else hi[k] = mid;
                }
            }
#pragma unroll
            for (int k = 0; k < FV_BATCH; ++k) {
                const int q = q0 + k * FV_THREADS;
                if (q < M) {
                    const int pos = atomicAdd(&s_hist[lo[k]], 1);
                    s_ub[q] = ((unsigned int)lo[k] << 16) | (unsigned int)pos;
                }
            }
        }
        __syncthreads();
        FV_T(11);
        {   // exclusive scan of the S <= 1024 bucket counts, one per thread
            const int val = tid < S ? s_hist[tid] : 0;
            int incl = val;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const int t = __shfl_up_sync(0xffffffffu, incl, o);
                if (lane >= o) incl += t;
            }
            if (lane == 31) s_wsum[warp] = incl;
            __syncthreads();
            if (warp == 0) {
                const int w = lane < FV_THREADS / 32 ? s_wsum[lane] : 0;
                int wi = w;
#pragma unroll
                for (int o = 1; o < 32; o <<= 1) {
                    const int t = __shfl_up_sync(0xffffffffu, wi, o);
                    if (lane >= o) wi += t;
                }
                s_wsum[lane] = wi - w;
            }
            __syncthreads();
            if (tid < S) s_hist[tid] = s_wsum[warp] + incl - val;
            if (tid == 0) s_hist[S] = M;
        }
        __syncthreads();
        FV_T(12);
        for (int q = tid; q < M; q += FV_THREADS) {
            const unsigned int ub = s_ub[q];
            s_bk[s_hist[ub >> 16] + (int)(ub & 0xffffu)] = (ub & 0xffff0000u) | s_uslot[q];
        }
        __syncthreads();
        FV_T(13);
        unsigned short *s_rank = reinterpret_cast<unsigned short *>(s_ub);
        for (int d = tid; d < M; d += FV_THREADS) {
            const unsigned int sb = s_bk[d];
            const int b = (int)(sb >> 16);
            const unsigned long long key = s_tab[sb & 0xffffu];
            const int lo = s_hist[b], hi = s_hist[b + 1];
            int smaller = 0;
            int q = lo;
            for (; q + 3 < hi; q += 4) {            // four mates in flight
                const unsigned int b0 = s_bk[q], b1 = s_bk[q + 1], b2 = s_bk[q + 2], b3 = s_bk[q + 3];
                const unsigned long long k0 = s_tab[b0 & 0xffffu], k1 = s_tab[b1 & 0xffffu];
                const unsigned long long k2 = s_tab[b2 & 0xffffu], k3 = s_tab[b3 & 0xffffu];
                smaller += (k0 < key) + (k1 < key) + (k2 < key) + (k3 < key);
            }
            for (; q < hi; ++q) smaller += (s_tab[s_bk[q] & 0xffffu] < key) ? 1 : 0;
            s_rank[d] = (unsigned short)(lo + smaller);
        }
        __syncthreads();
        FV_T(14);
        for (int d = tid; d < M; d += FV_THREADS) s_tab[s_bk[d] & 0xffffu] = s_rank[d];   // keys are not needed any more
        __syncthreads();
    }

    FV_T(4);
    // ---- decoupled look-back: unique counts of all earlier units (blocks are dispatched in index order)
    long long sum_all = 0, sum_seg = 0;
    for (int t = tid; t < u; t += FV_THREADS) {
        int v;
        unsigned spin = 0;
        while ((v = *reinterpret_cast<volatile int *>(&unit_m[t])) < 0) {
            __nanosleep(64);
            if (++spin > (1u << 24)) __trap();
        }
        sum_all += v;
        if (t >= u0) sum_seg += v;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        sum_all += __shfl_xor_sync(0xffffffffu, sum_all, o);
        sum_seg += __shfl_xor_sync(0xffffffffu, sum_seg, o);
    }
    if (lane == 0) { s_red[0][warp] = sum_all; s_red[1][warp] = sum_seg; }
    __syncthreads();
    sum_all = 0; sum_seg = 0;
    for (int w = 0; w < FV_THREADS / 32; ++w) { sum_all += s_red[0][w]; sum_seg += s_red[1][w]; }
    FV_T(5);
    const int64_t uo = sum_all - sum_seg;          // uniq_off[s]
    const int rbase = (int)sum_seg;                // rank of this unit's smallest key inside the segment
    if (tid == 0) {
        if (p == 0) uniq_off[s] = uo;
        if (u == U - 1) uniq_off[n_seg] = sum_all + M;
        if (M > 0) atomicAdd(&m[s], M);
    }
    if (M == 0) return;

    // ---- inverse map and first occurrence
    int *s_first = reinterpret_cast<int *>(s_ub);
    for (int r = tid; r < M; r += FV_THREADS) s_first[r] = 0x7fffffff;
    __syncthreads();
    FV_T(6);
    const int add = rbase + (collate ? (int)uo : 0);
    for (int base = 0; base < n; base += FV_BATCH * FV_THREADS) {
        unsigned int w[FV_BATCH];
        unsigned long long kk[FV_BATCH];
#pragma unroll
        for (int k = 0; k < FV_BATCH; ++k) {
            const int j = base + k * FV_THREADS + tid;
            w[k] = 0; kk[k] = 0;
            if (j < n) {
                if (KEY_SRC == 0) w[k] = reinterpret_cast<const unsigned int *>(pgrid + a + j)[3];
                else { w[k] = pslot[a + j]; if (P > 1) kk[k] = keys_in[a + j]; }
            }
        }
#pragma unroll
        for (int k = 0; k < FV_BATCH; ++k) {
            const int j = base + k * FV_THREADS + tid;
            bool own = j < n;
            if (KEY_SRC == 0) {
                own = own && (w[k] & 0xffff0000u) == tag;      // tagged by this unit's insert pass (same thread)
            } else if (own && P > 1) {
                const unsigned long long key = clean_key(kk[k], nullptr);
                own = key >= klo && key < khi;
            }
            if (own) {
                const int r = (int)s_tab[w[k] & 0xffffu];
                if (inverse) inverse[a + j] = add + r;
                atomicMin(&s_first[r], j);
            }
        }
    }
    __syncthreads();
    FV_T(7);
    for (int r0 = tid; r0 < M; r0 += FV_OUT_BATCH * FV_THREADS) {
        int f[FV_OUT_BATCH];
        int4 g[FV_OUT_BATCH];
#pragma unroll
        for (int k = 0; k < FV_OUT_BATCH; ++k) {
            const int r = r0 + k * FV_THREADS;
            f[k] = r < M ? s_first[r] : 0;
            if (KEY_SRC == 0 && voxel_xyz && r < M) g[k] = pgrid[a + f[k]];     // four gathers in flight
        }
#pragma unroll
        for (int k = 0; k < FV_OUT_BATCH; ++k) {
            const int r = r0 + k * FV_THREADS;
            if (r < M) {
                const int64_t o = uo + rbase + r;
                first[o] = f[k];
                if (KEY_SRC == 0 && voxel_xyz) {
                    voxel_xyz[o * 3 + 0] = g[k].x - g0; voxel_xyz[o * 3 + 1] = g[k].y - g1; voxel_xyz[o * 3 + 2] = g[k].z - g2;
                }
            }
        }
    }
    FV_T(8);
#ifdef XM3D_FV_TIMING
    if (threadIdx.x == 0 && g_fv_dbg) { g_fv_dbg[blockIdx.x * 32 + 24] = n; g_fv_dbg[blockIdx.x * 32 + 25] = M; g_fv_dbg[blockIdx.x * 32 + 26] = P;
        unsigned smid; asm("mov.u32 %0, %%smid;" : "=r"(smid)); g_fv_dbg[blockIdx.x * 32 + 27] = smid; }
#endif
}

// ---- elementwise hashes (drop-ins for fnv_hash_vec / ravel_hash_vec on float64 rows) --------
__global__ void __launch_bounds__(256)
fnv_f64_kernel(const double *__restrict__ coords, int64_t n, int dim, unsigned long long *__restrict__ keys) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    unsigned long long h = FNV_OFFSET;
    for (int j = 0; j < dim; ++j) {
        h *= FNV_PRIME;
        h ^= f64_to_u64_numpy(coords[i * dim + j]);
    }
    keys[i] = h;
}

// column min / max of a float64 [n, dim] matrix (dim <= 8) into ws[0..dim) / ws[8..8+dim)
__global__ void __launch_bounds__(256)
colminmax_f64_kernel(const double *__restrict__ coords, int64_t n, int dim, double *__restrict__ ws) {
    __shared__ double s_min[8][8], s_max[8][8];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    double mn[8], mx[8];
    for (int j = 0; j < 8; ++j) { mn[j] = INFINITY; mx[j] = -INFINITY; }
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
        for (int j = 0; j < dim; ++j) {
            const double v = coords[i * dim + j];
            mn[j] = fmin(mn[j], v); mx[j] = fmax(mx[j], v);
        }
    for (int j = 0; j < dim; ++j) {
        for (int o = 16; o > 0; o >>= 1) {
            mn[j] = fmin(mn[j], __shfl_xor_sync(0xffffffffu, mn[j], o));
            mx[j] = fmax(mx[j], __shfl_xor_sync(0xffffffffu, mx[j], o));
        }
        if (lane == 0) { s_min[warp][j] = mn[j]; s_max[warp][j] = mx[j]; }
    }
    __syncthreads();
    if (threadIdx.x < dim) {
        const int j = threadIdx.x;
        double a = INFINITY, b = -INFINITY;
        for (int w = 0; w < 8; ++w) { a = fmin(a, s_min[w][j]); b = fmax(b, s_max[w][j]); }
        // atomic min / max on doubles through compare-and-swap
        unsigned long long *pa = reinterpret_cast<unsigned long long *>(ws + j);
        unsigned long long old = *pa, assumed;
        do { assumed = old; if (__longlong_as_double(assumed) <= a) break;
             old = atomicCAS(pa, assumed, (unsigned long long)__double_as_longlong(a)); } while (old != assumed);
        unsigned long long *pb = reinterpret_cast<unsigned long long *>(ws + 8 + j);
        old = *pb;
        do { assumed = old; if (__longlong_as_double(assumed) >= b) break;
             old = atomicCAS(pb, assumed, (unsigned long long)__double_as_longlong(b)); } while (old != assumed);
    }
}

__global__ void init_minmax_kernel(double *ws) {
    if (threadIdx.x < 8) { ws[threadIdx.x] = INFINITY; ws[8 + threadIdx.x] = -INFINITY; }
}

// ravel_hash_vec (voxelization_utils.py:21-35): arr -= min; arr = uint64(arr); radix = max + 1;
// key = ((k + a_0) * radix_1 + a_1) * radix_2 + ... + a_last
__global__ void __launch_bounds__(256)
ravel_f64_kernel(const double *__restrict__ coords, int64_t n, int dim, const double *__restrict__ ws,
                 unsigned long long *__restrict__ keys) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    unsigned long long key = 0;
    for (int j = 0; j < dim - 1; ++j) {
        key += f64_to_u64_numpy(__dsub_rn(coords[i * dim + j], ws[j]));
        key *= f64_to_u64_numpy(__dsub_rn(ws[8 + j + 1], ws[j + 1])) + 1ull;
    }
    key += f64_to_u64_numpy(__dsub_rn(coords[i * dim + dim - 1], ws[dim - 1]));
    keys[i] = key;
}

struct VoxWs {
    Slot *tbl;
    int64_t *tbl_off, *spl_off, *total_eff;
    unsigned int *pslot, *uniq, *ubkt, *upos;
    unsigned long long *ukey, *bk_key, *bk_sb, *spl;
    int *hist, *grid_min;
    int4 *pgrid;                       // grid coordinates of every point (fast path, voxel coords)
    int *unit_off, *unit_seg, *unit_m, *ctl;
    int units_cap;
};

static VoxWs carve_vox(void *ws, int n_seg, int64_t cap, size_t *bytes) {
    Carver c(ws);
    VoxWs w;
    // sum of bucket_count(n_s) <= sum max(32, n_s / 8) <= cap / 8 + 32 n_seg
    const size_t nspl = (size_t)(cap / 8 + (int64_t)(SPL_MIN + 8) * n_seg + SPL_MAX);
    w.tbl = c.take<Slot>((size_t)(table_size(cap) + 32 * (int64_t)n_seg));
    w.tbl_off = c.take<int64_t>(n_seg + 1);
    w.spl_off = c.take<int64_t>(n_seg + 1);
    w.total_eff = c.take<int64_t>(1);
    w.pslot = c.take<unsigned int>(cap);
    w.uniq = c.take<unsigned int>(cap);
    w.ubkt = c.take<unsigned int>(cap);
    w.upos = c.take<unsigned int>(cap);
    w.ukey = c.take<unsigned long long>(cap);
    w.bk_key = c.take<unsigned long long>(cap);
    w.bk_sb = c.take<unsigned long long>(cap);
    w.spl = c.take<unsigned long long>(nspl);
    w.hist = c.take<int>(nspl);
    w.grid_min = c.take<int>(3 * (size_t)n_seg);
    w.pgrid = c.take<int4>((size_t)cap);
    w.units_cap = (int)(cap / FV_UNIT_PTS_MIN + n_seg);
    w.unit_off = c.take<int>((size_t)n_seg + 1);
    w.unit_seg = c.take<int>((size_t)w.units_cap);
    w.unit_m = c.take<int>((size_t)w.units_cap);
    w.ctl = c.take<int>(64);
    *bytes = c.off + 256;
    return w;
}

static int run_unique(const Xyz xyz, const unsigned long long *keys, const int64_t *seg_off, int n_seg,
                      int64_t cap, const double *rt, int *m, int64_t *uniq_off, int *first, int *counts,
                      int *inverse, int collate, int *voxel_xyz, int *grid_min_out, void *ws, size_t ws_bytes,
                      int *status, int path, cudaStream_t stream, const char *who) {
    size_t need = 0;
    VoxWs w = carve_vox(ws, n_seg, cap, &need);
    if (ws_bytes < need) {
        set_error("%s: workspace too small (%zu < %zu)", who, ws_bytes, need);
        return XM3D_ERR_WORKSPACE;
    }
    int *gmin = grid_min_out ? grid_min_out : w.grid_min;
    const unsigned blocks = (unsigned)((cap + VOX_THREADS - 1) / VOX_THREADS);
    // fast path: one shared-memory CTA per unit; the multi-kernel path below is gated on w.ctl
    // path: bits 0-7 = XM3D_VOX_AUTO / XM3D_VOX_MULTI_KERNEL, bits 8.. = points per key-range unit (0 = default)
    const int req_pts = path >> 8;
    const int unit_pts = req_pts <= 0 ? FV_UNIT_PTS : (req_pts < FV_UNIT_PTS_MIN ? FV_UNIT_PTS_MIN
                                                      : (req_pts > FV_UNIT_PTS ? FV_UNIT_PTS : req_pts));
    int64_t units_max64 = cap / unit_pts + n_seg;
    if (units_max64 > w.units_cap) units_max64 = w.units_cap;
    const int units_max = (int)units_max64;
    const bool fast_only = (path & 0xff) == XM3D_VOX_FAST_ONLY;
    if (fast_only && counts) { set_error("%s: XM3D_VOX_FAST_ONLY cannot return counts", who); return XM3D_ERR_UNSUPPORTED; }
    const int force_slow = (((path & 0xff) == XM3D_VOX_MULTI_KERNEL || counts != nullptr) ? 1 : 0) | (fast_only ? 2 : 0);
    static std::atomic<uint64_t> smem_set{0};
    if (first_use_on_device(&smem_set)) {
        cudaFuncSetAttribute(vox_fast_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)FV_SMEM);
        cudaFuncSetAttribute(vox_fast_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)FV_SMEM);
    }
    vox_plan_kernel<<<1, 1024, 0, stream>>>(seg_off, n_seg, cap, w.tbl_off, w.spl_off, w.total_eff, m,
                                            xyz ? gmin : nullptr, status, unit_pts, units_max, force_slow,
                                            w.unit_off, w.unit_seg, w.unit_m, w.ctl);
    count_launches(1);
    if (xyz && blocks) {
        vox_min_kernel<<<blocks, VOX_THREADS, 0, stream>>>(xyz, seg_off, n_seg, w.total_eff, rt, gmin, w.pgrid, status);
        count_launches(1);
    }
    if (!(force_slow & 1)) {
        if (xyz)
            vox_fast_kernel<0><<<units_max, FV_THREADS, FV_SMEM, stream>>>(
                w.pgrid, nullptr, seg_off, n_seg, w.total_eff, w.unit_off, w.unit_seg, w.unit_m, w.ctl, gmin, w.pslot, m,
                uniq_off, first, inverse, collate, voxel_xyz, status);
        else
            vox_fast_kernel<1><<<units_max, FV_THREADS, FV_SMEM, stream>>>(
                nullptr, keys, seg_off, n_seg, w.total_eff, w.unit_off, w.unit_seg, w.unit_m, w.ctl, nullptr, w.pslot, m,
                uniq_off, first, inverse, collate, nullptr, status);
        count_launches(1);
    }
    if (fast_only) return check_launch(who);      // the eight kernels of the multi-kernel path are not even launched
    vox_clear_kernel<<<(unsigned)(sm_count() * 8), 256, 0, stream>>>(w.tbl, w.tbl_off, n_seg, w.total_eff, first, m, w.ctl);
    count_launches(1);
    if (counts) cudaMemsetAsync(counts, 0, sizeof(int) * cap, stream);
    if (xyz) {
        vox_sample_kernel<0><<<n_seg, 1024, 0, stream>>>(xyz, nullptr, seg_off, w.total_eff, rt, gmin, w.spl_off, w.spl,
                                                        w.hist, w.ctl);
        count_launches(1);
        if (blocks) {
            vox_insert_kernel<0><<<blocks, VOX_THREADS, 0, stream>>>(xyz, nullptr, seg_off, n_seg, w.total_eff, rt, gmin,
                                                                     w.tbl, w.tbl_off, w.spl_off, w.spl, w.hist, w.pslot,
                                                                     w.uniq, w.ukey, w.ubkt, w.upos, m, status, w.ctl);
            count_launches(1);
        }
    } else {
        vox_sample_kernel<1><<<n_seg, 1024, 0, stream>>>(Xyz{nullptr, nullptr}, keys, seg_off, w.total_eff, nullptr, nullptr, w.spl_off,
                                                        w.spl, w.hist, w.ctl);
        count_launches(1);
        if (blocks) {
            vox_insert_kernel<1><<<blocks, VOX_THREADS, 0, stream>>>(Xyz{nullptr, nullptr}, keys, seg_off, n_seg, w.total_eff, nullptr,
                                                                     nullptr, w.tbl, w.tbl_off, w.spl_off, w.spl, w.hist,
                                                                     w.pslot, w.uniq, w.ukey, w.ubkt, w.upos, m, status, w.ctl);
            count_launches(1);
        }
    }
    vox_bscan_kernel<<<n_seg, 256, 0, stream>>>(w.spl_off, n_seg, w.hist, m, uniq_off, w.ctl);
    count_launches(1);
    if (blocks) {
        vox_scatter_kernel<<<blocks, VOX_THREADS, 0, stream>>>(seg_off, n_seg, w.total_eff, m, w.spl_off, w.hist, w.uniq,
                                                               w.ukey, w.ubkt, w.upos, w.bk_key, w.bk_sb, w.ctl);
        count_launches(1);
        vox_rank_kernel<<<blocks, VOX_THREADS, 0, stream>>>(seg_off, n_seg, w.total_eff, m, w.spl_off, w.hist, w.bk_key,
                                                            w.bk_sb, w.tbl, w.tbl_off, w.ctl);
        count_launches(1);
        vox_inverse_kernel<<<blocks, VOX_THREADS, 0, stream>>>(seg_off, n_seg, w.total_eff, w.tbl, w.tbl_off, w.pslot,
                                                               uniq_off, collate, inverse, first, counts, w.ctl);
        count_launches(1);
        if (xyz && voxel_xyz) {
            vox_coords_kernel<<<blocks, VOX_THREADS, 0, stream>>>(seg_off, n_seg, uniq_off, first, xyz, rt, gmin,
                                                                  voxel_xyz, w.ctl);
            count_launches(1);
        }
    }
    return check_launch(who);
}

}  // namespace xm3d

using namespace xm3d;

extern "C" int xm3d_voxel_path_info(const void *ws, int32_t n_seg, int64_t cap, int32_t *ctl_host,
                                    xm3d_stream_t stream_) {
    XM3D_REQUIRE(ws && ctl_host && n_seg > 0 && cap >= 0, "bad arguments");
    size_t need = 0;
    const VoxWs w = carve_vox(const_cast<void *>(ws), n_seg, cap, &need);
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    if (cudaMemcpyAsync(ctl_host, w.ctl, 2 * sizeof(int), cudaMemcpyDeviceToHost, stream) != cudaSuccess ||
        cudaStreamSynchronize(stream) != cudaSuccess)
        return check_launch("xm3d_voxel_path_info");
    return XM3D_OK;
}

#ifdef XM3D_FV_TIMING
extern "C" __attribute__((visibility("default"))) int xm3d_voxel_debug(void *dev_buf) {
    return (int)cudaMemcpyToSymbol(g_fv_dbg, &dev_buf, sizeof(void *));
}
#endif

extern "C" size_t xm3d_unique_ws_bytes(int32_t n_seg, int64_t cap) {
    size_t b = 0;
    carve_vox(nullptr, n_seg, cap, &b);
    return b;
}
extern "C" size_t xm3d_voxelize_ws_bytes(int32_t n_seg, int64_t cap) { return xm3d_unique_ws_bytes(n_seg, cap); }

extern "C" int xm3d_unique_batch(const uint64_t *keys, const int64_t *seg_off, int32_t n_seg, int64_t cap,
                                 int32_t *m, int64_t *uniq_off, int32_t *first, int32_t *counts,
                                 int32_t *inverse, int32_t collate, int32_t path, void *ws, size_t ws_bytes,
                                 int32_t *status, xm3d_stream_t stream) {
    XM3D_REQUIRE(n_seg > 0 && cap >= 0, "bad sizes");
    XM3D_REQUIRE(keys && seg_off && m && uniq_off && first && ws, "null pointer");
    XM3D_REQUIRE(cap < (int64_t)1 << 31, "cap must fit int32");
    return run_unique(Xyz{nullptr, nullptr}, reinterpret_cast<const unsigned long long *>(keys), seg_off, n_seg, cap,
                      nullptr, m, uniq_off, first, counts, inverse, collate, nullptr, nullptr, ws, ws_bytes, status, path,
                      static_cast<cudaStream_t>(stream), "xm3d_unique_batch");
}

extern "C" int xm3d_voxelize_batch(const void *xyz, int32_t xyz_f64, const int64_t *seg_off, int32_t n_seg, int64_t cap,
                                   const double *rt, int32_t *m, int64_t *uniq_off, int32_t *first,
                                   int32_t *inverse, int32_t collate, int32_t *voxel_xyz, int32_t *grid_min,
                                   int32_t path, void *ws, size_t ws_bytes, int32_t *status, xm3d_stream_t stream) {
    XM3D_REQUIRE(n_seg > 0 && cap >= 0, "bad sizes");
    XM3D_REQUIRE(xyz && seg_off && rt && m && uniq_off && first && ws, "null pointer");
    XM3D_REQUIRE(cap < (int64_t)1 << 31, "cap must fit int32");
    const Xyz src = xyz_f64 ? Xyz{nullptr, static_cast<const double *>(xyz)} : Xyz{static_cast<const float *>(xyz), nullptr};
    return run_unique(src, nullptr, seg_off, n_seg, cap, rt, m, uniq_off, first, nullptr, inverse, collate,
                      voxel_xyz, grid_min, ws, ws_bytes, status, path, static_cast<cudaStream_t>(stream),
                      "xm3d_voxelize_batch");
}

extern "C" int xm3d_fnv_hash_f64(const double *coords, int64_t n, int32_t dim, uint64_t *keys,
                                 xm3d_stream_t stream) {
    XM3D_REQUIRE(n >= 0 && dim > 0, "bad sizes");
    if (n == 0) return XM3D_OK;
    XM3D_REQUIRE(coords && keys, "null pointer");
    fnv_f64_kernel<<<(unsigned)((n + 255) / 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(
        coords, n, dim, reinterpret_cast<unsigned long long *>(keys));
    count_launches(1);
    return check_launch("xm3d_fnv_hash_f64");
}

extern "C" size_t xm3d_ravel_ws_bytes(int32_t dim) { (void)dim; return 256; }

extern "C" int xm3d_ravel_hash_f64(const double *coords, int64_t n, int32_t dim, uint64_t *keys, void *ws,
                                   size_t ws_bytes, xm3d_stream_t stream_) {
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    XM3D_REQUIRE(n >= 0 && dim > 0 && dim <= 8, "dim must be in 1..8");
    if (n == 0) return XM3D_OK;
    XM3D_REQUIRE(coords && keys && ws && ws_bytes >= 128, "null pointer / workspace");
    double *mm = static_cast<double *>(ws);
    init_minmax_kernel<<<1, 32, 0, stream>>>(mm);
    count_launches(1);
    unsigned blocks = (unsigned)((n + 255) / 256);
    if (blocks > (unsigned)sm_count() * 8) blocks = sm_count() * 8;
    colminmax_f64_kernel<<<blocks, 256, 0, stream>>>(coords, n, dim, mm);
    count_launches(1);
    ravel_f64_kernel<<<(unsigned)((n + 255) / 256), 256, 0, stream>>>(coords, n, dim, mm,
                                                                       reinterpret_cast<unsigned long long *>(keys));
    count_launches(1);
    return check_launch("xm3d_ravel_hash_f64");
}
