// loss_contra mask selection on the device (models/utils/criterion.py:80-146), for every scene of a batch at once.
//
// The reference walks the K masks of every scene in Python with four `.item()` host syncs per mask:
//   mask_3d = sigmoid(mask[:, x, y]) >= 0.5; if no mask has >= 10 points: mask_3d[0, :] = True      (:83-88)
//   keep = sum(mask_3d, 1) >= 10                                                                     (:90)
//   per kept mask: novel_num = #(binary_gt == 0), base_num = len - novel_num, base_num_ = #(binary_gt == 1),
//                  novel_num_ = len - base_num_;  novel candidate iff novel_num > 1.8 base_num and novel_num > 10,
//                  else base candidate iff base_num_ > 20 novel_num_ and base_num_ > 150;
//                  score = mean(sigmoid(mask)[sigmoid(mask) > 0.5]) over the whole image                (:100-121)
//   pooled masks = top-4 novel + top-1 base by descending score (stable)                               (:124-139)
// Here: one counting pass over the visible points (warp ballots, shared-memory histograms, integer atomics), one
// image reduction per CANDIDATE mask (non-candidates return at once), one warp per scene for the ranking, and one
// pass that compacts the selected masks into a 5-bit membership word per point — the input of xm3d_pool_batch
// (k = 5).  No host round trip; all integer results are exact, the score is a float64 sum of float32 sigmoids.
#include "common.cuh"
#include "vec.cuh"

namespace xm3d {

constexpr int CS_THREADS = 1024;
constexpr int CS_MAXK = 32 * MAX_WORDS;

__device__ __forceinline__ float sigmoid_f32(float x) { return __fdiv_rn(1.0f, 1.0f + expf(-x)); }

// counts[s][m] = {points of mask m, of which binary_gt == 0, of which binary_gt == 1}; totals[s] = the same over all
// points of scene s (what mask 0 holds once the guard sets it to all-True)
__global__ void __launch_bounds__(CS_THREADS)
contra_count_kernel(const uint32_t *__restrict__ member, int words, int k, const float *__restrict__ binary_gt,
                    const int64_t *__restrict__ seg_off, int n_seg, int64_t cap, int *__restrict__ counts,
                    int *__restrict__ totals) {
    __shared__ int s_hist[CS_MAXK * 3 + 3];
    int64_t total = seg_off[n_seg];
    if (total > cap) total = 0;
    const int64_t b0 = (int64_t)blockIdx.x * CS_THREADS;
    if (b0 >= total) return;
    const int64_t b1 = min(total, b0 + CS_THREADS);
    const int tid = threadIdx.x, lane = tid & 31;
    const int64_t i = b0 + tid;
    const int tail = k & 31;
    for (int j = tid; j < k * 3 + 3; j += CS_THREADS) s_hist[j] = 0;
    __syncthreads();
    for (int s = seg_of(seg_off, n_seg, b0); s < n_seg && seg_off[s] < b1; ++s) {
        const int64_t lo = max(seg_off[s], b0), hi = min(seg_off[s + 1], b1);
        if (lo >= hi) continue;
        const bool valid = i >= lo && i < hi;
        const float gt = valid ? __ldg(binary_gt + i) : -1.f;
        const bool g0 = valid && gt == 0.f, g1 = valid && gt == 1.f;
        const uint32_t bv = __ballot_sync(0xffffffffu, valid);
        if (bv) {                                                      // warp-uniform
            const uint32_t t0 = __ballot_sync(0xffffffffu, g0), t1 = __ballot_sync(0xffffffffu, g1);
            if (lane == 0) {
                atomicAdd(&s_hist[k * 3 + 0], __popc(bv));
                atomicAdd(&s_hist[k * 3 + 1], __popc(t0));
                atomicAdd(&s_hist[k * 3 + 2], __popc(t1));
            }
            for (int w = 0; w < words; ++w) {
                uint32_t bits = valid ? __ldg(member + (size_t)i * words + w) : 0u;
                if (tail && w == (k >> 5)) bits &= (1u << tail) - 1u;
                if (w * 32 >= k) bits = 0u;
                uint32_t any = bits;                                   // masks that occur in this warp at all
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) any |= __shfl_xor_sync(0xffffffffu, any, o);
                while (any) {
                    const int b = __ffs(any) - 1;
                    any &= any - 1;
                    const bool in = (bits >> b) & 1u;
                    const uint32_t c = __ballot_sync(0xffffffffu, in), c0 = __ballot_sync(0xffffffffu, in && g0),
                                   c1 = __ballot_sync(0xffffffffu, in && g1);
                    if (lane == 0) {
                        const int m = w * 32 + b;
                        atomicAdd(&s_hist[m * 3 + 0], __popc(c));
                        if (c0) atomicAdd(&s_hist[m * 3 + 1], __popc(c0));
                        if (c1) atomicAdd(&s_hist[m * 3 + 2], __popc(c1));
                    }
                }
            }
        }
        __syncthreads();
        for (int j = tid; j < k * 3 + 3; j += CS_THREADS) {
            const int v = s_hist[j];
            if (v) {
                if (j < k * 3) atomicAdd(&counts[(size_t)s * k * 3 + j], v);
                else atomicAdd(&totals[s * 4 + (j - k * 3)], v);
                s_hist[j] = 0;
            }
        }
        __syncthreads();
    }
}

// effective counts of mask m of scene s after the ">= 10 points else row 0 all True" guard
struct ContraCounts { int cnt, c0, c1; bool guard; };
__device__ __forceinline__ ContraCounts contra_counts(const int *__restrict__ counts, const int *__restrict__ totals,
                                                     int s, int m, int k, bool guard) {
    ContraCounts r;
    r.guard = guard;
    if (guard && m == 0) { r.cnt = totals[s * 4]; r.c0 = totals[s * 4 + 1]; r.c1 = totals[s * 4 + 2]; }
    else { const int *c = counts + ((size_t)s * k + m) * 3; r.cnt = c[0]; r.c0 = c[1]; r.c1 = c[2]; }
    return r;
}
__device__ __forceinline__ int contra_kind(const ContraCounts &c) {
    if (c.cnt < 10) return 0;                                           // not kept (:90)
    const int novel_num = c.c0, base_num = c.cnt - c.c0, base_num_ = c.c1, novel_num_ = c.cnt - c.c1;
    if ((double)novel_num > 1.8 * (double)base_num && novel_num > 10) return 1;
    if (base_num_ > 20 * novel_num_ && base_num_ > 150) return 2;
    return 0;
}

// one CTA per (mask, scene): classification, then for candidates the mean sigmoid over the pixels with sigmoid > 0.5
__global__ void __launch_bounds__(256)
contra_score_kernel(const float *__restrict__ mask_logits, int hw, int k, int n_seg, const int *__restrict__ counts,
                    int *__restrict__ totals, int8_t *__restrict__ kind, float *__restrict__ score,
                    int *__restrict__ eff_counts) {
    __shared__ int s_any;
    __shared__ double s_sum[8];
    __shared__ int s_n[8];
    const int m = blockIdx.x, s = blockIdx.y, tid = threadIdx.x;
    if (tid == 0) s_any = 0;
    __syncthreads();
    int mine = 0;
    for (int j = tid; j < k; j += 256) mine |= counts[((size_t)s * k + j) * 3] >= 10;
    if (mine) s_any = 1;
    __syncthreads();
    const bool guard = !s_any;
    if (m == 0 && tid == 0) totals[s * 4 + 3] = guard ? 1 : 0;          // read by contra_member_kernel
    const ContraCounts cc = contra_counts(counts, totals, s, m, k, guard);
    const int kd = contra_kind(cc);
    if (tid == 0) {
        kind[s * k + m] = (int8_t)kd;
        int *e = eff_counts + ((size_t)s * k + m) * 3;
        e[0] = cc.cnt; e[1] = cc.c0; e[2] = cc.c1;
        if (!kd) score[s * k + m] = __int_as_float(0x7fc00000);          // NaN: not a candidate
    }
    if (!kd) return;
    const float *p = mask_logits + ((size_t)s * k + m) * hw;
    double sum = 0.0;
    int n = 0;
    for (int j = tid; j < hw; j += 256) {
        const float sg = sigmoid_f32(__ldg(p + j));
        if (sg > 0.5f) { sum += (double)sg; ++n; }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        sum += __shfl_xor_sync(0xffffffffu, sum, o);
        n += __shfl_xor_sync(0xffffffffu, n, o);
    }
    if ((tid & 31) == 0) { s_sum[tid >> 5] = sum; s_n[tid >> 5] = n; }
    __syncthreads();
    if (tid == 0) {
        double t = 0.0;
        int c = 0;
        for (int w = 0; w < 8; ++w) { t += s_sum[w]; c += s_n[w]; }     // fixed order: deterministic
        score[s * k + m] = c > 0 ? (float)(t / (double)c) : __int_as_float(0x7fc00000);
    }
}

// one thread per scene: top-4 novel then top-1 base, descending score, stable (first of equal scores wins)
__global__ void contra_select_kernel(const int8_t *__restrict__ kind, const float *__restrict__ score, int k, int n_seg,
                                     int *__restrict__ sel, int *__restrict__ n_sel) {
    const int s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= n_seg) return;
    const int8_t *kd = kind + (size_t)s * k;
    const float *sc = score + (size_t)s * k;
    int out[5] = {-1, -1, -1, -1, -1};
    int cnt = 0;
    for (int want = 1; want <= 2; ++want) {
        const int take = want == 1 ? 4 : 1;
        for (int r = 0; r < take; ++r) {
            int best = -1;
            float bs = 0.f;
            for (int m = 0; m < k; ++m) {
                if (kd[m] != want) continue;
                bool used = false;
                for (int q = 0; q < cnt; ++q) used |= out[q] == m;
                if (used) continue;
                const float v = sc[m] == sc[m] ? sc[m] : -INFINITY;      // NaN scores rank last
                if (best < 0 || v > bs) { best = m; bs = v; }
            }
            if (best < 0) break;
            out[cnt++] = best;
        }
    }
    for (int q = 0; q < 5; ++q) sel[s * 5 + q] = out[q];
    n_sel[s] = cnt;
}

// bit j of sel_member[i] = point i lies in the j-th selected mask of its scene (guard: mask 0 = every point)
__global__ void __launch_bounds__(256)
contra_member_kernel(const uint32_t *__restrict__ member, int words, const int64_t *__restrict__ seg_off, int n_seg,
                     int64_t cap, const int *__restrict__ totals, const int *__restrict__ sel,
                     const int *__restrict__ n_sel, uint32_t *__restrict__ sel_member) {
    int64_t total = seg_off[n_seg];
    if (total > cap) total = 0;
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= total) return;
    const int s = seg_of(seg_off, n_seg, i);
    uint32_t out = 0u;
    const int ns = n_sel[s];
    for (int j = 0; j < ns; ++j) {
        const int m = sel[s * 5 + j];
        uint32_t bit = (__ldg(member + (size_t)i * words + (m >> 5)) >> (m & 31)) & 1u;
        if (m == 0 && totals[s * 4 + 3]) bit = 1u;                      // the guard made mask 0 all-True (:87-88)
        out |= bit << j;
    }
    sel_member[i] = out;
}

}  // namespace xm3d

using namespace xm3d;

extern "C" size_t xm3d_contra_ws_bytes(int32_t n_seg, int32_t k) {
    Carver cv(nullptr);
    cv.take<int>((size_t)n_seg * k * 3);
    cv.take<int>((size_t)n_seg * 4);
    return cv.off + 256;
}

extern "C" int xm3d_contra_select_batch(const uint32_t *member, int32_t k, const float *binary_gt,
                                        const int64_t *seg_off, int32_t n_seg, int64_t cap, const float *mask_logits,
                                        int32_t h, int32_t w, int32_t *counts, int8_t *kind, float *score, int32_t *sel,
                                        int32_t *n_sel, uint32_t *sel_member, void *ws, size_t ws_bytes,
                                        xm3d_stream_t stream_) {
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    XM3D_REQUIRE(n_seg > 0 && k > 0 && k <= CS_MAXK && cap >= 0 && h > 0 && w > 0, "bad sizes");
    XM3D_REQUIRE(member && binary_gt && seg_off && mask_logits && counts && kind && score && sel && n_sel && sel_member && ws,
                 "null pointer");
    if (ws_bytes < xm3d_contra_ws_bytes(n_seg, k)) {
        set_error("xm3d_contra_select_batch: workspace too small");
        return XM3D_ERR_WORKSPACE;
    }
    Carver cv(ws);
    int *raw = cv.take<int>((size_t)n_seg * k * 3);
    int *totals = cv.take<int>((size_t)n_seg * 4);
    cudaMemsetAsync(raw, 0, sizeof(int) * (size_t)n_seg * k * 3, stream);
    cudaMemsetAsync(totals, 0, sizeof(int) * (size_t)n_seg * 4, stream);
    const int words = words_for(k);
    if (cap > 0) {
        contra_count_kernel<<<(unsigned)((cap + CS_THREADS - 1) / CS_THREADS), CS_THREADS, 0, stream>>>(
            member, words, k, binary_gt, seg_off, n_seg, cap, raw, totals);
        count_launches(1);
    }
    contra_score_kernel<<<dim3((unsigned)k, (unsigned)n_seg), 256, 0, stream>>>(mask_logits, h * w, k, n_seg, raw, totals, kind,
                                                                               score, counts);
    count_launches(1);
    contra_select_kernel<<<(unsigned)((n_seg + 63) / 64), 64, 0, stream>>>(kind, score, k, n_seg, sel, n_sel);
    count_launches(1);
    if (cap > 0) {
        contra_member_kernel<<<(unsigned)((cap + 255) / 256), 256, 0, stream>>>(member, words, seg_off, n_seg, cap, totals, sel,
                                                                              n_sel, sel_member);
        count_launches(1);
    }
    return check_launch("xm3d_contra_select_batch");
}
