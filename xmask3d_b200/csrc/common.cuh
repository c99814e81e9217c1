// Shared helpers of libxm3d (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include <atomic>

#include "../../include/xm3d.h"

#if defined(__CUDA_ARCH__) && (__CUDA_ARCH__ < 1000)
#error "libxm3d is written for sm_100a (B200) only"
#endif

namespace xm3d {

void set_error(const char *fmt, ...);
int check_launch(const char *what);  // cudaGetLastError -> XM3D_OK / XM3D_ERR_CUDA
void count_launches(int n);          // bookkeeping for xm3d_launch_count()
int sm_count();                      // of the CURRENT device (cached per device)
// One-time per-device setup (cudaFuncSetAttribute and the like): true exactly once per (flag, current device),
// thread-safe.  Kernel attributes are per device, so a process-wide `static bool` would be wrong for a process
// that drives two GPUs.
bool first_use_on_device(std::atomic<uint64_t> *flag);
extern thread_local cudaEvent_t g_pool_ev[2];   // optional timing hook (xm3d_set_pool_events)

#define XM3D_REQUIRE(cond, msg)                      \
    do {                                             \
        if (!(cond)) {                               \
            xm3d::set_error("%s: %s", __func__, msg); \
            return XM3D_ERR_BAD_ARG;                 \
        }                                            \
    } while (0)

static inline size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }

// Bump allocator over the caller's workspace; every block is 256-byte aligned.
struct Carver {
    char *base;
    size_t off;
    explicit Carver(void *p) : base(static_cast<char *>(p)), off(0) {}
    template <typename T>
    T *take(size_t count) {
        T *p = base ? reinterpret_cast<T *>(base + off) : nullptr;
        off += align_up(count * sizeof(T), 256);
        return p;
    }
};

#ifdef __CUDACC__
__device__ __forceinline__ int lane_id() { return threadIdx.x & 31; }

// Largest s with off[s] <= i  (off is an ascending int64 [n+1] array, off[0] == 0).
__device__ __forceinline__ int seg_of(const int64_t *__restrict__ off, int n, int64_t i) {
    int lo = 0, hi = n;  // invariant: off[lo] <= i < off[hi]
    while (hi - lo > 1) {
        int mid = (lo + hi) >> 1;
        if (off[mid] <= i) lo = mid; else hi = mid;
    }
    return lo;
}

// Read-once stream: no L1 allocation and an L2 evict-first policy, so that a multi-GB stream does
// not push the L2-resident working sets of concurrently running kernels (hash tables, depth
// images, membership words) out to DRAM.
__device__ __forceinline__ unsigned long long l2_evict_first_policy() {
    unsigned long long pol;
    asm("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
    return pol;
}
__device__ __forceinline__ float4 ldg_stream4(const float *p) {
    float4 r;
    asm volatile("ld.global.nc.L1::no_allocate.L2::cache_hint.v4.f32 {%0,%1,%2,%3}, [%4], %5;"
                 : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w) : "l"(p), "l"(l2_evict_first_policy()));
    return r;
}

// ---- mbarrier + 1-D bulk async copy (TMA engine, SASS: UBLKCP) ----------------------------
__device__ __forceinline__ uint32_t smem_u32(const void *p) {
    return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}
__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_fence_init() {
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes)
                 : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t *bar, uint32_t phase) {
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok) : "r"(smem_u32(bar)), "r"(phase) : "memory");
    return ok != 0;
}
// Bounded wait: a wrong transaction count must trap, never hang the GPU box.
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t phase) {
    for (uint32_t spin = 0; !mbar_try_wait(bar, phase); ++spin)
        if (spin > (1u << 26)) __trap();
}
// The same with a suspend-time hint: the waiting thread is parked by the hardware until the phase completes (or the
// hint expires) instead of polling — in warp-specialised kernels the polling loops of the idle roles otherwise take
// issue slots from the working warps (ncu on pool_mma2_kernel: 1.3 G warp instructions, half of them polls).
__device__ __forceinline__ void mbar_wait_parked(uint64_t *bar, uint32_t phase) {
    for (uint32_t spin = 0;; ++spin) {
        uint32_t ok;
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(ok) : "r"(smem_u32(bar)), "r"(phase), "r"(20000u) : "memory");     // up to 20 us per try
        if (ok) return;
        if (spin > (1u << 18)) __trap();                     // ~5 s: a wrong transaction count must trap, never hang
    }
}
// Wait of a role with slack (epilogue, TMA producer): sleep between polls so that the polling loop does not take
// issue slots from the working warps of the same SM sub-partition.
__device__ __forceinline__ void mbar_wait_sleep(uint64_t *bar, uint32_t phase, uint32_t ns) {
    for (uint32_t spin = 0; !mbar_try_wait(bar, phase); ++spin) {
        __nanosleep(ns);
        if (spin > (1u << 24)) __trap();
    }
}
// 32 x 32 bit-matrix transpose across the lanes of a warp: afterwards bit p of lane i = bit i of lane p before.
__device__ __forceinline__ uint32_t warp_transpose32(uint32_t x) {
    const int lane = threadIdx.x & 31;
#pragma unroll
    for (int j = 16; j >= 1; j >>= 1) {
        const uint32_t m = j == 16 ? 0x0000ffffu : j == 8 ? 0x00ff00ffu : j == 4 ? 0x0f0f0f0fu : j == 2 ? 0x33333333u : 0x55555555u;
        const uint32_t y = __shfl_xor_sync(0xffffffffu, x, j);
        x = (lane & j) ? (((y >> j) & m) | (x & ~m)) : ((x & m) | ((y & m) << j));
    }
    return x;
}
// N independent transposes, round by round: the N shuffles of a round issue back to back, so the dependent chain of
// one transpose (5 x (shuffle latency + 6 ALU) ~ 300 cycles) is shared by all of them.
template <int N>
__device__ __forceinline__ void warp_transpose32_n(uint32_t (&x)[N]) {
    const int lane = threadIdx.x & 31;
#pragma unroll
    for (int j = 16; j >= 1; j >>= 1) {
        const uint32_t m = j == 16 ? 0x0000ffffu : j == 8 ? 0x00ff00ffu : j == 4 ? 0x0f0f0f0fu : j == 2 ? 0x33333333u : 0x55555555u;
        uint32_t y[N];
#pragma unroll
        for (int i = 0; i < N; ++i) y[i] = __shfl_xor_sync(0xffffffffu, x[i], j);
#pragma unroll
        for (int i = 0; i < N; ++i) x[i] = (lane & j) ? (((y[i] >> j) & m) | (x[i] & ~m)) : ((x[i] & m) | ((y[i] & m) << j));
    }
}
// global -> shared bulk copy; bytes % 16 == 0, both addresses 16-byte aligned.
__device__ __forceinline__ void bulk_g2s(void *dst_smem, const void *src_gmem, uint32_t bytes, uint64_t *bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(smem_u32(dst_smem)), "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
#endif  // __CUDACC__

}  // namespace xm3d
