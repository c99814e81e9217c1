// float / float2 / float4 helpers shared by the stage-3 kernels.
#pragma once
#include "common.cuh"

namespace xm3d {

constexpr int MAX_WORDS = 8;                  // k <= 256 masks per segment

template <int VEC> struct VecT;
template <> struct VecT<4> { using type = float4; };
template <> struct VecT<2> { using type = float2; };
template <> struct VecT<1> { using type = float; };

// streaming (read-once) loads: non-coherent path, no L1 allocation
template <int VEC>
__device__ __forceinline__ typename VecT<VEC>::type ld_stream(const float *p);
template <>
__device__ __forceinline__ float4 ld_stream<4>(const float *p) { return ldg_stream4(p); }
template <>
__device__ __forceinline__ float2 ld_stream<2>(const float *p) {
    float2 r;
    asm volatile("ld.global.nc.L1::no_allocate.v2.f32 {%0,%1}, [%2];" : "=f"(r.x), "=f"(r.y) : "l"(p));
    return r;
}
template <>
__device__ __forceinline__ float ld_stream<1>(const float *p) {
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

static inline int words_for(int k) { return (k + 31) / 32; }

}  // namespace xm3d
