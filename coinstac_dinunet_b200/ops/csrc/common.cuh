// Shared device helpers for the sm_100a kernels of coinstac_dinunet_b200.
// Everything here is plain CUDA C++ + inline PTX; no CUTLASS, no torch headers, so a .cu file
// compiles in seconds with:  nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <stdint.h>

#define B200_SM_COUNT 148
#define COINN_API extern "C" __attribute__((visibility("default")))

// error code convention of the C ABI: 0 == ok, otherwise cudaError_t of the failing call
#define COINN_CHECK_LAUNCH() do { cudaError_t e_ = cudaGetLastError(); if (e_ != cudaSuccess) return (int)e_; } while (0)

namespace coinn {

__device__ __forceinline__ uint32_t lane_id() { return threadIdx.x & 31u; }

// ---- 128-bit streaming loads / stores (bypass L1 allocation: touched exactly once) ----------
__device__ __forceinline__ float4 ld_stream_f4(const float4* p) {
    float4 r;
    asm volatile("ld.global.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];"
                 : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w) : "l"(p));
    return r;
}
__device__ __forceinline__ uint4 ld_stream_u4(const uint4* p) {
    uint4 r;
    asm volatile("ld.global.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];"
                 : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p));
    return r;
}
__device__ __forceinline__ uint2 ld_stream_u2(const uint2* p) {
    uint2 r;
    asm volatile("ld.global.L1::no_allocate.v2.u32 {%0,%1}, [%2];" : "=r"(r.x), "=r"(r.y) : "l"(p));
    return r;
}
__device__ __forceinline__ void st_stream_f4(float4* p, const float4& v) {
    asm volatile("st.global.L1::no_allocate.v4.f32 [%0], {%1,%2,%3,%4};"
                 :: "l"(p), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}
__device__ __forceinline__ void st_stream_u2(uint2* p, const uint2& v) {
    asm volatile("st.global.L1::no_allocate.v2.u32 [%0], {%1,%2};" :: "l"(p), "r"(v.x), "r"(v.y) : "memory");
}

// ---- system-scope flags (cross-GPU, over NVLink) -------------------------------------------
__device__ __forceinline__ void st_release_sys(uint32_t* p, uint32_t v) {
    asm volatile("st.release.sys.global.u32 [%0], %1;" :: "l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ uint32_t ld_acquire_sys(const uint32_t* p) {
    uint32_t v;
    asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}

// ---- NVLS (in-switch reduction / broadcast on a multicast address) --------------------------
__device__ __forceinline__ float4 multimem_ld_reduce_f4(const float4* mc) {
    float4 r;
    asm volatile("multimem.ld_reduce.relaxed.sys.global.add.v4.f32 {%0,%1,%2,%3}, [%4];"
                 : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w) : "l"(mc) : "memory");
    return r;
}
// 8 bf16 summed in the switch with fp32 accumulation
__device__ __forceinline__ uint4 multimem_ld_reduce_bf16x8(const uint4* mc) {
    uint4 r;
    asm volatile("multimem.ld_reduce.relaxed.sys.global.add.acc::f32.v4.bf16x2 {%0,%1,%2,%3}, [%4];"
                 : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(mc) : "memory");
    return r;
}
__device__ __forceinline__ void multimem_st_f4(float4* mc, const float4& v) {
    asm volatile("multimem.st.relaxed.sys.global.v4.f32 [%0], {%1,%2,%3,%4};"
                 :: "l"(mc), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}
__device__ __forceinline__ void multimem_st_u2(uint2* mc, const uint2& v) {
    // 64-bit multicast store expressed as v2.f32 (bit pattern preserved)
    asm volatile("multimem.st.relaxed.sys.global.v2.f32 [%0], {%1,%2};"
                 :: "l"(mc), "f"(__uint_as_float(v.x)), "f"(__uint_as_float(v.y)) : "memory");
}

// ---- bf16 packing -----------------------------------------------------------------------------
__device__ __forceinline__ uint32_t pack_bf16x2(float lo, float hi) {
    __nv_bfloat162 t = __floats2bfloat162_rn(lo, hi);
    return *reinterpret_cast<uint32_t*>(&t);
}
__device__ __forceinline__ float2 unpack_bf16x2(uint32_t u) {
    __nv_bfloat162 t = *reinterpret_cast<__nv_bfloat162*>(&u);
    return __bfloat1622float2(t);
}
__device__ __forceinline__ float2 unpack_f16x2(uint32_t u) {
    __half2 t = *reinterpret_cast<__half2*>(&u);
    return __half22float2(t);
}

// ---- warp / block reductions --------------------------------------------------------------------
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}
// block-wide sum; `scratch` needs >= 32 floats of shared memory; result valid in all threads
__device__ __forceinline__ float block_sum(float v, float* scratch) {
    v = warp_sum(v);
    __syncthreads();
    if (lane_id() == 0) scratch[threadIdx.x >> 5] = v;
    __syncthreads();
    float t = (threadIdx.x < (blockDim.x + 31) / 32) ? scratch[threadIdx.x] : 0.f;
    if (threadIdx.x < 32) { t = warp_sum(t); if (threadIdx.x == 0) scratch[0] = t; }
    __syncthreads();
    return scratch[0];
}

}  // namespace coinn
