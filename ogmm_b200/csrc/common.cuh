// Shared helpers for the ogmm_b200 kernels (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <math.h>

#include "../../include/ogmm_b200.h"

namespace ogmm {

// ---- status / thread-local last error -----------------------------------------------------
void set_error(const char* fmt, ...);
int cuda_status(cudaError_t e, const char* what);

#define OGMM_REQUIRE(cond, code, ...)                 \
    do {                                              \
        if (!(cond)) {                                \
            ::ogmm::set_error(__VA_ARGS__);           \
            return (code);                            \
        }                                             \
    } while (0)

#define OGMM_LAUNCH_CHECK(what)                                              \
    do {                                                                     \
        int st__ = ::ogmm::cuda_status(cudaGetLastError(), what);            \
        if (st__ != OGMM_OK) return st__;                                    \
    } while (0)

static inline cudaStream_t as_stream(ogmm_stream_t s) { return reinterpret_cast<cudaStream_t>(s); }

constexpr int kWarp = 32;
constexpr unsigned kFull = 0xffffffffu;

// ---- warp / block reductions -----------------------------------------------------------------
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(kFull, v, o);
    return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(kFull, v, o));
    return v;
}
__device__ __forceinline__ unsigned long long warp_max_u64(unsigned long long v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        unsigned long long w = __shfl_xor_sync(kFull, v, o);
        v = w > v ? w : v;
    }
    return v;
}

// Block-wide sum of one float; `scratch` holds >= 32 floats.  All threads get the result.
template <int NT>
__device__ __forceinline__ float block_sum(float v, float* scratch) {
    constexpr int NW = NT / kWarp;
    v = warp_sum(v);
    __syncthreads();                       // protect scratch reuse
    if ((threadIdx.x & 31) == 0) scratch[threadIdx.x >> 5] = v;
    __syncthreads();
    float r = (threadIdx.x & 31) < NW ? scratch[threadIdx.x & 31] : 0.f;
    return warp_sum(r);
}

// Packed FP32 FMA (sm_100 FFMA2): d.x += a.x * b.x, d.y += a.y * b.y in one issue slot.
__device__ __forceinline__ void ffma2_pair(float2& d, const float2 a, const float2 b) {
    asm("fma.rn.f32x2 %0, %1, %2, %0;"
        : "+l"(*reinterpret_cast<unsigned long long*>(&d))
        : "l"(*reinterpret_cast<const unsigned long long*>(&a)), "l"(*reinterpret_cast<const unsigned long long*>(&b)));
}

constexpr int kJC = 16;                 // column chunk held in registers

// ---- 16-column butterfly: every lane enters with 16 partial sums, lane l leaves with the warp
// total of column (l >> 1) & 15 (lanes l and l^1 hold the same column).  16 shuffles.
__device__ __forceinline__ float butterfly16(float (&v)[kJC], int lane) {
    {
        const bool hi = lane & 16;
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            float send = hi ? v[i] : v[i + 8];
            float keep = hi ? v[i + 8] : v[i];
            v[i] = keep + __shfl_xor_sync(kFull, send, 16);
        }
    }
    {
        const bool hi = lane & 8;
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            float send = hi ? v[i] : v[i + 4];
            float keep = hi ? v[i + 4] : v[i];
            v[i] = keep + __shfl_xor_sync(kFull, send, 8);
        }
    }
    {
        const bool hi = lane & 4;
#pragma unroll
        for (int i = 0; i < 2; ++i) {
            float send = hi ? v[i] : v[i + 2];
            float keep = hi ? v[i + 2] : v[i];
            v[i] = keep + __shfl_xor_sync(kFull, send, 4);
        }
    }
    {
        const bool hi = lane & 2;
        float send = hi ? v[0] : v[1];
        float keep = hi ? v[1] : v[0];
        v[0] = keep + __shfl_xor_sync(kFull, send, 2);
    }
    return v[0] + __shfl_xor_sync(kFull, v[0], 1);
}

// Streaming (read-once) global loads: do not allocate in L1.
__device__ __forceinline__ float4 ldg_stream4(const float* p) {
    float4 r;
    asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];"
                 : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w) : "l"(p));
    return r;
}
__device__ __forceinline__ float ldg_stream(const float* p) {
    float r;
    asm volatile("ld.global.nc.L1::no_allocate.f32 %0, [%1];" : "=f"(r) : "l"(p));
    return r;
}

// torch.nan_to_num(x, nan=nan_value): NaN -> nan_value, +inf -> FLT_MAX, -inf -> -FLT_MAX.
__device__ __forceinline__ float nan_to_num(float x, float nan_value) {
    if (x != x) return nan_value;
    if (x == INFINITY) return 3.402823466e+38f;
    if (x == -INFINITY) return -3.402823466e+38f;
    return x;
}

}  // namespace ogmm
