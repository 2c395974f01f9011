// Shared device helpers for the blades_b200 sm_100a kernels.
#pragma once
#include <cuda_runtime.h>
#include <cstdint>
#include <cfloat>

#define BL_MAX_ROWS 512      // max clients per aggregation (N <= 512, SURVEY config #5)
#define BL_MAX_PEERS 8       // GPUs on one NVSwitch node

// torch.nan_to_num semantics: NaN -> 0, +inf -> FLT_MAX, -inf -> -FLT_MAX (reference client.py:198)
__device__ __forceinline__ float bl_sanitize(float x) {
    if (x != x) return 0.0f;
    return fminf(fmaxf(x, -FLT_MAX), FLT_MAX);
}

// streaming (read-once) global load: bypass L1 allocation; works for local and peer-mapped memory.
__device__ __forceinline__ float bl_ldg_stream(const float* p) {
    float v;
    asm volatile("ld.global.nc.L1::no_allocate.f32 %0, [%1];" : "=f"(v) : "l"(p));
    return v;
}
__device__ __forceinline__ float4 bl_ldg_stream4(const float* p) {
    float4 v;
    asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];"
                 : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p));
    return v;
}

// Replicated output + fused server step:
//   out[g][c]   = agg                      for every replica g (peer pointers over NVLink)
//   theta[g][c] = theta_src[c] + lr * agg  (momentum-free SGD on every replica of theta)
struct BlEpilogue {
    float* out[BL_MAX_PEERS];
    float* theta[BL_MAX_PEERS];
    const float* theta_src;
    float lr;
    int n_out;
    int n_theta;
};

__device__ __forceinline__ void bl_epilogue_store(const BlEpilogue& ep, long long c, float agg) {
#pragma unroll
    for (int g = 0; g < BL_MAX_PEERS; ++g)
        if (g < ep.n_out) ep.out[g][c] = agg;
    if (ep.n_theta > 0) {
        float t = ep.theta_src[c] + ep.lr * agg;
#pragma unroll
        for (int g = 0; g < BL_MAX_PEERS; ++g)
            if (g < ep.n_theta) ep.theta[g][c] = t;
    }
}
