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

// coherent 16 B load (peer data written by another GPU's kernel before a device barrier: never the .nc path)
__device__ __forceinline__ float4 bl_ld_volatile4(const float* p) {
    float4 v;
    asm volatile("ld.volatile.global.v4.f32 {%0,%1,%2,%3}, [%4];"
                 : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p) : "memory");
    return v;
}

// Replicated output + fused server step:
//   out[g][c]   = agg                      for every replica g
//   theta[g][c] = theta_src[c] + lr * agg  (momentum-free SGD on every replica of theta)
// Two ways to reach the replicas: `mc_out` / `mc_theta` are NVLS multicast addresses of the symmetric allocation --
// ONE multimem.st per value, the NVSwitch replicates it into every GPU's copy (the reference's broadcast of the model
// to every actor, simulator.py:222-233, done by the switch); when the fabric has no multicast object (or on one GPU)
// they are null and the kernel issues one plain store per peer pointer.
struct BlEpilogue {
    float* out[BL_MAX_PEERS];
    float* theta[BL_MAX_PEERS];
    const float* theta_src;
    float lr;
    int n_out;
    int n_theta;
    float* mc_out;
    float* mc_theta;
};

__device__ __forceinline__ void bl_mc_store(float* p, float v) {
    asm volatile("multimem.st.relaxed.sys.global.f32 [%0], %1;" :: "l"(p), "f"(v) : "memory");
}
__device__ __forceinline__ void bl_mc_store4(float* p, float4 v) {
    asm volatile("multimem.st.relaxed.sys.global.v4.f32 [%0], {%1,%2,%3,%4};"
                 :: "l"(p), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}
// in-switch sum of the same address on every replica (NVLS reduction)
__device__ __forceinline__ float4 bl_mc_ld_reduce4(const float* p) {
    float4 v;
    asm volatile("multimem.ld_reduce.relaxed.sys.global.add.v4.f32 {%0,%1,%2,%3}, [%4];"
                 : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p) : "memory");
    return v;
}

__device__ __forceinline__ void bl_epilogue_store(const BlEpilogue& ep, long long c, float agg) {
    if (ep.mc_out) {
        bl_mc_store(ep.mc_out + c, agg);
    } else {
#pragma unroll
        for (int g = 0; g < BL_MAX_PEERS; ++g)
            if (g < ep.n_out) ep.out[g][c] = agg;
    }
    if (ep.n_theta > 0) {
        const float t = ep.theta_src[c] + ep.lr * agg;
        if (ep.mc_theta) {
            bl_mc_store(ep.mc_theta + c, t);
        } else {
#pragma unroll
            for (int g = 0; g < BL_MAX_PEERS; ++g)
                if (g < ep.n_theta) ep.theta[g][c] = t;
        }
    }
}

// four consecutive coordinates (c % 4 == 0, 16 B aligned replicas): 16 B stores / one v4 multicast store
__device__ __forceinline__ void bl_epilogue_store4(const BlEpilogue& ep, long long c, float4 agg) {
    if (ep.mc_out) {
        bl_mc_store4(ep.mc_out + c, agg);
    } else {
#pragma unroll
        for (int g = 0; g < BL_MAX_PEERS; ++g)
            if (g < ep.n_out) *reinterpret_cast<float4*>(ep.out[g] + c) = agg;
    }
    if (ep.n_theta > 0) {
        const float4 s = *reinterpret_cast<const float4*>(ep.theta_src + c);
        const float4 t = make_float4(s.x + ep.lr * agg.x, s.y + ep.lr * agg.y, s.z + ep.lr * agg.z, s.w + ep.lr * agg.w);
        if (ep.mc_theta) {
            bl_mc_store4(ep.mc_theta + c, t);
        } else {
#pragma unroll
            for (int g = 0; g < BL_MAX_PEERS; ++g)
                if (g < ep.n_theta) *reinterpret_cast<float4*>(ep.theta[g] + c) = t;
        }
    }
}
