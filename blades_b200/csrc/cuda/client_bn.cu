// Per-client ("ghost") BatchNorm2d for client-batched training -- K9 of SURVEY 2.7.
// x is the concatenated batch [n_clients*B, C, H, W] (NCHW); statistics are taken per (client, channel)
// over that client's B*H*W values, exactly what each client would compute alone
// (reference client.py:178-193 with torchvision-style BatchNorm).  One fused kernel per direction:
//   fwd: mean/var (two passes, second from L1/L2) -> y = (x-mean)*rstd*gamma + beta; saves mean, rstd
//   bwd: dbeta_c, dgamma_c per client -> written (scaled by alpha = -lr) straight into the client's
//        row of the update matrix; dx = gamma*rstd*(gy - (dbeta + xhat*dgamma)/m)
// Mapping: a block owns (client, 256-float tile of the [C*HW] plane); thread t owns the same plane
// offset in every sample, so all 256 threads read 1 KB contiguous per sample and every thread's
// elements belong to one channel (requires HW to be a power of two <= 256; other shapes use the
// PyTorch composite path).
#include "common.cuh"

struct ClientBNParams {
    const float* x;      // [n*B, C, HW]
    const float* gy;     // bwd only
    float* y;            // fwd: output; bwd: dx
    const float* gamma;  // [C]
    const float* beta;   // [C]
    float* mean;         // [n, C]
    float* rstd;         // [n, C]
    float* dgamma;       // bwd: &U[0][off_gamma], row stride = ld
    float* dbeta;        // bwd: &U[0][off_beta]
    long long ld;
    int n, B, C, HW;
    float eps, alpha;
    // fused activation (NHWC kernels only):
    //   fwd: y = relu?(bn(x) + res)          res optional (residual branch), relu flag
    //   bwd: g = relu ? gy * (act > 0) : gy  act = the forward OUTPUT; gmask (optional, may alias gy) receives g --
    //        the gradient the residual branch needs
    const float* res;
    const float* act;
    float* gmask;
    int relu;
    int pad_;
};

__device__ __forceinline__ float4 bn_relu_mask(float4 g, float4 a) {
    return make_float4(a.x > 0.f ? g.x : 0.f, a.y > 0.f ? g.y : 0.f, a.z > 0.f ? g.z : 0.f, a.w > 0.f ? g.w : 0.f);
}
// y_pre = x * (gamma * rstd) + (beta - mean * gamma * rstd): ONE definition shared by the forward kernels and by the
// backward kernels that recompute the ReLU mask from x instead of reading the forward output (same expressions ->
// same contraction -> bit-identical y_pre, so (y_pre > 0) is exactly (act > 0) whenever no residual was added).
__device__ __forceinline__ void bn_affine(const float4 ga, const float4 be, const float4 mean, const float4 rstd,
                                          float4& g, float4& sh) {
    g = make_float4(ga.x * rstd.x, ga.y * rstd.y, ga.z * rstd.z, ga.w * rstd.w);
    sh = make_float4(be.x - mean.x * g.x, be.y - mean.y * g.y, be.z - mean.z * g.z, be.w - mean.w * g.w);
}
__device__ __forceinline__ float4 bn_pre(const float4 v, const float4 g, const float4 sh) {
    return make_float4(fmaf(v.x, g.x, sh.x), fmaf(v.y, g.y, sh.y), fmaf(v.z, g.z, sh.z), fmaf(v.w, g.w, sh.w));
}
__device__ __forceinline__ float4 bn_act(float4 o, const float4* res, long long idx, int relu) {
    if (res != nullptr) { const float4 r = res[idx]; o.x += r.x; o.y += r.y; o.z += r.z; o.w += r.w; }
    if (relu) { o.x = fmaxf(o.x, 0.f); o.y = fmaxf(o.y, 0.f); o.z = fmaxf(o.z, 0.f); o.w = fmaxf(o.w, 0.f); }
    return o;
}

__device__ __forceinline__ float channel_reduce(float v, int HW, float* smem) {
    // sum over the HW threads that own the same channel (HW power of two, segments aligned)
    const int span = HW < 32 ? HW : 32;
    for (int o = span >> 1; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    if (HW <= 32) return v;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    __syncthreads();
    if (lane == 0) smem[warp] = v;
    __syncthreads();
    const int wpc = HW >> 5;                       // warps per channel
    const int w0 = (warp / wpc) * wpc;
    float s = 0.f;
    for (int i = 0; i < wpc; ++i) s += smem[w0 + i];
    return s;
}

__global__ void __launch_bounds__(256)
client_bn_fwd_kernel(const __grid_constant__ ClientBNParams p) {
    __shared__ float red[8];
    const int c = blockIdx.y;
    const int plane = p.C * p.HW;
    const int e = blockIdx.x * 256 + threadIdx.x;
    const bool live = e < plane;
    const int ch = live ? e / p.HW : 0;
    const float* xb = p.x + ((long long)c * p.B) * plane + e;
    const float m = (float)(p.B * p.HW);
    float s = 0.f;
    if (live)
        for (int b = 0; b < p.B; ++b) s += xb[(long long)b * plane];
    const float mean = channel_reduce(s, p.HW, red) / m;
    float q = 0.f;
    if (live)
        for (int b = 0; b < p.B; ++b) { const float d = xb[(long long)b * plane] - mean; q = fmaf(d, d, q); }
    const float var = channel_reduce(q, p.HW, red) / m;
    const float rstd = rsqrtf(var + p.eps);
    if (live) {
        const float g = p.gamma[ch] * rstd, sh = p.beta[ch] - mean * g;
        float* yb = p.y + ((long long)c * p.B) * plane + e;
        for (int b = 0; b < p.B; ++b) yb[(long long)b * plane] = fmaf(xb[(long long)b * plane], g, sh);
        if (e % p.HW == 0) { p.mean[c * p.C + ch] = mean; p.rstd[c * p.C + ch] = rstd; }
    }
}

__global__ void __launch_bounds__(256)
client_bn_bwd_kernel(const __grid_constant__ ClientBNParams p) {
    __shared__ float red[8];
    const int c = blockIdx.y;
    const int plane = p.C * p.HW;
    const int e = blockIdx.x * 256 + threadIdx.x;
    const bool live = e < plane;
    const int ch = live ? e / p.HW : 0;
    const long long base = ((long long)c * p.B) * plane + e;
    const float mean = live ? p.mean[c * p.C + ch] : 0.f;
    const float rstd = live ? p.rstd[c * p.C + ch] : 0.f;
    float sb = 0.f, sg = 0.f;
    if (live)
        for (int b = 0; b < p.B; ++b) {
            const float g = p.gy[base + (long long)b * plane];
            const float xh = (p.x[base + (long long)b * plane] - mean) * rstd;
            sb += g;
            sg = fmaf(g, xh, sg);
        }
    const float dbeta = channel_reduce(sb, p.HW, red);
    const float dgamma = channel_reduce(sg, p.HW, red);
    if (live) {
        if (e % p.HW == 0) {
            p.dgamma[(long long)c * p.ld + ch] = bl_sanitize(p.alpha * dgamma);
            p.dbeta[(long long)c * p.ld + ch] = bl_sanitize(p.alpha * dbeta);
        }
        if (p.y != nullptr) {
            const float inv_m = 1.f / (float)(p.B * p.HW);
            const float k = p.gamma[ch] * rstd;
            for (int b = 0; b < p.B; ++b) {
                const float g = p.gy[base + (long long)b * plane];
                const float xh = (p.x[base + (long long)b * plane] - mean) * rstd;
                p.y[base + (long long)b * plane] = k * (g - (dbeta + xh * dgamma) * inv_m);
            }
        }
    }
}

static bool bn_shape_ok(int HW) { return HW >= 1 && HW <= 256 && (HW & (HW - 1)) == 0; }

extern "C" int bl_client_bn_fwd(const ClientBNParams* p, void* stream) {
    if (!bn_shape_ok(p->HW)) return -1;
    dim3 grid((p->C * p->HW + 255) / 256, p->n);
    client_bn_fwd_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(*p);
    return (int)cudaGetLastError();
}
extern "C" int bl_client_bn_bwd(const ClientBNParams* p, void* stream) {
    if (!bn_shape_ok(p->HW)) return -1;
    dim3 grid((p->C * p->HW + 255) / 256, p->n);
    client_bn_bwd_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(*p);
    return (int)cudaGetLastError();
}
extern "C" int bl_sizeof_bn_params() { return (int)sizeof(ClientBNParams); }

// ---------------------------------------------------------------------------------------------
// NHWC (channels_last) variants: x is physically [n*B*HW rows][C].  A block owns (client, 32-channel tile);
// thread = (row group 0..31, channel quad 0..7): every row segment is one coalesced 128 B read made of
// eight float4 loads, and each thread keeps 4 independent rows in flight (unrolled) -- the first version
// (one float per thread per row, no unrolling) was latency bound at ~1/3 of the NCHW kernel's speed.
// The block strides over the client's R = B*HW rows three times (sum, squared deviation, normalise); passes
// two and three are served by L2.  C % 4 == 0 uses the vector path, anything else a scalar twin.
constexpr int kBnTile = 32, kBnQuads = 8, kBnGroups = 32;

__device__ __forceinline__ float4 bn_group_reduce4(float4 v, float4 (*red)[kBnQuads]) {
    const int q = threadIdx.x % kBnQuads, rg = threadIdx.x / kBnQuads;
    __syncthreads();
    red[rg][q] = v;
    __syncthreads();
    float4 s = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll 8
    for (int i = 0; i < kBnGroups; ++i) {
        const float4 t = red[i][q];
        s.x += t.x; s.y += t.y; s.z += t.z; s.w += t.w;
    }
    return s;
}

__global__ void __launch_bounds__(256)
client_bn_nhwc_fwd_kernel(const __grid_constant__ ClientBNParams p) {
    __shared__ float4 red[kBnGroups][kBnQuads];
    const int c = blockIdx.y;
    const int q = threadIdx.x % kBnQuads, rg = threadIdx.x / kBnQuads;
    const int ch = blockIdx.x * kBnTile + q * 4;
    const bool live = ch < p.C;
    const int R = p.B * p.HW;
    const float4* xb = reinterpret_cast<const float4*>(p.x + (long long)c * R * p.C + ch);
    const long long rs = p.C / 4;                     // row stride in float4
    const float m = (float)R;
    float4 s = make_float4(0.f, 0.f, 0.f, 0.f);
    if (live) {
#pragma unroll 4
        for (int r = rg; r < R; r += kBnGroups) {
            const float4 v = xb[(long long)r * rs];
            s.x += v.x; s.y += v.y; s.z += v.z; s.w += v.w;
        }
    }
    float4 mean = bn_group_reduce4(s, red);
    mean.x /= m; mean.y /= m; mean.z /= m; mean.w /= m;
    float4 qv = make_float4(0.f, 0.f, 0.f, 0.f);
    if (live) {
#pragma unroll 4
        for (int r = rg; r < R; r += kBnGroups) {
            const float4 v = xb[(long long)r * rs];
            float d;
            d = v.x - mean.x; qv.x = fmaf(d, d, qv.x);
            d = v.y - mean.y; qv.y = fmaf(d, d, qv.y);
            d = v.z - mean.z; qv.z = fmaf(d, d, qv.z);
            d = v.w - mean.w; qv.w = fmaf(d, d, qv.w);
        }
    }
    const float4 var = bn_group_reduce4(qv, red);
    if (live) {
        const float4 rstd = make_float4(rsqrtf(var.x / m + p.eps), rsqrtf(var.y / m + p.eps),
                                        rsqrtf(var.z / m + p.eps), rsqrtf(var.w / m + p.eps));
        const float4 ga = *reinterpret_cast<const float4*>(p.gamma + ch);
        const float4 be = *reinterpret_cast<const float4*>(p.beta + ch);
        float4 g, sh;
        bn_affine(ga, be, mean, rstd, g, sh);
        float4* yb = reinterpret_cast<float4*>(p.y + (long long)c * R * p.C + ch);
        const float4* rb = p.res ? reinterpret_cast<const float4*>(p.res + (long long)c * R * p.C + ch) : nullptr;
#pragma unroll 4
        for (int r = rg; r < R; r += kBnGroups) {
            const float4 v = xb[(long long)r * rs];
            yb[(long long)r * rs] = bn_act(bn_pre(v, g, sh), rb, (long long)r * rs, p.relu);
        }
        if (rg == 0) {
            *reinterpret_cast<float4*>(p.mean + c * p.C + ch) = mean;
            *reinterpret_cast<float4*>(p.rstd + c * p.C + ch) = rstd;
        }
    }
}

__global__ void __launch_bounds__(256)
client_bn_nhwc_bwd_kernel(const __grid_constant__ ClientBNParams p) {
    __shared__ float4 red[kBnGroups][kBnQuads];
    const int c = blockIdx.y;
    const int q = threadIdx.x % kBnQuads, rg = threadIdx.x / kBnQuads;
    const int ch = blockIdx.x * kBnTile + q * 4;
    const bool live = ch < p.C;
    const int R = p.B * p.HW;
    const long long base = (long long)c * R * p.C + ch;
    const float4* xb = reinterpret_cast<const float4*>(p.x + base);
    const float4* gb = reinterpret_cast<const float4*>(p.gy + base);
    const float4* ab = reinterpret_cast<const float4*>(p.act + base);       // only dereferenced when p.relu
    float4* mb = p.gmask ? reinterpret_cast<float4*>(p.gmask + base) : nullptr;
    const long long rs = p.C / 4;
    float4 mean = make_float4(0.f, 0.f, 0.f, 0.f), rstd = mean;
    if (live) {
        mean = *reinterpret_cast<const float4*>(p.mean + c * p.C + ch);
        rstd = *reinterpret_cast<const float4*>(p.rstd + c * p.C + ch);
    }
    // ReLU mask: from the forward output `act`, or -- act == nullptr, unit without a residual -- recomputed from x
    const bool remask = p.relu && p.act == nullptr;
    float4 fg = make_float4(0.f, 0.f, 0.f, 0.f), fsh = fg;
    if (live && remask)
        bn_affine(*reinterpret_cast<const float4*>(p.gamma + ch), *reinterpret_cast<const float4*>(p.beta + ch), mean, rstd,
                  fg, fsh);
    float4 sb = make_float4(0.f, 0.f, 0.f, 0.f), sg = sb;
    if (live) {
#pragma unroll 4
        for (int r = rg; r < R; r += kBnGroups) {
            float4 g = gb[(long long)r * rs];
            const float4 v = xb[(long long)r * rs];
            if (p.relu) g = bn_relu_mask(g, remask ? bn_pre(v, fg, fsh) : ab[(long long)r * rs]);
            if (mb != nullptr) mb[(long long)r * rs] = g;
            sb.x += g.x; sb.y += g.y; sb.z += g.z; sb.w += g.w;
            sg.x = fmaf(g.x, (v.x - mean.x) * rstd.x, sg.x);
            sg.y = fmaf(g.y, (v.y - mean.y) * rstd.y, sg.y);
            sg.z = fmaf(g.z, (v.z - mean.z) * rstd.z, sg.z);
            sg.w = fmaf(g.w, (v.w - mean.w) * rstd.w, sg.w);
        }
    }
    const float4 dbeta = bn_group_reduce4(sb, red);
    const float4 dgamma = bn_group_reduce4(sg, red);
    if (live) {
        if (rg == 0) {
            float* dg = p.dgamma + (long long)c * p.ld + ch;
            float* db = p.dbeta + (long long)c * p.ld + ch;
            dg[0] = bl_sanitize(p.alpha * dgamma.x); dg[1] = bl_sanitize(p.alpha * dgamma.y);
            dg[2] = bl_sanitize(p.alpha * dgamma.z); dg[3] = bl_sanitize(p.alpha * dgamma.w);
            db[0] = bl_sanitize(p.alpha * dbeta.x); db[1] = bl_sanitize(p.alpha * dbeta.y);
            db[2] = bl_sanitize(p.alpha * dbeta.z); db[3] = bl_sanitize(p.alpha * dbeta.w);
        }
        if (p.y != nullptr) {
            const float inv_m = 1.f / (float)R;
            const float4 ga = *reinterpret_cast<const float4*>(p.gamma + ch);
            const float4 k = make_float4(ga.x * rstd.x, ga.y * rstd.y, ga.z * rstd.z, ga.w * rstd.w);
            float4* yb = reinterpret_cast<float4*>(p.y + base);
#pragma unroll 4
            for (int r = rg; r < R; r += kBnGroups) {
                float4 g = gb[(long long)r * rs];
                const float4 v = xb[(long long)r * rs];
                if (p.relu) g = bn_relu_mask(g, remask ? bn_pre(v, fg, fsh) : ab[(long long)r * rs]);   // idempotent if gmask aliased gy
                float4 o;
                o.x = k.x * (g.x - (dbeta.x + (v.x - mean.x) * rstd.x * dgamma.x) * inv_m);
                o.y = k.y * (g.y - (dbeta.y + (v.y - mean.y) * rstd.y * dgamma.y) * inv_m);
                o.z = k.z * (g.z - (dbeta.z + (v.z - mean.z) * rstd.z * dgamma.z) * inv_m);
                o.w = k.w * (g.w - (dbeta.w + (v.w - mean.w) * rstd.w * dgamma.w) * inv_m);
                yb[(long long)r * rs] = o;
            }
        }
    }
}

// ---------------------------------------------------------------------------------------------
// Cluster variants (default).  The plain NHWC kernels above launch only ceil(C/32) * n CTAs (200 for the stem BN of
// ResNet-18 at 100 clients: 1.35 CTAs per SM) and stream x from DRAM up to three times, because the ~2 MB
// per-client slices of all resident CTAs together do not fit L2.  Here a thread-block CLUSTER of S CTAs owns one
// (client, channel tile): CTA r stages rows [r*R/S, (r+1)*R/S) of the tile in shared memory while it accumulates
// its partial sums, the S partials are exchanged through distributed shared memory (one float4 per channel quad),
// and every later pass (exact two-pass variance, normalisation; dx in the backward) runs out of shared memory.
// x (and gy) cross HBM exactly once, y/dx once; the grid is S times larger.
#include <cooperative_groups.h>
namespace cg = cooperative_groups;

template <int QUADS>
__device__ __forceinline__ float4 bn_cl_reduce4(float4 v, float4* red) {      // red: [256/QUADS][QUADS]
    constexpr int GROUPS = 256 / QUADS;
    const int q = threadIdx.x % QUADS, rg = threadIdx.x / QUADS;
    __syncthreads();
    red[rg * QUADS + q] = v;
    __syncthreads();
    float4 s = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll 8
    for (int i = 0; i < GROUPS; ++i) {
        const float4 t = red[i * QUADS + q];
        s.x += t.x; s.y += t.y; s.z += t.z; s.w += t.w;
    }
    return s;
}

template <int QUADS>
__device__ __forceinline__ float4 bn_cl_exchange(cg::cluster_group& cluster, float4* part, float4 mine, int slot) {
    // publish this CTA's partial, then sum the partials of every CTA of the cluster through DSMEM
    const int q = threadIdx.x % QUADS, rg = threadIdx.x / QUADS;
    if (rg == 0) part[slot * QUADS + q] = mine;
    cluster.sync();
    float4 s = make_float4(0.f, 0.f, 0.f, 0.f);
    const unsigned S = cluster.num_blocks();
    for (unsigned j = 0; j < S; ++j) {
        const float4 t = cluster.map_shared_rank(part, j)[slot * QUADS + q];
        s.x += t.x; s.y += t.y; s.z += t.z; s.w += t.w;
    }
    return s;
}

template <int QUADS>
__global__ void __launch_bounds__(256)
client_bn_nhwc_fwd_cl_kernel(const __grid_constant__ ClientBNParams p) {
    constexpr int GROUPS = 256 / QUADS;
    extern __shared__ __align__(16) float4 bn_tile[];          // [Rc][QUADS]
    __shared__ float4 red[256];
    __shared__ float4 part[2 * QUADS];
    cg::cluster_group cluster = cg::this_cluster();
    const int S = (int)cluster.num_blocks(), rank = (int)cluster.block_rank();
    const int c = blockIdx.y;
    const int q = threadIdx.x % QUADS, rg = threadIdx.x / QUADS;
    const int ch = (blockIdx.x / S) * (QUADS * 4) + q * 4;
    const bool live = ch < p.C;
    const int R = p.B * p.HW, Rc = R / S;
    const long long base = ((long long)c * R + (long long)rank * Rc) * p.C + ch;
    const float4* xb = reinterpret_cast<const float4*>(p.x + base);
    const long long rs = p.C / 4;
    const float m = (float)R;
    float4 s = make_float4(0.f, 0.f, 0.f, 0.f);
    if (live) {
#pragma unroll 8
        for (int r = rg; r < Rc; r += GROUPS) {
            const float4 v = __ldcs(&xb[(long long)r * rs]);
            bn_tile[r * QUADS + q] = v;
            s.x += v.x; s.y += v.y; s.z += v.z; s.w += v.w;
        }
    }
    float4 mean = bn_cl_exchange<QUADS>(cluster, part, bn_cl_reduce4<QUADS>(s, red), 0);
    mean.x /= m; mean.y /= m; mean.z /= m; mean.w /= m;
    float4 qv = make_float4(0.f, 0.f, 0.f, 0.f);
    if (live) {
#pragma unroll 8
        for (int r = rg; r < Rc; r += GROUPS) {
            const float4 v = bn_tile[r * QUADS + q];
            float d;
            d = v.x - mean.x; qv.x = fmaf(d, d, qv.x);
            d = v.y - mean.y; qv.y = fmaf(d, d, qv.y);
            d = v.z - mean.z; qv.z = fmaf(d, d, qv.z);
            d = v.w - mean.w; qv.w = fmaf(d, d, qv.w);
        }
    }
    const float4 var = bn_cl_exchange<QUADS>(cluster, part, bn_cl_reduce4<QUADS>(qv, red), 1);
    if (live) {
        const float4 rstd = make_float4(rsqrtf(var.x / m + p.eps), rsqrtf(var.y / m + p.eps),
                                        rsqrtf(var.z / m + p.eps), rsqrtf(var.w / m + p.eps));
        const float4 ga = *reinterpret_cast<const float4*>(p.gamma + ch);
        const float4 be = *reinterpret_cast<const float4*>(p.beta + ch);
        float4 g, sh;
        bn_affine(ga, be, mean, rstd, g, sh);
        float4* yb = reinterpret_cast<float4*>(p.y + base);
        const float4* rb = p.res ? reinterpret_cast<const float4*>(p.res + base) : nullptr;
#pragma unroll 8
        for (int r = rg; r < Rc; r += GROUPS) {
            const float4 v = bn_tile[r * QUADS + q];
            yb[(long long)r * rs] = bn_act(bn_pre(v, g, sh), rb, (long long)r * rs, p.relu);
        }
        if (rg == 0 && rank == 0) {
            *reinterpret_cast<float4*>(p.mean + c * p.C + ch) = mean;
            *reinterpret_cast<float4*>(p.rstd + c * p.C + ch) = rstd;
        }
    }
    cluster.sync();                 // peers may still be reading this CTA's partials
}

template <int QUADS>
__global__ void __launch_bounds__(256)
client_bn_nhwc_bwd_cl_kernel(const __grid_constant__ ClientBNParams p) {
    constexpr int GROUPS = 256 / QUADS;
    extern __shared__ __align__(16) float4 bn_tile[];          // x tile [Rc][QUADS], then gy tile [Rc][QUADS]
    __shared__ float4 red[256];
    __shared__ float4 part[2 * QUADS];
    cg::cluster_group cluster = cg::this_cluster();
    const int S = (int)cluster.num_blocks(), rank = (int)cluster.block_rank();
    const int c = blockIdx.y;
    const int q = threadIdx.x % QUADS, rg = threadIdx.x / QUADS;
    const int ch = (blockIdx.x / S) * (QUADS * 4) + q * 4;
    const bool live = ch < p.C;
    const int R = p.B * p.HW, Rc = R / S;
    const long long base = ((long long)c * R + (long long)rank * Rc) * p.C + ch;
    const float4* xb = reinterpret_cast<const float4*>(p.x + base);
    const float4* gb = reinterpret_cast<const float4*>(p.gy + base);
    const float4* ab = reinterpret_cast<const float4*>(p.act + base);       // only dereferenced when p.relu
    float4* mb = p.gmask ? reinterpret_cast<float4*>(p.gmask + base) : nullptr;
    float4* xt = bn_tile;
    float4* gt = bn_tile + (size_t)Rc * QUADS;
    const long long rs = p.C / 4;
    float4 mean = make_float4(0.f, 0.f, 0.f, 0.f), rstd = mean;
    if (live) {
        mean = *reinterpret_cast<const float4*>(p.mean + c * p.C + ch);
        rstd = *reinterpret_cast<const float4*>(p.rstd + c * p.C + ch);
    }
    const bool remask = p.relu && p.act == nullptr;
    float4 fg = make_float4(0.f, 0.f, 0.f, 0.f), fsh = fg;
    if (live && remask)
        bn_affine(*reinterpret_cast<const float4*>(p.gamma + ch), *reinterpret_cast<const float4*>(p.beta + ch), mean, rstd,
                  fg, fsh);
    float4 sb = make_float4(0.f, 0.f, 0.f, 0.f), sg = sb;
    if (live) {
#pragma unroll 4
        for (int r = rg; r < Rc; r += GROUPS) {
            float4 g = __ldcs(&gb[(long long)r * rs]);
            const float4 v = __ldcs(&xb[(long long)r * rs]);
            if (p.relu) {
                float4 a;
                if (remask) a = bn_pre(v, fg, fsh);
                else a = __ldcs(&ab[(long long)r * rs]);
                g = bn_relu_mask(g, a);
            }
            if (mb != nullptr) mb[(long long)r * rs] = g;
            const float4 xh = make_float4((v.x - mean.x) * rstd.x, (v.y - mean.y) * rstd.y, (v.z - mean.z) * rstd.z,
                                          (v.w - mean.w) * rstd.w);
            xt[r * QUADS + q] = xh;             // keep xhat, not x
            gt[r * QUADS + q] = g;
            sb.x += g.x; sb.y += g.y; sb.z += g.z; sb.w += g.w;
            sg.x = fmaf(g.x, xh.x, sg.x); sg.y = fmaf(g.y, xh.y, sg.y);
            sg.z = fmaf(g.z, xh.z, sg.z); sg.w = fmaf(g.w, xh.w, sg.w);
        }
    }
    const float4 pb = bn_cl_reduce4<QUADS>(sb, red);
    const float4 pg = bn_cl_reduce4<QUADS>(sg, red);
    if (rg == 0) part[QUADS + q] = pg;
    const float4 dbeta = bn_cl_exchange<QUADS>(cluster, part, pb, 0);      // its cluster.sync also publishes part[1]
    float4 dgamma = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int j = 0; j < S; ++j) {
        const float4 t = cluster.map_shared_rank(part, j)[QUADS + q];
        dgamma.x += t.x; dgamma.y += t.y; dgamma.z += t.z; dgamma.w += t.w;
    }
    if (live) {
        if (rg == 0 && rank == 0) {
            float* dg = p.dgamma + (long long)c * p.ld + ch;
            float* db = p.dbeta + (long long)c * p.ld + ch;
            dg[0] = bl_sanitize(p.alpha * dgamma.x); dg[1] = bl_sanitize(p.alpha * dgamma.y);
            dg[2] = bl_sanitize(p.alpha * dgamma.z); dg[3] = bl_sanitize(p.alpha * dgamma.w);
            db[0] = bl_sanitize(p.alpha * dbeta.x); db[1] = bl_sanitize(p.alpha * dbeta.y);
            db[2] = bl_sanitize(p.alpha * dbeta.z); db[3] = bl_sanitize(p.alpha * dbeta.w);
        }
        if (p.y != nullptr) {
            const float inv_m = 1.f / (float)R;
            const float4 ga = *reinterpret_cast<const float4*>(p.gamma + ch);
            const float4 k = make_float4(ga.x * rstd.x, ga.y * rstd.y, ga.z * rstd.z, ga.w * rstd.w);
            float4* yb = reinterpret_cast<float4*>(p.y + base);
#pragma unroll 8
            for (int r = rg; r < Rc; r += GROUPS) {
                const float4 g = gt[r * QUADS + q], xh = xt[r * QUADS + q];
                float4 o;
                o.x = k.x * (g.x - (dbeta.x + xh.x * dgamma.x) * inv_m);
                o.y = k.y * (g.y - (dbeta.y + xh.y * dgamma.y) * inv_m);
                o.z = k.z * (g.z - (dbeta.z + xh.z * dgamma.z) * inv_m);
                o.w = k.w * (g.w - (dbeta.w + xh.w * dgamma.w) * inv_m);
                yb[(long long)r * rs] = o;
            }
        }
    }
    cluster.sync();
}

namespace {
constexpr size_t kBnClMaxSmem = 200 * 1024;

// Cluster size S (power of two <= 8, dividing R) and tile width for `tiles_per_row` staged tiles per row:
// smallest S whose tile fits 64 KB, grown while the grid is below ~4 CTAs per SM; 16-channel tiles when even
// S = 8 needs more than 96 KB with 32 channels.  Returns false when nothing fits (caller uses the plain kernel).
bool bn_cl_plan(const ClientBNParams* p, int staged, int* S_out, int* quads_out, size_t* smem_out) {
    static const bool enabled = [] { const char* e = getenv("BLADES_BN_CLUSTER"); return !(e && e[0] == '0'); }();
    if (!enabled) return false;
    const long long R = (long long)p->B * p->HW;
    // measured (profiles/round_launches_r1.txt vs round_kernels2): the cluster form wins for long slices staged in
    // <= 64 KB per CTA (stem BN fwd 269 -> 175 us, layer1 fwd 40 -> 28 us, bwd 74 -> 64 us) and loses for short
    // slices (R <= 512: 12 -> 18 us, launch/sync overhead) and for the stem backward, whose two staged tiles need
    // 128 KB per CTA (one CTA per SM: 277 -> 425 us).
    if (R < 1024) return false;
    int quads = 8;
    auto bytes = [&](int S, int qd) { return (size_t)(R / S) * qd * 16 * staged; };
    int S = 1;
    while (S < 8 && R % (2 * S) == 0 && bytes(S, quads) > 64 * 1024) S *= 2;
    if (bytes(S, quads) > 96 * 1024 && p->C % 16 == 0) quads = 4;
    if (bytes(S, quads) > 64 * 1024) return false;
    auto ctas = [&](int S_, int qd) { return (long long)((p->C + qd * 4 - 1) / (qd * 4)) * p->n * S_; };
    while (S < 8 && R % (2 * S) == 0 && R / (2 * S) >= 256 / quads && ctas(S, quads) < 600) S *= 2;
    *S_out = S; *quads_out = quads; *smem_out = bytes(S, quads);
    return true;
}

template <typename K>
int bn_cl_launch(K kernel, const ClientBNParams* p, int S, int quads, size_t smem, cudaStream_t stream) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned)(((p->C + quads * 4 - 1) / (quads * 4)) * S), (unsigned)p->n);
    cfg.blockDim = dim3(256);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = (unsigned)S; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr; cfg.numAttrs = 1;
    return (int)cudaLaunchKernelEx(&cfg, kernel, *p);
}

void bn_cl_set_attrs() {
    static const bool once = [] {
        cudaFuncSetAttribute(client_bn_nhwc_fwd_cl_kernel<8>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kBnClMaxSmem);
        cudaFuncSetAttribute(client_bn_nhwc_fwd_cl_kernel<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kBnClMaxSmem);
        cudaFuncSetAttribute(client_bn_nhwc_bwd_cl_kernel<8>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kBnClMaxSmem);
        cudaFuncSetAttribute(client_bn_nhwc_bwd_cl_kernel<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kBnClMaxSmem);
        return true;
    }();
    (void)once;
}
}  // namespace

// The plan the launchers below would use (no CUDA call: usable without a GPU by the tests).  ``staged`` = 1 for
// the forward (x tile), 2 for the backward (xhat and gy tiles).  Returns the cluster size S (0 = plain kernel) and
// writes the channel quads per tile and the dynamic shared memory per CTA.
extern "C" int bl_client_bn_cluster_plan(int n, int B, int C, int HW, int staged, int* quads_out, long long* smem_out) {
    ClientBNParams p = {};
    p.n = n; p.B = B; p.C = C; p.HW = HW;
    int S = 0, quads = 0;
    size_t smem = 0;
    if (C % 4 != 0 || !bn_cl_plan(&p, staged, &S, &quads, &smem)) return 0;
    if (quads_out) *quads_out = quads;
    if (smem_out) *smem_out = (long long)smem;
    return S;
}

extern "C" int bl_client_bn_nhwc_fwd(const ClientBNParams* p, void* stream) {
    if (p->C % 4 != 0) return -1;
    int S, quads; size_t smem;
    if (bn_cl_plan(p, 1, &S, &quads, &smem)) {
        bn_cl_set_attrs();
        return quads == 8 ? bn_cl_launch(client_bn_nhwc_fwd_cl_kernel<8>, p, S, quads, smem, (cudaStream_t)stream)
                          : bn_cl_launch(client_bn_nhwc_fwd_cl_kernel<4>, p, S, quads, smem, (cudaStream_t)stream);
    }
    dim3 grid((p->C + kBnTile - 1) / kBnTile, p->n);
    client_bn_nhwc_fwd_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(*p);
    return (int)cudaGetLastError();
}
extern "C" int bl_client_bn_nhwc_bwd(const ClientBNParams* p, void* stream) {
    if (p->C % 4 != 0) return -1;
    int S, quads; size_t smem;
    if (bn_cl_plan(p, 2, &S, &quads, &smem)) {
        bn_cl_set_attrs();
        return quads == 8 ? bn_cl_launch(client_bn_nhwc_bwd_cl_kernel<8>, p, S, quads, smem, (cudaStream_t)stream)
                          : bn_cl_launch(client_bn_nhwc_bwd_cl_kernel<4>, p, S, quads, smem, (cudaStream_t)stream);
    }
    dim3 grid((p->C + kBnTile - 1) / kBnTile, p->n);
    client_bn_nhwc_bwd_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(*p);
    return (int)cudaGetLastError();
}
