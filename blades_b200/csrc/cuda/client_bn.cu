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
};

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
// NHWC (channels_last) variants: x is physically [n*B*HW rows][C]; a block owns (client, 32-channel tile):
// thread = (row group 0..7, channel 0..31), so each row segment is one coalesced 128 B read; the block
// strides over the client's R = B*HW rows three times (sum, squared deviation, normalise) -- passes two and
// three are served by L2.  Any HW / C is supported.
constexpr int kBnTile = 32, kBnGroups = 8;

__device__ __forceinline__ float bn_group_reduce(float v, float (*red)[kBnTile]) {
    const int ch = threadIdx.x % kBnTile, rg = threadIdx.x / kBnTile;
    __syncthreads();
    red[rg][ch] = v;
    __syncthreads();
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < kBnGroups; ++i) s += red[i][ch];
    return s;
}

__global__ void __launch_bounds__(256)
client_bn_nhwc_fwd_kernel(const __grid_constant__ ClientBNParams p) {
    __shared__ float red[kBnGroups][kBnTile];
    const int c = blockIdx.y;
    const int ch = blockIdx.x * kBnTile + threadIdx.x % kBnTile, rg = threadIdx.x / kBnTile;
    const bool live = ch < p.C;
    const int R = p.B * p.HW;
    const float* xb = p.x + (long long)c * R * p.C + ch;
    const float m = (float)R;
    float s = 0.f;
    if (live) for (int r = rg; r < R; r += kBnGroups) s += xb[(long long)r * p.C];
    const float mean = bn_group_reduce(s, red) / m;
    float q = 0.f;
    if (live) for (int r = rg; r < R; r += kBnGroups) { const float d = xb[(long long)r * p.C] - mean; q = fmaf(d, d, q); }
    const float var = bn_group_reduce(q, red) / m;
    const float rstd = rsqrtf(var + p.eps);
    if (live) {
        const float g = p.gamma[ch] * rstd, sh = p.beta[ch] - mean * g;
        float* yb = p.y + (long long)c * R * p.C + ch;
        for (int r = rg; r < R; r += kBnGroups) yb[(long long)r * p.C] = fmaf(xb[(long long)r * p.C], g, sh);
        if (rg == 0) { p.mean[c * p.C + ch] = mean; p.rstd[c * p.C + ch] = rstd; }
    }
}

__global__ void __launch_bounds__(256)
client_bn_nhwc_bwd_kernel(const __grid_constant__ ClientBNParams p) {
    __shared__ float red[kBnGroups][kBnTile];
    const int c = blockIdx.y;
    const int ch = blockIdx.x * kBnTile + threadIdx.x % kBnTile, rg = threadIdx.x / kBnTile;
    const bool live = ch < p.C;
    const int R = p.B * p.HW;
    const long long base = (long long)c * R * p.C + ch;
    const float mean = live ? p.mean[c * p.C + ch] : 0.f;
    const float rstd = live ? p.rstd[c * p.C + ch] : 0.f;
    float sb = 0.f, sg = 0.f;
    if (live)
        for (int r = rg; r < R; r += kBnGroups) {
            const float g = p.gy[base + (long long)r * p.C];
            const float xh = (p.x[base + (long long)r * p.C] - mean) * rstd;
            sb += g;
            sg = fmaf(g, xh, sg);
        }
    const float dbeta = bn_group_reduce(sb, red);
    const float dgamma = bn_group_reduce(sg, red);
    if (live) {
        if (rg == 0) {
            p.dgamma[(long long)c * p.ld + ch] = bl_sanitize(p.alpha * dgamma);
            p.dbeta[(long long)c * p.ld + ch] = bl_sanitize(p.alpha * dbeta);
        }
        if (p.y != nullptr) {
            const float inv_m = 1.f / (float)R;
            const float k = p.gamma[ch] * rstd;
            for (int r = rg; r < R; r += kBnGroups) {
                const float g = p.gy[base + (long long)r * p.C];
                const float xh = (p.x[base + (long long)r * p.C] - mean) * rstd;
                p.y[base + (long long)r * p.C] = k * (g - (dbeta + xh * dgamma) * inv_m);
            }
        }
    }
}

extern "C" int bl_client_bn_nhwc_fwd(const ClientBNParams* p, void* stream) {
    dim3 grid((p->C + kBnTile - 1) / kBnTile, p->n);
    client_bn_nhwc_fwd_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(*p);
    return (int)cudaGetLastError();
}
extern "C" int bl_client_bn_nhwc_bwd(const ClientBNParams* p, void* stream) {
    dim3 grid((p->C + kBnTile - 1) / kBnTile, p->n);
    client_bn_nhwc_bwd_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(*p);
    return (int)cudaGetLastError();
}
