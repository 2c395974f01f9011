// Small memory-bound pieces of the client-batched training pass that used to be ATen / cuDNN launches
// (reference call site: the autograd graph behind client.py:178-193):
//   * NHWC max pooling forward (value + 1-byte argmax) and backward (gather form, no atomics)
//   * NHWC global average pooling forward / backward
//   * per-client cross-entropy: mean CE over each client's B samples, clamped like the reference's
//     torch.clamp(loss, 0, 1e6) (client.py:146), and d(sum_c loss_c)/d(logits) in one launch
//   * per-client column sums (bias gradients) written into the update matrix with the -lr scale
//   * row-padding copy (weight matrices whose row length is not a multiple of 4 floats cannot be TMA sources)
#include "common.cuh"

// ------------------------------------------------------------------------------------------ max pooling
struct PoolParams {
    const float* x;          // fwd: input [NB][H][W][C];  bwd: gy [NB][Ho][Wo][C]
    float* y;                // fwd: output [NB][Ho][Wo][C]; bwd: gx [NB][H][W][C]
    unsigned char* idx;      // [NB][Ho][Wo][C] argmax position r*k + s inside the window
    int NB, H, W, C, Ho, Wo, k, s, p;
};

// One block per output row (b, ho); threads run over (wo, channel quad) with 32-bit index arithmetic -- the flat
// grid-stride form spent its time in 64-bit divisions (stem pooling: 100 us forward / 210 us backward against
// 40 us of memory traffic).
__global__ void __launch_bounds__(256)
maxpool_nhwc_fwd_kernel(const __grid_constant__ PoolParams p) {
    const int c4 = p.C >> 2;
    const unsigned ho = blockIdx.x % (unsigned)p.Ho;
    const long long b = blockIdx.x / (unsigned)p.Ho;
    const float* xb = p.x + b * p.H * p.W * p.C;
    for (unsigned e = threadIdx.x; e < (unsigned)(p.Wo * c4); e += blockDim.x) {
        const int wo = (int)(e / (unsigned)c4);
        const int c = (int)(e - (unsigned)wo * (unsigned)c4) * 4;
        float4 best = make_float4(-INFINITY, -INFINITY, -INFINITY, -INFINITY);
        int bx = -1, by = -1, bz = -1, bw = -1;
        for (int r = 0; r < p.k; ++r) {
            const int h = (int)ho * p.s - p.p + r;
            if (h < 0 || h >= p.H) continue;
            for (int q = 0; q < p.k; ++q) {
                const int w = wo * p.s - p.p + q;
                if (w < 0 || w >= p.W) continue;
                const float4 v = *reinterpret_cast<const float4*>(xb + (h * p.W + w) * p.C + c);
                const int pos = r * p.k + q;
                // ATen's rule: start at the first window element; replace when larger or NaN
                if (bx < 0 || v.x > best.x || v.x != v.x) { best.x = v.x; bx = pos; }
                if (by < 0 || v.y > best.y || v.y != v.y) { best.y = v.y; by = pos; }
                if (bz < 0 || v.z > best.z || v.z != v.z) { best.z = v.z; bz = pos; }
                if (bw < 0 || v.w > best.w || v.w != v.w) { best.w = v.w; bw = pos; }
            }
        }
        const long long o = ((b * p.Ho + ho) * p.Wo + wo) * p.C + c;
        *reinterpret_cast<float4*>(p.y + o) = best;
        *reinterpret_cast<uchar4*>(p.idx + o) = make_uchar4((unsigned char)bx, (unsigned char)by, (unsigned char)bz,
                                                            (unsigned char)bw);
    }
}

// gather form (no atomics), one block per input row (b, h)
__global__ void __launch_bounds__(256)
maxpool_nhwc_bwd_kernel(const __grid_constant__ PoolParams p) {
    const int c4 = p.C >> 2;
    const int h = (int)(blockIdx.x % (unsigned)p.H);
    const long long b = blockIdx.x / (unsigned)p.H;
    // output windows containing row h: ho*s - p <= h <= ho*s - p + k - 1   (uniform per block)
    int ho0 = (h + p.p - p.k + 1 + p.s - 1); ho0 = ho0 < 0 ? 0 : ho0 / p.s;
    const int ho1 = min((h + p.p) / p.s, p.Ho - 1);
    const float* gb = p.x + b * p.Ho * p.Wo * p.C;
    const unsigned char* ib = p.idx + b * p.Ho * p.Wo * p.C;
    for (unsigned e = threadIdx.x; e < (unsigned)(p.W * c4); e += blockDim.x) {
        const int w = (int)(e / (unsigned)c4);
        const int c = (int)(e - (unsigned)w * (unsigned)c4) * 4;
        float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
        int wo0 = (w + p.p - p.k + 1 + p.s - 1); wo0 = wo0 < 0 ? 0 : wo0 / p.s;
        const int wo1 = min((w + p.p) / p.s, p.Wo - 1);
        for (int ho = ho0; ho <= ho1; ++ho)
            for (int wo = wo0; wo <= wo1; ++wo) {
                const int pos = (h - (ho * p.s - p.p)) * p.k + (w - (wo * p.s - p.p));
                const int o = (ho * p.Wo + wo) * p.C + c;
                const uchar4 id = *reinterpret_cast<const uchar4*>(ib + o);
                const float4 g = *reinterpret_cast<const float4*>(gb + o);
                if (id.x == pos) acc.x += g.x;
                if (id.y == pos) acc.y += g.y;
                if (id.z == pos) acc.z += g.z;
                if (id.w == pos) acc.w += g.w;
            }
        *reinterpret_cast<float4*>(p.y + ((b * p.H + h) * p.W + w) * p.C + c) = acc;
    }
}

static int pool_grid(long long total) {
    long long g = (total + 255) / 256;
    if (g > 148LL * 16) g = 148LL * 16;
    return (int)(g < 1 ? 1 : g);
}

extern "C" int bl_maxpool_nhwc_fwd(const float* x, float* y, unsigned char* idx, int NB, int H, int W, int C, int Ho,
                                   int Wo, int k, int s, int pad, void* stream) {
    if (C % 4 != 0 || k * k > 255) return -1;
    PoolParams p{x, y, idx, NB, H, W, C, Ho, Wo, k, s, pad};
    if ((long long)NB * Ho > 0x7fffffffLL || (long long)H * W * C > 0x7fffffffLL) return -1;
    const int tf = Wo * (C / 4);
    maxpool_nhwc_fwd_kernel<<<(unsigned)(NB * Ho), tf >= 256 ? 256 : (tf + 31) / 32 * 32, 0, (cudaStream_t)stream>>>(p);
    return (int)cudaGetLastError();
}
extern "C" int bl_maxpool_nhwc_bwd(const float* gy, float* gx, const unsigned char* idx, int NB, int H, int W, int C,
                                   int Ho, int Wo, int k, int s, int pad, void* stream) {
    if (C % 4 != 0 || k * k > 255) return -1;
    PoolParams p{gy, gx, const_cast<unsigned char*>(idx), NB, H, W, C, Ho, Wo, k, s, pad};
    if ((long long)NB * H > 0x7fffffffLL || (long long)H * W * C > 0x7fffffffLL) return -1;
    const int tb = W * (C / 4);
    maxpool_nhwc_bwd_kernel<<<(unsigned)(NB * H), tb >= 256 ? 256 : (tb + 31) / 32 * 32, 0, (cudaStream_t)stream>>>(p);
    return (int)cudaGetLastError();
}

// ------------------------------------------------------------------------------------------ global average pooling
__global__ void __launch_bounds__(256)
avgpool_nhwc_fwd_kernel(const float* __restrict__ x, float* __restrict__ y, long long NB, int HW, int C) {
    const int c4 = C >> 2;
    const long long total = NB * c4;
    const float inv = 1.f / (float)HW;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const int c = (int)(i % c4) * 4;
        const long long b = i / c4;
        float4 s = make_float4(0.f, 0.f, 0.f, 0.f);
        for (int j = 0; j < HW; ++j) {
            const float4 v = *reinterpret_cast<const float4*>(x + (b * HW + j) * C + c);
            s.x += v.x; s.y += v.y; s.z += v.z; s.w += v.w;
        }
        *reinterpret_cast<float4*>(y + b * C + c) = make_float4(s.x * inv, s.y * inv, s.z * inv, s.w * inv);
    }
}
__global__ void __launch_bounds__(256)
avgpool_nhwc_bwd_kernel(const float* __restrict__ gy, float* __restrict__ gx, long long NB, int HW, int C) {
    const int c4 = C >> 2;
    const long long total = NB * HW * c4;
    const float inv = 1.f / (float)HW;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const int c = (int)(i % c4) * 4;
        const long long b = (i / c4) / HW;
        const float4 g = *reinterpret_cast<const float4*>(gy + b * C + c);
        *reinterpret_cast<float4*>(gx + (i / c4) * C + c) = make_float4(g.x * inv, g.y * inv, g.z * inv, g.w * inv);
    }
}
extern "C" int bl_avgpool_nhwc_fwd(const float* x, float* y, long long NB, int HW, int C, void* stream) {
    if (C % 4 != 0) return -1;
    avgpool_nhwc_fwd_kernel<<<pool_grid(NB * (C / 4)), 256, 0, (cudaStream_t)stream>>>(x, y, NB, HW, C);
    return (int)cudaGetLastError();
}
extern "C" int bl_avgpool_nhwc_bwd(const float* gy, float* gx, long long NB, int HW, int C, void* stream) {
    if (C % 4 != 0) return -1;
    avgpool_nhwc_bwd_kernel<<<pool_grid(NB * HW * (C / 4)), 256, 0, (cudaStream_t)stream>>>(gy, gx, NB, HW, C);
    return (int)cudaGetLastError();
}

// ------------------------------------------------------------------------------------------ per-client cross-entropy
// One block per client.  loss_c = mean_b CE(logits[c*B+b], target[c*B+b]); the objective of the fused pass is
// sum_c min(max(loss_c, 0), clamp_c), so  dlogits = (softmax - onehot) / B  where loss_c < clamp_c, else 0.
struct ClientCEParams {
    const float* logits;     // [n*B][ldl]
    const long long* target; // [n*B]
    const float* clamp;      // [n]
    float* loss;             // [n]  (unclamped mean, like the reference logs it)
    float* dlogits;          // [n*B][ldg]; columns >= C are written as zeros (nullptr: evaluation only)
    float* hits;             // optional [n]: number of samples whose argmax equals the target (top-1 evaluation)
    int n, B, C, ldl, ldg;
};

__global__ void __launch_bounds__(128)
client_ce_kernel(const __grid_constant__ ClientCEParams p) {
    __shared__ float red[128];
    __shared__ float redh[128];
    __shared__ float s_scale;
    const int c = blockIdx.x;
    float local = 0.f, nhit = 0.f;
    for (int b = threadIdx.x; b < p.B; b += blockDim.x) {
        const float* z = p.logits + (long long)(c * p.B + b) * p.ldl;
        float m = -INFINITY;
        int am = 0;
        for (int j = 0; j < p.C; ++j)
            if (z[j] > m) { m = z[j]; am = j; }              // first maximum, like torch.argmax
        float se = 0.f;
        for (int j = 0; j < p.C; ++j) se += expf(z[j] - m);
        const long long t = p.target[c * p.B + b];
        const float zt = (t >= 0 && t < p.C) ? z[t] : 0.f;
        local += (m + logf(se)) - zt;
        nhit += (am == t) ? 1.f : 0.f;
    }
    red[threadIdx.x] = local;
    redh[threadIdx.x] = nhit;
    __syncthreads();
    for (int o = 64; o > 0; o >>= 1) {
        if ((int)threadIdx.x < o) { red[threadIdx.x] += red[threadIdx.x + o]; redh[threadIdx.x] += redh[threadIdx.x + o]; }
        __syncthreads();
    }
    if (threadIdx.x == 0) {
        const float mean = red[0] / (float)p.B;
        p.loss[c] = mean;
        if (p.hits != nullptr) p.hits[c] = redh[0];
        // d/dx min(max(x, 0), clamp): 1 inside (0, clamp); NaN losses give no gradient
        s_scale = (p.clamp != nullptr && mean > 0.f && mean < p.clamp[c]) ? 1.f / (float)p.B : 0.f;
    }
    if (p.dlogits == nullptr) return;
    __syncthreads();
    const float scale = s_scale;
    for (int b = threadIdx.x; b < p.B; b += blockDim.x) {
        const float* z = p.logits + (long long)(c * p.B + b) * p.ldl;
        float* g = p.dlogits + (long long)(c * p.B + b) * p.ldg;
        float m = -INFINITY;
        for (int j = 0; j < p.C; ++j) m = fmaxf(m, z[j]);
        float se = 0.f;
        for (int j = 0; j < p.C; ++j) se += expf(z[j] - m);
        const float inv = 1.f / se;
        const long long t = p.target[c * p.B + b];
        for (int j = 0; j < p.ldg; ++j) {
            float v = 0.f;
            if (j < p.C) v = (expf(z[j] - m) * inv - (j == t ? 1.f : 0.f)) * scale;
            g[j] = v;
        }
    }
}
extern "C" int bl_client_ce(const float* logits, const long long* target, const float* clamp, float* loss,
                            float* dlogits, float* hits, int n, int B, int C, int ldl, int ldg, void* stream) {
    if (n < 1 || B < 1 || C < 1 || (dlogits != nullptr && ldg < C) || ldl < C) return -1;
    ClientCEParams p{logits, target, clamp, loss, dlogits, hits, n, B, C, ldl, ldg};
    client_ce_kernel<<<n, 128, 0, (cudaStream_t)stream>>>(p);
    return (int)cudaGetLastError();
}

// ------------------------------------------------------------------------------------------ per-client column sums
// out[c*ld_out + j] = sanitize(alpha * sum_t g[(c*T + t)*ldg + j])     (bias gradients into the update rows)
__global__ void __launch_bounds__(128)
client_colsum_kernel(const float* __restrict__ g, float* __restrict__ out, int T, int C, long long ldg,
                     long long ld_out, float alpha) {
    const int c = blockIdx.y;
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= C) return;
    const float* gp = g + (long long)c * T * ldg + j;
    float s = 0.f;
    for (int t = 0; t < T; ++t) s += gp[(long long)t * ldg];
    out[(long long)c * ld_out + j] = bl_sanitize(alpha * s);
}
extern "C" int bl_client_colsum(const float* g, float* out, int n, int T, int C, long long ldg, long long ld_out,
                                float alpha, void* stream) {
    if (n < 1 || C < 1) return 0;
    dim3 grid((C + 127) / 128, n);
    client_colsum_kernel<<<grid, 128, 0, (cudaStream_t)stream>>>(g, out, T, C, ldg, ld_out, alpha);
    return (int)cudaGetLastError();
}

// ------------------------------------------------------------------------------------------ row padding copy
__global__ void __launch_bounds__(256)
pad_rows_kernel(const float* __restrict__ src, float* __restrict__ dst, long long rows, int cols, long long lds,
                int ldd) {
    const long long total = rows * ldd;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const long long r = i / ldd;
        const int j = (int)(i - r * ldd);
        dst[i] = j < cols ? src[r * lds + j] : 0.f;
    }
}
extern "C" int bl_pad_rows(const float* src, float* dst, long long rows, int cols, long long lds, int ldd, void* stream) {
    if (rows < 1) return 0;
    pad_rows_kernel<<<pool_grid(rows * ldd), 256, 0, (cudaStream_t)stream>>>(src, dst, rows, cols, lds, ldd);
    return (int)cudaGetLastError();
}

// ---------------------------------------------------------------------------------------------
// out[i] = nan_to_num(a[i] - b[i]): the client update "theta_after - theta_before" of the time-sliced (fedavg /
// custom-client) path, sanitised where it is produced (reference client.py:195-198) -- one streaming launch.
__global__ void __launch_bounds__(256)
diff_rows_kernel(const float* __restrict__ a, const float* __restrict__ b, float* __restrict__ out, long long n4,
                 long long n) {
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += stride) {
        const float4 x = reinterpret_cast<const float4*>(a)[i], y = reinterpret_cast<const float4*>(b)[i];
        reinterpret_cast<float4*>(out)[i] = make_float4(bl_sanitize(x.x - y.x), bl_sanitize(x.y - y.y),
                                                        bl_sanitize(x.z - y.z), bl_sanitize(x.w - y.w));
    }
    for (long long i = n4 * 4 + (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride)
        out[i] = bl_sanitize(a[i] - b[i]);
}

extern "C" int bl_diff_rows(const float* a, const float* b, float* out, long long n, void* stream) {
    if (n <= 0) return 0;
    const bool vec = ((uintptr_t)a % 16 == 0) && ((uintptr_t)b % 16 == 0) && ((uintptr_t)out % 16 == 0);
    const long long n4 = vec ? n / 4 : 0;
    long long work = vec ? n4 : n;
    unsigned grid = (unsigned)((work + 255) / 256);
    if (grid > 148u * 8) grid = 148u * 8;
    if (grid < 1) grid = 1;
    diff_rows_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(a, b, out, n4, n);
    return (int)cudaGetLastError();
}
