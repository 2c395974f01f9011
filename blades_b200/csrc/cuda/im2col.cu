// Row-major im2col for the per-client weight-gradient GEMM (K9): NCHW activations ->
// cols[(b, ho, wo)][(cin, r, s)], i.e. exactly the [n_clients*T, K] "MN-major" B operand the grouped
// wgrad kernel consumes and with K ordered like the flattened conv weight [Cout][Cin*kh*kw].
// (torch's F.unfold launches one small kernel per sample -- 64 000 launches for a 3200-image batch --
// and produces the transposed [K, L] layout; this is one launch with coalesced 128 B row writes.)
#include "common.cuh"

struct Im2colParams {
    const float* x;
    float* out;
    int NB, Cin, H, W, kh, kw, sh, sw, ph, pw, dh, dw, Ho, Wo;
    long long rows;     // NB*Ho*Wo
    int K;              // Cin*kh*kw
};

constexpr int kRowsPerBlock = 8;

__global__ void __launch_bounds__(256)
im2col_rows_kernel(const __grid_constant__ Im2colParams p) {
    __shared__ long long s_base[kRowsPerBlock];
    __shared__ int s_h0[kRowsPerBlock], s_w0[kRowsPerBlock];
    const long long row0 = (long long)blockIdx.x * kRowsPerBlock;
    if (threadIdx.x < kRowsPerBlock) {
        const long long row = row0 + threadIdx.x;
        if (row < p.rows) {
            const int L = p.Ho * p.Wo;
            const long long b = row / L;
            const int l = (int)(row - b * L);
            const int ho = l / p.Wo, wo = l - ho * p.Wo;
            s_base[threadIdx.x] = b * (long long)p.Cin * p.H * p.W;
            s_h0[threadIdx.x] = ho * p.sh - p.ph;
            s_w0[threadIdx.x] = wo * p.sw - p.pw;
        } else {
            s_base[threadIdx.x] = -1;
        }
    }
    __syncthreads();
    const int taps = p.kh * p.kw;
    const long long HW = (long long)p.H * p.W;
    for (int col = threadIdx.x; col < p.K; col += blockDim.x) {
        const int cin = col / taps;
        const int t = col - cin * taps;
        const int r = t / p.kw, s = t - r * p.kw;
        const int dh = r * p.dh, dw = s * p.dw;
        const long long coff = cin * HW;
#pragma unroll
        for (int i = 0; i < kRowsPerBlock; ++i) {
            const long long base = s_base[i];
            if (base < 0) break;
            const int h = s_h0[i] + dh, w = s_w0[i] + dw;
            float v = 0.f;
            if (h >= 0 && h < p.H && w >= 0 && w < p.W) v = __ldg(p.x + base + coff + (long long)h * p.W + w);
            p.out[(row0 + i) * p.K + col] = v;
        }
    }
}

extern "C" int bl_im2col_rows(const float* x, float* out, int NB, int Cin, int H, int W, int kh, int kw, int sh,
                              int sw, int ph, int pw, int dh, int dw, int Ho, int Wo, void* stream) {
    Im2colParams p{x, out, NB, Cin, H, W, kh, kw, sh, sw, ph, pw, dh, dw, Ho, Wo, (long long)NB * Ho * Wo,
                   Cin * kh * kw};
    if (p.rows <= 0) return 0;
    const long long blocks = (p.rows + kRowsPerBlock - 1) / kRowsPerBlock;
    if (blocks > 0x7fffffffLL) return -1;
    im2col_rows_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(p);
    return (int)cudaGetLastError();
}

// ---------------------------------------------------------------------------------------------
// NHWC variant: x [NB, H, W, Cin] (the physical layout of a channels_last tensor) ->
// cols[(b, ho, wo)][(r, s, cin)] with row stride ldk (>= K = kh*kw*Cin, padded to a multiple of 4 so the
// wgrad kernel's TMA row stride is 16 B aligned; pad columns are written as zeros).  K is ordered like a
// channels_last conv weight [Cout][kh][kw][Cin]; every tap is one contiguous run of Cin floats, moved as
// float4 when Cin % 4 == 0.
struct Im2colNhwcParams {
    const float* x;
    float* out;
    int NB, Cin, H, W, kh, kw, sh_, sw_, ph, pw, dh, dw, Ho, Wo;
    long long rows;
    int K, ldk;
    long long sb, sh, sw, sc;      // element strides of x (NHWC: H*W*C, W*C, C, 1; an NCHW input works too)
};

template <int VEC>
__global__ void __launch_bounds__(256)
im2col_nhwc_kernel(const __grid_constant__ Im2colNhwcParams p) {
    __shared__ long long s_base[kRowsPerBlock];
    __shared__ int s_h0[kRowsPerBlock], s_w0[kRowsPerBlock];
    const long long row0 = (long long)blockIdx.x * kRowsPerBlock;
    if (threadIdx.x < kRowsPerBlock) {
        const long long row = row0 + threadIdx.x;
        if (row < p.rows) {
            const int L = p.Ho * p.Wo;
            const long long b = row / L;
            const int l = (int)(row - b * L);
            const int ho = l / p.Wo, wo = l - ho * p.Wo;
            s_base[threadIdx.x] = b * p.sb;
            s_h0[threadIdx.x] = ho * p.sh_ - p.ph;
            s_w0[threadIdx.x] = wo * p.sw_ - p.pw;
        } else {
            s_base[threadIdx.x] = -1;
        }
    }
    __syncthreads();
    const int groups = p.ldk / VEC;                 // ldk % VEC == 0 by construction
    for (int g = threadIdx.x; g < groups; g += blockDim.x) {
        const int col = g * VEC;
        const bool pad = col >= p.K;
        const int tap = pad ? 0 : col / p.Cin;
        const int cin = pad ? 0 : col - tap * p.Cin;
        const int r = tap / p.kw, s = tap - r * p.kw;
        const int dh = r * p.dh, dw = s * p.dw;
#pragma unroll
        for (int i = 0; i < kRowsPerBlock; ++i) {
            const long long base = s_base[i];
            if (base < 0) break;
            const int h = s_h0[i] + dh, w = s_w0[i] + dw;
            const bool in = !pad && h >= 0 && h < p.H && w >= 0 && w < p.W;
            float* dst = p.out + (row0 + i) * p.ldk + col;
            if (VEC == 4) {
                float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
                if (in) v = __ldg(reinterpret_cast<const float4*>(p.x + base + h * p.sh + w * p.sw + cin));
                *reinterpret_cast<float4*>(dst) = v;
            } else {
                float v = 0.f;
                if (in) v = __ldg(p.x + base + h * p.sh + w * p.sw + cin * p.sc);
                *dst = v;
            }
        }
    }
}

// (A float4-store variant for scalar-channel wide rows -- the 7x7 stem, Cin = 3, K = 147 -- with the tap offsets in a
// shared-memory table measured SLOWER than the element-per-thread kernel above: 280 vs 221 us; removed.)
// Narrow rows (ldk <= 32: the stem, K = 3 * 3 * 3 = 27 -> 28): a lane is a column, a warp walks 32 rows at a time.
// Each lane decodes ONE of those rows (the divisions), the decoded (image base, h0, w0) travel by shuffle, and every
// store instruction writes one whole row contiguously.  The general kernel keeps 28 of its 256 threads busy on this
// shape: 222 us for 92 MB of output (stem, 3200 x 32 x 32 inputs); this form is bound by the stores.
__global__ void __launch_bounds__(256)
im2col_nhwc_narrow_kernel(const __grid_constant__ Im2colNhwcParams p) {
    const int lane = threadIdx.x & 31;
    const long long warp = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const long long nwarps = ((long long)gridDim.x * blockDim.x) >> 5;
    const int col = lane;
    const bool live = col < p.ldk;
    const bool pad = col >= p.K;
    const int tap = pad ? 0 : col / p.Cin;
    const int cin = pad ? 0 : col - tap * p.Cin;
    const int r = tap / p.kw, sq = tap - r * p.kw;
    const int dh = r * p.dh, dw = sq * p.dw;
    const int L = p.Ho * p.Wo;
    for (long long row_base = warp * 32; row_base < p.rows; row_base += nwarps * 32) {
        const long long row = row_base + lane;
        long long base = -1;
        int h0 = 0, w0 = 0;
        if (row < p.rows) {
            const long long b = row / L;
            const int l = (int)(row - b * L);
            const int ho = l / p.Wo, wo = l - ho * p.Wo;
            base = b * p.sb;
            h0 = ho * p.sh_ - p.ph;
            w0 = wo * p.sw_ - p.pw;
        }
        const int n_here = (int)min((long long)32, p.rows - row_base);
#pragma unroll 4
        for (int i = 0; i < n_here; ++i) {
            const long long bi = __shfl_sync(0xffffffffu, base, i);
            const int h = __shfl_sync(0xffffffffu, h0, i) + dh, w = __shfl_sync(0xffffffffu, w0, i) + dw;
            if (live) {
                float v = 0.f;
                if (!pad && h >= 0 && h < p.H && w >= 0 && w < p.W) v = __ldg(p.x + bi + h * p.sh + w * p.sw + cin * p.sc);
                p.out[(row_base + i) * p.ldk + col] = v;
            }
        }
    }
}

extern "C" int bl_im2col_nhwc(const float* x, float* out, int NB, int Cin, int H, int W, int kh, int kw, int sh,
                              int sw, int ph, int pw, int dh, int dw, int Ho, int Wo, int ldk, long long xsb,
                              long long xsh, long long xsw, long long xsc, void* stream) {
    Im2colNhwcParams p{x, out, NB, Cin, H, W, kh, kw, sh, sw, ph, pw, dh, dw, Ho, Wo, (long long)NB * Ho * Wo,
                       Cin * kh * kw, ldk, xsb, xsh, xsw, xsc};
    if (p.rows <= 0) return 0;
    if (ldk < p.K) return -1;
    const long long blocks = (p.rows + kRowsPerBlock - 1) / kRowsPerBlock;
    if (blocks > 0x7fffffffLL) return -1;
    const bool vec = (Cin % 4 == 0) && (ldk % 4 == 0) && (((uintptr_t)x) % 16 == 0) && (((uintptr_t)out) % 16 == 0) &&
                     xsc == 1 && xsw % 4 == 0 && xsh % 4 == 0 && xsb % 4 == 0;
    if (vec) im2col_nhwc_kernel<4><<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(p);
    else if (ldk <= 32) {
        long long g = (p.rows + 255) / 256;              // 8 warps x 32 rows per block and pass
        if (g > 148LL * 16) g = 148LL * 16;
        im2col_nhwc_narrow_kernel<<<(unsigned)g, 256, 0, (cudaStream_t)stream>>>(p);
    } else im2col_nhwc_kernel<1><<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(p);
    return (int)cudaGetLastError();
}
