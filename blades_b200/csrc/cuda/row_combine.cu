// Weighted row combine  out[c] = sum_i w_i * U[i][c]  -- K2 of SURVEY 2.7 (mean.py:72,
// geomed.py:78, fltrust.py:37, centeredclipping.py:41-44, clustering.py:42) with the K8 epilogue.
// Pull mode: each launch owns a coordinate range and reads that range of every row, local or
// peer (NVLink) -- the "all-gather" never materialises.  Memory-bound streaming kernel: float4
// loads, 8 rows in flight per thread, fp32 FMA accumulation.
#include "common.cuh"

struct CombineParams {
    const float* rows[BL_MAX_ROWS + 1];
    float w[BL_MAX_ROWS + 1];
    int n_rows;
    long long c0, c1;
    BlEpilogue ep;
    int ep_vec;               // set by the launcher: every output replica is 16 B aligned -> 16 B / v4 multicast stores
    const float* w_dev;       // weights in device memory (written by the on-device Gram solvers, gram_solve.cu):
                              // overrides w[]; rows whose weight is 0 are never read (Krum selects 1 of N rows)
};

// Device-side weights: warp 0 of every CTA compacts the rows with a non-zero weight into shared memory, in row order
// (deterministic summation order), so the streaming loop below only touches rows that contribute.
struct CombineList {
    const float* rows[BL_MAX_ROWS + 1];
    float w[BL_MAX_ROWS + 1];
    int n;
};

__device__ __forceinline__ void combine_compact(const CombineParams& p, CombineList& L) {
    if (threadIdx.x < 32) {
        const unsigned lane = threadIdx.x;
        int off = 0;
        for (int base = 0; base < p.n_rows; base += 32) {
            const int i = base + (int)lane;
            const float w = i < p.n_rows ? p.w_dev[i] : 0.f;
            const unsigned m = __ballot_sync(0xffffffffu, w != 0.f);
            if (w != 0.f) {
                const int slot = off + __popc(m & ((1u << lane) - 1u));
                L.rows[slot] = p.rows[i];
                L.w[slot] = w;
            }
            off += __popc(m);
        }
        if (lane == 0) L.n = off;
    }
    __syncthreads();
}

template <bool VEC, bool DEV>
__global__ void __launch_bounds__(256)
row_combine_kernel(const __grid_constant__ CombineParams p) {
    __shared__ CombineList L;
    if (DEV) combine_compact(p, L);
    const int n_rows = DEV ? L.n : p.n_rows;
#define ROW(i) (DEV ? L.rows[i] : p.rows[i])
#define WGT(i) (DEV ? L.w[i] : p.w[i])
    const long long stride = (long long)gridDim.x * blockDim.x * (VEC ? 4 : 1);
    for (long long c = p.c0 + ((long long)blockIdx.x * blockDim.x + threadIdx.x) * (VEC ? 4 : 1);
         c < p.c1; c += stride) {
        if (VEC) {
            float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
            int i = 0;
            for (; i + 8 <= n_rows; i += 8) {
                float4 x[8];
#pragma unroll
                for (int k = 0; k < 8; ++k) x[k] = bl_ldg_stream4(ROW(i + k) + c);
#pragma unroll
                for (int k = 0; k < 8; ++k) {
                    const float w = WGT(i + k);
                    acc.x = fmaf(w, bl_sanitize(x[k].x), acc.x);
                    acc.y = fmaf(w, bl_sanitize(x[k].y), acc.y);
                    acc.z = fmaf(w, bl_sanitize(x[k].z), acc.z);
                    acc.w = fmaf(w, bl_sanitize(x[k].w), acc.w);
                }
            }
            for (; i < n_rows; ++i) {
                const float4 x = bl_ldg_stream4(ROW(i) + c);
                const float w = WGT(i);
                acc.x = fmaf(w, bl_sanitize(x.x), acc.x);
                acc.y = fmaf(w, bl_sanitize(x.y), acc.y);
                acc.z = fmaf(w, bl_sanitize(x.z), acc.z);
                acc.w = fmaf(w, bl_sanitize(x.w), acc.w);
            }
            // c0 % 4 == 0 and (c1 - c0) % 4 == 0 are guaranteed by the launcher for VEC
            if (p.ep_vec) {
                bl_epilogue_store4(p.ep, c, acc);
            } else {
                bl_epilogue_store(p.ep, c + 0, acc.x);
                bl_epilogue_store(p.ep, c + 1, acc.y);
                bl_epilogue_store(p.ep, c + 2, acc.z);
                bl_epilogue_store(p.ep, c + 3, acc.w);
            }
        } else {
            float acc = 0.f;
            for (int i = 0; i < n_rows; ++i)
                acc = fmaf(WGT(i), bl_sanitize(bl_ldg_stream(ROW(i) + c)), acc);
            bl_epilogue_store(p.ep, c, acc);
        }
    }
#undef ROW
#undef WGT
}

template <bool VEC>
static void launch_combine_kernel(const CombineParams& p, unsigned grid, unsigned block, cudaStream_t st) {
    if (p.w_dev) row_combine_kernel<VEC, true><<<grid, block, 0, st>>>(p);
    else row_combine_kernel<VEC, false><<<grid, block, 0, st>>>(p);
}

extern "C" int bl_row_combine(const CombineParams* p, int num_sms, void* stream) {
    cudaStream_t st = (cudaStream_t)stream;
    if (p->n_rows < 1 || p->n_rows > BL_MAX_ROWS + 1) return -1;
    long long cols = p->c1 - p->c0;
    if (cols <= 0) return 0;
    bool vec = (p->c0 % 4 == 0);
    for (int i = 0; i < p->n_rows && vec; ++i) vec = (((uintptr_t)p->rows[i]) % 16 == 0);
    const unsigned cap = (unsigned)(num_sms > 0 ? num_sms * 8 : 148 * 8);
    if (vec && cols >= 4) {
        CombineParams q = *p;
        bool ev = ((uintptr_t)p->ep.mc_out % 16 == 0) && ((uintptr_t)p->ep.mc_theta % 16 == 0) && ((uintptr_t)p->ep.theta_src % 16 == 0);
        for (int g = 0; g < p->ep.n_out && ev; ++g) ev = ((uintptr_t)p->ep.out[g] % 16 == 0);
        for (int g = 0; g < p->ep.n_theta && ev; ++g) ev = ((uintptr_t)p->ep.theta[g] % 16 == 0);
        q.ep_vec = ev ? 1 : 0;
        q.c1 = p->c0 + (cols / 4) * 4;
        long long work = (q.c1 - q.c0) / 4;
        unsigned grid = (unsigned)((work + 255) / 256);
        if (grid > cap) grid = cap;
        launch_combine_kernel<true>(q, grid, 256, st);
        if (q.c1 < p->c1) {
            CombineParams t = *p;
            t.c0 = q.c1;
            launch_combine_kernel<false>(t, 1, 32, st);
        }
    } else {
        unsigned grid = (unsigned)((cols + 255) / 256);
        if (grid > cap) grid = cap;
        launch_combine_kernel<false>(*p, grid, 256, st);
    }
    return (int)cudaGetLastError();
}

extern "C" int bl_sizeof_combine_params() { return (int)sizeof(CombineParams); }
