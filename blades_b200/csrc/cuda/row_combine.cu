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
};

template <bool VEC>
__global__ void __launch_bounds__(256)
row_combine_kernel(const __grid_constant__ CombineParams p) {
    const long long stride = (long long)gridDim.x * blockDim.x * (VEC ? 4 : 1);
    for (long long c = p.c0 + ((long long)blockIdx.x * blockDim.x + threadIdx.x) * (VEC ? 4 : 1);
         c < p.c1; c += stride) {
        if (VEC) {
            float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
            int i = 0;
            for (; i + 8 <= p.n_rows; i += 8) {
                float4 x[8];
#pragma unroll
                for (int k = 0; k < 8; ++k) x[k] = bl_ldg_stream4(p.rows[i + k] + c);
#pragma unroll
                for (int k = 0; k < 8; ++k) {
                    const float w = p.w[i + k];
                    acc.x = fmaf(w, bl_sanitize(x[k].x), acc.x);
                    acc.y = fmaf(w, bl_sanitize(x[k].y), acc.y);
                    acc.z = fmaf(w, bl_sanitize(x[k].z), acc.z);
                    acc.w = fmaf(w, bl_sanitize(x[k].w), acc.w);
                }
            }
            for (; i < p.n_rows; ++i) {
                const float4 x = bl_ldg_stream4(p.rows[i] + c);
                const float w = p.w[i];
                acc.x = fmaf(w, bl_sanitize(x.x), acc.x);
                acc.y = fmaf(w, bl_sanitize(x.y), acc.y);
                acc.z = fmaf(w, bl_sanitize(x.z), acc.z);
                acc.w = fmaf(w, bl_sanitize(x.w), acc.w);
            }
            // c0 % 4 == 0 and (c1 - c0) % 4 == 0 are guaranteed by the launcher for VEC
            bl_epilogue_store(p.ep, c + 0, acc.x);
            bl_epilogue_store(p.ep, c + 1, acc.y);
            bl_epilogue_store(p.ep, c + 2, acc.z);
            bl_epilogue_store(p.ep, c + 3, acc.w);
        } else {
            float acc = 0.f;
            for (int i = 0; i < p.n_rows; ++i)
                acc = fmaf(p.w[i], bl_sanitize(bl_ldg_stream(p.rows[i] + c)), acc);
            bl_epilogue_store(p.ep, c, acc);
        }
    }
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
        q.c1 = p->c0 + (cols / 4) * 4;
        long long work = (q.c1 - q.c0) / 4;
        unsigned grid = (unsigned)((work + 255) / 256);
        if (grid > cap) grid = cap;
        row_combine_kernel<true><<<grid, 256, 0, st>>>(q);
        if (q.c1 < p->c1) {
            CombineParams t = *p;
            t.c0 = q.c1;
            row_combine_kernel<false><<<1, 32, 0, st>>>(t);
        }
    } else {
        unsigned grid = (unsigned)((cols + 255) / 256);
        if (grid > cap) grid = cap;
        row_combine_kernel<false><<<grid, 256, 0, st>>>(*p);
    }
    return (int)cudaGetLastError();
}

extern "C" int bl_sizeof_combine_params() { return (int)sizeof(CombineParams); }
