// Micro-benchmark: compare-exchange formulations for the register sorting networks (SortNet<40> body, 40 rows).
//   V=0  lo = min, hi = max                      (2 ALU-pipe FMNMX)
//   V=1  lo = min, hi = a + b - lo as two IMADs  (1 ALU + 2 FMA-pipe, exact integer arithmetic on the bit patterns)
//   V=2  two of three comparators in the IMAD form, V=3 one of two, V=4 one of three
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o build/ce_bench scripts/exp/ce_bench.cu
#include <cuda_runtime.h>
#include <cstdio>
#include <vector>
struct P { const float* rows[40]; float* out; int one, mone; };
#define CE_PLAIN(a, b) { float lo_ = fminf(v[a], v[b]); v[b] = fmaxf(v[a], v[b]); v[a] = lo_; }
#define CE_IMAD(a, b) { float lo_ = fminf(v[a], v[b]); int t_, h_; \
   asm("mad.lo.s32 %0, %1, %2, %3;" : "=r"(t_) : "r"(__float_as_int(lo_)), "r"(mone), "r"(__float_as_int(v[b]))); \
   asm("mad.lo.s32 %0, %1, %2, %3;" : "=r"(h_) : "r"(__float_as_int(v[a])), "r"(one), "r"(t_)); \
   v[b] = __int_as_float(h_); v[a] = lo_; }
template <int V>
__global__ void __launch_bounds__(256) k(const __grid_constant__ P p) {
    const int one = p.one, mone = p.mone;
    float v[40];
    unsigned c = blockIdx.x * blockDim.x + threadIdx.x;
#pragma unroll
    for (int i = 0; i < 40; ++i) v[i] = __ldcs(p.rows[i] + c);
#define CE(a, b) { constexpr int K_ = __COUNTER__; \
    if constexpr (V == 1 || (V == 2 && K_ % 3 != 0) || (V == 3 && K_ % 2 == 0) || (V == 4 && K_ % 3 == 0)) CE_IMAD(a, b) else CE_PLAIN(a, b) }
#include "body40.inc"
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < 40; ++i) s = fmaf(s, 1.0001f, v[i]);      // keep every output live, order-sensitive
    p.out[c] = s;
}
template <int V> float run(const P& p, unsigned n, int iters) {
    cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
    float best = 1e9f;
    for (int it = 0; it < iters + 2; ++it) {
        cudaEventRecord(a); k<V><<<n / 256, 256>>>(p); cudaEventRecord(b); cudaEventSynchronize(b);
        float ms; cudaEventElapsedTime(&ms, a, b); if (it >= 2 && ms < best) best = ms;
    }
    return best;
}
int main() {
    const unsigned n = 16u << 20;
    P p; float* buf; cudaMalloc(&buf, (size_t)41 * n * 4);
    std::vector<float> h((size_t)n);
    unsigned s = 12345u;
    for (int r = 0; r < 40; ++r) {
        for (size_t i = 0; i < n; ++i) { s = s * 1664525u + 1013904223u; h[i] = ((int)(s >> 8) - (1 << 23)) * 1e-7f; }
        cudaMemcpy(buf + (size_t)r * n, h.data(), (size_t)n * 4, cudaMemcpyHostToDevice);
        p.rows[r] = buf + (size_t)r * n;
    }
    p.out = buf + (size_t)40 * n; p.one = 1; p.mone = -1;
    std::vector<float> ref(n), got(n);
    float t0 = run<0>(p, n, 5); cudaMemcpy(ref.data(), p.out, (size_t)n * 4, cudaMemcpyDeviceToHost);
    printf("V=0 plain        %.3f ms  %.0f GB/s\n", t0, 40.0 * n * 4 / t0 / 1e6);
#define RUN(V, name) { float t = run<V>(p, n, 5); cudaMemcpy(got.data(), p.out, (size_t)n * 4, cudaMemcpyDeviceToHost); \
    size_t bad = 0; for (size_t i = 0; i < n; ++i) bad += (got[i] != ref[i]); \
    printf("V=%d %-12s %.3f ms  %.0f GB/s  mismatches %zu\n", V, name, t, 40.0 * n * 4 / t / 1e6, bad); }
    RUN(1, "imad all") RUN(2, "imad 2/3") RUN(3, "imad 1/2") RUN(4, "imad 1/3")
    printf("%s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
