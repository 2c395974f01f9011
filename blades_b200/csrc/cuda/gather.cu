// Zero-copy mini-batch gather: the selected training samples are read by the GPU straight out of the
// PINNED HOST copy of each client's dataset (UVA-mapped, PCIe reads issued by the copy kernel) and land in the
// device staging buffer in batch order.  Replaces host-side batch assembly + pinned staging + cudaMemcpy
// (reference: one small ``.to(device)`` per batch per client, client.py:186): the host only uploads the
// index list (a few KB) per round.
#include "common.cuh"
#include <cstdlib>

struct GatherParams {
    const unsigned long long* src_x;   // [n_clients] host-pinned base pointers of the sample arrays
    const unsigned long long* src_y;   // [n_clients] host-pinned base pointers of the int64 label arrays
    const long long* idx;              // [n_clients * per_client] sample index inside the client's array
    float* dst_x;                      // [n_clients * per_client][sample_floats]
    long long* dst_y;
    int per_client;
    int sample_floats;
    long long total;
};

// Persistent and deliberately SMALL: the kernel is PCIe bound (~55 GB/s, 39 MB per headline round = 0.7 ms) and runs
// on a side stream WHILE the previous round trains.  One CTA per sample (3200 x 256 threads) kept every SM's 2048
// thread slots occupied by warps parked on PCIe reads for those 0.7 ms, and the training kernels queued behind them:
// the end-to-end round was 0.75 ms longer than the device-timed one.  Now 64 CTAs of 128 threads loop over the
// samples with 4 independent 16 B loads in flight per thread (~0.5 MB outstanding, far more than PCIe latency x
// bandwidth needs) and leaves the SMs to the round: end-to-end 135 -> 146 rounds/s at 150 device-timed.
constexpr int kGatherThreads = 128;

__global__ void __launch_bounds__(kGatherThreads)
gather_samples_kernel(const __grid_constant__ GatherParams p) {
    for (long long t = blockIdx.x; t < p.total; t += gridDim.x) {
        const int c = (int)(t / p.per_client);
        const long long s = p.idx[t];
        const float* src = reinterpret_cast<const float*>(p.src_x[c]) + s * p.sample_floats;
        float* dst = p.dst_x + t * p.sample_floats;
        if ((p.sample_floats & 3) == 0 && ((reinterpret_cast<uintptr_t>(src) | reinterpret_cast<uintptr_t>(dst)) & 15) == 0) {
            const float4* s4 = reinterpret_cast<const float4*>(src);
            float4* d4 = reinterpret_cast<float4*>(dst);
            const int n4 = p.sample_floats / 4;
            int i = threadIdx.x;
            for (; i + 3 * kGatherThreads < n4; i += 4 * kGatherThreads) {
                const float4 v0 = s4[i], v1 = s4[i + kGatherThreads], v2 = s4[i + 2 * kGatherThreads],
                             v3 = s4[i + 3 * kGatherThreads];
                d4[i] = v0; d4[i + kGatherThreads] = v1; d4[i + 2 * kGatherThreads] = v2; d4[i + 3 * kGatherThreads] = v3;
            }
            for (; i < n4; i += kGatherThreads) d4[i] = s4[i];
        } else {
            for (int i = threadIdx.x; i < p.sample_floats; i += kGatherThreads) dst[i] = src[i];
        }
        if (threadIdx.x == 0) p.dst_y[t] = reinterpret_cast<const long long*>(p.src_y[c])[s];
    }
}

extern "C" int bl_gather_samples(const GatherParams* p, void* stream) {
    if (p->total <= 0) return 0;
    if (p->total > 0x7fffffffLL) return -1;
    static int ctas = 0;
    if (ctas == 0) {
        const char* e = getenv("BLADES_GATHER_CTAS");
        int dev = 0, sms = 148;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
        ctas = e ? atoi(e) : 64;             // measured e2e rounds/s on B200: 64 CTAs 146.1, 148: 143.3, 296: 141.9
        if (ctas < 1) ctas = 64;
        if (ctas > 8 * sms) ctas = 8 * sms;
    }
    const unsigned grid = (unsigned)(p->total < ctas ? p->total : ctas);
    gather_samples_kernel<<<grid, kGatherThreads, 0, (cudaStream_t)stream>>>(*p);
    return (int)cudaGetLastError();
}
extern "C" int bl_sizeof_gather_params() { return (int)sizeof(GatherParams); }
