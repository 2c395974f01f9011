// Zero-copy mini-batch gather: the selected training samples are read by the GPU straight out of the
// PINNED HOST copy of each client's dataset (UVA-mapped, PCIe reads issued by the copy kernel) and land in the
// device staging buffer in batch order.  Replaces host-side batch assembly + pinned staging + cudaMemcpy
// (reference: one small ``.to(device)`` per batch per client, client.py:186): the host only uploads the
// index list (a few KB) per round.
#include "common.cuh"

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

__global__ void __launch_bounds__(256)
gather_samples_kernel(const __grid_constant__ GatherParams p) {
    const long long t = blockIdx.x;
    if (t >= p.total) return;
    const int c = (int)(t / p.per_client);
    const long long s = p.idx[t];
    const float* src = reinterpret_cast<const float*>(p.src_x[c]) + s * p.sample_floats;
    float* dst = p.dst_x + t * p.sample_floats;
    if ((p.sample_floats & 3) == 0 && ((reinterpret_cast<uintptr_t>(src) | reinterpret_cast<uintptr_t>(dst)) & 15) == 0) {
        const float4* s4 = reinterpret_cast<const float4*>(src);
        float4* d4 = reinterpret_cast<float4*>(dst);
        for (int i = threadIdx.x; i < p.sample_floats / 4; i += blockDim.x) d4[i] = s4[i];
    } else {
        for (int i = threadIdx.x; i < p.sample_floats; i += blockDim.x) dst[i] = src[i];
    }
    if (threadIdx.x == 0) p.dst_y[t] = reinterpret_cast<const long long*>(p.src_y[c])[s];
}

extern "C" int bl_gather_samples(const GatherParams* p, void* stream) {
    if (p->total <= 0) return 0;
    if (p->total > 0x7fffffffLL) return -1;
    gather_samples_kernel<<<(unsigned)p->total, 256, 0, (cudaStream_t)stream>>>(*p);
    return (int)cudaGetLastError();
}
extern "C" int bl_sizeof_gather_params() { return (int)sizeof(GatherParams); }
