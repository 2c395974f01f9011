// In-fabric all-reduce of a small symmetric buffer -- the cross-GPU sum of the Gram partials (reference krum.py:85-90
// computes its pairwise distances on the driver from the gathered updates; here every GPU holds the partial Gram of
// its coordinate range and the N x N sums meet inside the NVSwitch).
//
// Two-shot, one kernel: rank r owns the r-th slice of the buffer,
//   NVLS:   v = multimem.ld_reduce.add.v4.f32 [mc + i]   (the switch sums the G replicas)
//           multimem.st.v4.f32 [mc + i], v               (and replicates the sum into every GPU's copy)
//   no multicast object: the slice is summed with plain 16 B peer loads and written with one 16 B store per peer.
// The caller brackets the launch with the symmetric-memory device barrier (partials complete / sums landed).
#include "common.cuh"

struct NvlsReduceParams {
    float* peers[BL_MAX_PEERS];    // replica base pointers (peer-mapped), used when mc == nullptr
    float* mc;                     // multicast address of the same buffer (nullptr: P2P path)
    long long count;               // floats, multiple of 4
    int rank, world;
};

__global__ void __launch_bounds__(256)
nvls_allreduce_kernel(const __grid_constant__ NvlsReduceParams p) {
    const long long vecs = p.count / 4;
    const long long per = (vecs + p.world - 1) / p.world;
    const long long v0 = per * p.rank, v1 = min(vecs, v0 + per);
    for (long long i = v0 + (long long)blockIdx.x * blockDim.x + threadIdx.x; i < v1; i += (long long)gridDim.x * blockDim.x) {
        if (p.mc) {
            const float4 v = bl_mc_ld_reduce4(p.mc + 4 * i);
            bl_mc_store4(p.mc + 4 * i, v);
        } else {
            float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
            for (int g = 0; g < BL_MAX_PEERS; ++g)
                if (g < p.world) {
                    const float4 x = bl_ld_volatile4(p.peers[g] + 4 * i);
                    acc.x += x.x; acc.y += x.y; acc.z += x.z; acc.w += x.w;
                }
#pragma unroll
            for (int g = 0; g < BL_MAX_PEERS; ++g)
                if (g < p.world) *reinterpret_cast<float4*>(p.peers[g] + 4 * i) = acc;
        }
    }
}

extern "C" int bl_nvls_allreduce(const NvlsReduceParams* p, void* stream) {
    if (p->count <= 0) return 0;
    if (p->count % 4 != 0 || p->world < 1 || p->world > BL_MAX_PEERS) return -1;
    const long long per = (p->count / 4 + p->world - 1) / p->world;
    unsigned grid = (unsigned)((per + 255) / 256);
    if (grid > 148u * 4) grid = 148u * 4;
    if (grid < 1) grid = 1;
    nvls_allreduce_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(*p);
    return (int)cudaGetLastError();
}

extern "C" int bl_sizeof_nvls_reduce_params() { return (int)sizeof(NvlsReduceParams); }

// ---------------------------------------------------------------------------------------------------------------------
// Strided 2-D copy on the copy engines (cudaMemcpy2DAsync over UVA: local or NVLink peer memory on either side).  The
// push half of the sharded aggregation: as soon as the backward pass has finished a window of update coordinates, every
// rank DMAs its rows' slice of that window to the rank that aggregates it -- on a side stream, by the copy engines, while
// the SMs run the rest of the backward pass -- so the selection kernel later streams LOCAL HBM only.
extern "C" int bl_copy2d_async(void* dst, long long dpitch, const void* src, long long spitch, long long width_bytes,
                               long long height, void* stream) {
    if (width_bytes <= 0 || height <= 0) return 0;
    return (int)cudaMemcpy2DAsync(dst, (size_t)dpitch, src, (size_t)spitch, (size_t)width_bytes, (size_t)height,
                                  cudaMemcpyDefault, (cudaStream_t)stream);
}

// Zero fill by the driver's memset path (copy engine / memset node) -- no elementwise kernel launch.
extern "C" int bl_memset_zero_async(void* dst, long long bytes, void* stream) {
    if (bytes <= 0) return 0;
    return (int)cudaMemsetAsync(dst, 0, (size_t)bytes, (cudaStream_t)stream);
}
