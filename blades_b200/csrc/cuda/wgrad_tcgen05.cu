// Grouped (per-client) weight-gradient GEMM with the client-update epilogue -- K9 of SURVEY 2.7.
//
//     out[c][m][n] = alpha * sum_t  A[c*T + t][m] * B[c*T + t][n]        c = client, t = its T rows
//
// A = output-gradients [n_clients*T, M], B = layer inputs (or im2col rows) [n_clients*T, N], both
// row-major, i.e. "MN-major" UMMA operands: no transposes are materialised.  `out` is a strided window
// into the shard's update matrix U_g (batch stride = d floats), so the GEMM epilogue IS the client's
// SGD step + update diff + nan_to_num + save_update of the reference (client.py:127-131,178-198):
// alpha = -lr.  The problem is output-bound (K = T is 32..8192 while each client writes M*N floats),
// so the kernel is organised around streaming 128x256 fp32 tiles out of TMEM at HBM write speed:
//   warp 0   TMA producer: per stage 4 (A) + BN/32 (B) boxes of [KT rows x 128 B], SWIZZLE_128B_ATOM_32B
//            (the only smem layout tcgen05 accepts for MN-major tf32 operands)
//   warp 1   MMA issuer: tcgen05.mma.kind::tf32, M=128, N=BN, one K=8 atom per instruction,
//            accumulators double-buffered in TMEM (2 x 256 columns)
//   warps 2-9 epilogue (two per TMEM lane quarter): tcgen05.ld -> scale/sanitise -> smem transpose -> coalesced stores
// Persistent CTAs walk (client, m-tile, n-tile) tiles round-robin; smem and TMEM pipelines run
// across tile boundaries so the next tile's loads/MMAs overlap the current tile's stores.
#include "common.cuh"
#include "tc_common.cuh"
#include <cstring>
#include <cstdlib>

struct WgradParams {
    CUtensorMap map_a;       // 2D: (M, n_clients*T), box (32, KT)
    CUtensorMap map_b;       // 2D: (N, n_clients*T), box (32, KT)
    int n_clients, T, M, N;
    int KT;                  // rows per stage (8/16/32, divides T)
    int BN;                  // tile N (multiple of 32, <= 256)
    int m_tiles, n_tiles;
    int stages;
    float alpha;
    float* out;              // out[c*batch_stride + m*N + n]
    long long batch_stride;
    int vec_ok;              // 16 B aligned stores possible
    // ---- implicit-GEMM convolution mode: B rows are gathered from the NHWC activation by TMA (no im2col)
    CUtensorMap map_x;       // 4D: (Cin, W, H, NB), box (32, bw*cs, bh*cs, bb), element strides (1, cs, cs, 1)
    int implicit;
    int Cin, kw, taps;       // N index n = tap * Cin + cin,  tap = r * kw + s
    int cs, cp;              // conv stride / padding
    int Ho, Wo, Bc;          // output spatial size, samples per client
    int bh, bb;              // box: full output rows (Wo) x bh output rows x bb samples  (KT = Wo*bh*bb)
};

namespace {
constexpr int kWThreads = 320;           // TMA, MMA + 8 epilogue warps (2 per TMEM lane quarter)
constexpr int kEpiLd = 36;               // padded row (floats): 16 B aligned, conflict-free float4 access

__global__ void __launch_bounds__(kWThreads, 1)
wgrad_tcgen05_kernel(const __grid_constant__ WgradParams p) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    const uint32_t box_bytes = (uint32_t)p.KT * 128u;
    const uint32_t a_bytes = 4u * box_bytes;
    const uint32_t nb_blocks = (uint32_t)p.BN / 32u;
    const uint32_t stage_bytes = a_bytes + nb_blocks * box_bytes;
    uint8_t* tiles = smem_raw;
    uint64_t* bars = reinterpret_cast<uint64_t*>(tiles + (size_t)p.stages * stage_bytes);
    uint64_t* full = bars;
    uint64_t* empty = bars + p.stages;
    uint64_t* tfull = bars + 2 * p.stages;        // [2] accumulator ready
    uint64_t* tempty = bars + 2 * p.stages + 2;   // [2] accumulator drained
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * p.stages + 4);
    float* epi = reinterpret_cast<float*>(tiles + (size_t)p.stages * stage_bytes + 256);   // 4 x [32][kEpiLd]

    const int warp = bl::uniform_warp_idx(), lane = threadIdx.x & 31;
    const long long total_tiles = (long long)p.n_clients * p.m_tiles * p.n_tiles;
    const int ksteps = p.T / p.KT;

    if (warp == 0 && lane == 0) {
        bl::tma_prefetch_desc(&p.map_a);
        bl::tma_prefetch_desc(&p.map_b);
        for (int s = 0; s < p.stages; ++s) { bl::mbar_init(&full[s], 1); bl::mbar_init(&empty[s], 1); }
        for (int i = 0; i < 2; ++i) { bl::mbar_init(&tfull[i], 1); bl::mbar_init(&tempty[i], 8); }
        bl::fence_barrier_init();
    }
    if (warp == 1) bl::tmem_alloc<512>(tmem_slot);
    bl::tc_fence_before();
    __syncthreads();
    bl::tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    const uint32_t tiles0 = bl::smem_u32(tiles);
    const uint32_t full0 = bl::smem_u32(full), empty0 = bl::smem_u32(empty);
    const uint32_t tfull0 = bl::smem_u32(tfull), tempty0 = bl::smem_u32(tempty);

    // Warps 0 and 1 run their loops CONVERGED (all lanes, uniform values, election inside the asm): see tc_common.cuh
    // "warp-uniform issue".  Stage / parity are running counters (no integer divisions in the issue loops).
    if (warp == 0) {
        // ================= TMA producer =================
        int s = 0;
        uint32_t ph = 0;
        const int hchunks = p.implicit ? p.Ho / p.bh : 1;
        for (long long tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
            const int nt = (int)(tile % p.n_tiles);
            const int mt = (int)((tile / p.n_tiles) % p.m_tiles);
            const int c = (int)(tile / ((long long)p.n_tiles * p.m_tiles));
            int kb = 0, kh = 0;         // implicit mode: sample-group / row-chunk counters of the K step
            for (int ks = 0; ks < ksteps; ++ks) {
                bl::mbar_wait_u32(empty0 + 8u * s, ph ^ 1u);
                const uint32_t fb = full0 + 8u * s;
                bl::mbar_arrive_expect_tx_e(fb, stage_bytes);
                const uint32_t dst = tiles0 + (uint32_t)s * stage_bytes;
                const int row = c * p.T + ks * p.KT;
#pragma unroll
                for (int j = 0; j < 4; ++j)
                    bl::tma_load_2d_e(dst + j * box_bytes, &p.map_a, fb, mt * 128 + j * 32, row);
                if (!p.implicit) {
                    for (uint32_t j = 0; j < nb_blocks; ++j)
                        bl::tma_load_2d_e(dst + a_bytes + j * box_bytes, &p.map_b, fb, nt * p.BN + (int)j * 32, row);
                } else {
                    // K chunk ks = output positions (b0..b0+bb) x (ho0..ho0+bh) x (0..Wo) of client c; for the
                    // 32-channel block (tap, cin0) the matching input pixels are one strided 4-D TMA box; conv
                    // padding = out-of-range coordinates, zero-filled by TMA.
                    const int b0 = kb * p.bb, ho0 = kh * p.bh;
                    int n0 = nt * p.BN;
                    int tap = n0 / p.Cin, cin0 = n0 - tap * p.Cin;          // one division per K step, then running
                    int r = tap / p.kw, sx = tap - r * p.kw;
                    for (uint32_t j = 0; j < nb_blocks; ++j) {
                        const bool valid = tap < p.taps;
                        bl::tma_load_4d_e(dst + a_bytes + j * box_bytes, &p.map_x, fb,
                                          valid ? cin0 : p.Cin,           // beyond the channel extent -> zeros
                                          sx - p.cp, ho0 * p.cs + r - p.cp, c * p.Bc + b0);
                        cin0 += 32;
                        if (cin0 >= p.Cin) { cin0 = 0; ++tap; if (++sx == p.kw) { sx = 0; ++r; } }
                    }
                    if (++kh == hchunks) { kh = 0; ++kb; }
                }
                if (++s == p.stages) { s = 0; ph ^= 1u; }
            }
        }
    } else if (warp == 1) {
        // ================= MMA issuer =================
        const uint32_t idesc = bl::umma_idesc_tf32(128, (uint32_t)p.BN, 1, 1);
        uint32_t tcount = 0;
        int s = 0;
        uint32_t ph = 0;
        // MN-major tf32: SWIZZLE_128B_BASE32B; LBO = stride between 32-float column blocks, SBO = stride between 4-row
        // K atoms (rows are contiguous: 512 B); one 8-row K atom (1024 B) per MMA.  The start-address field of the
        // descriptors (address >> 4) advances by plain adds.
        const uint64_t ad0 = bl::umma_smem_desc(tiles0, box_bytes, 512u, bl::kLayoutSw128Base32B);
        const uint64_t bd0 = bl::umma_smem_desc(tiles0 + a_bytes, box_bytes, 512u, bl::kLayoutSw128Base32B);
        for (long long tile = blockIdx.x; tile < total_tiles; tile += gridDim.x, ++tcount) {
            const uint32_t buf = tcount & 1u;
            const uint32_t tph = (tcount >> 1) & 1u;
            bl::mbar_wait_u32(tempty0 + 8u * buf, tph ^ 1u);             // epilogue drained this accumulator
            bl::tc_fence_after();
            const uint32_t d_tmem = tmem_base + buf * 256u;
            for (int ks = 0; ks < ksteps; ++ks) {
                bl::mbar_wait_u32(full0 + 8u * s, ph);
                bl::tc_fence_after();
                const uint64_t so = (uint64_t)(((uint32_t)s * stage_bytes) >> 4);
                for (int ka = 0; ka < p.KT / 8; ++ka)
                    bl::umma_tf32_e(d_tmem, ad0 + so + (uint64_t)(ka * 64), bd0 + so + (uint64_t)(ka * 64), idesc,
                                    (ks > 0 || ka > 0) ? 1u : 0u);
                bl::umma_commit_e(empty0 + 8u * s);
                if (ks == ksteps - 1) bl::umma_commit_e(tfull0 + 8u * buf);
                if (++s == p.stages) { s = 0; ph ^= 1u; }
            }
        }
    } else {
        // ================= epilogue =================
        const int q = warp & 3;                  // TMEM lane quarter this warp may access
        const int half = (warp - 2) >> 2;        // the two warps of a quarter split the column chunks
        uint32_t tcount = 0;
        for (long long tile = blockIdx.x; tile < total_tiles; tile += gridDim.x, ++tcount) {
            const int nt = (int)(tile % p.n_tiles);
            const int mt = (int)((tile / p.n_tiles) % p.m_tiles);
            const int c = (int)(tile / ((long long)p.n_tiles * p.m_tiles));
            const uint32_t buf = tcount & 1u;
            const uint32_t tph = (tcount >> 1) & 1u;
            bl::mbar_wait(&tfull[buf], tph);
            bl::tc_fence_after();
            // TMEM -> registers (lane = row) -> per-warp smem transpose -> coalesced stores: every STG.128
            // covers 4 rows x 128 contiguous bytes (a lane-per-row store would touch 32 rows x 16 B).
            float* stg = epi + (warp - 2) * (32 * kEpiLd);
            const int rsub = lane >> 3, csub = (lane & 7) * 4;
            float* obase = p.out + (long long)c * p.batch_stride;
            for (int cb = half * 32; cb < p.BN; cb += 64) {
                float v[32];
                bl::tmem_ld_32x32(tmem_base + ((uint32_t)(q * 32) << 16) + buf * 256u + (uint32_t)cb, v);
                const int n0 = nt * p.BN + cb;
                if (n0 >= p.N) break;
#pragma unroll
                for (int j = 0; j < 32; j += 4)
                    *reinterpret_cast<float4*>(stg + lane * kEpiLd + j) =
                        make_float4(bl_sanitize(p.alpha * v[j]), bl_sanitize(p.alpha * v[j + 1]),
                                    bl_sanitize(p.alpha * v[j + 2]), bl_sanitize(p.alpha * v[j + 3]));
                __syncwarp();
#pragma unroll
                for (int r0 = 0; r0 < 32; r0 += 4) {
                    const int r = r0 + rsub;
                    const int m = mt * 128 + q * 32 + r;
                    const float4 t = *reinterpret_cast<const float4*>(stg + r * kEpiLd + csub);
                    if (m < p.M) {
                        float* dst = obase + (long long)m * p.N + n0 + csub;
                        if (p.vec_ok && n0 + csub + 4 <= p.N) {
                            *reinterpret_cast<float4*>(dst) = t;
                        } else {
                            if (n0 + csub + 0 < p.N) dst[0] = t.x;
                            if (n0 + csub + 1 < p.N) dst[1] = t.y;
                            if (n0 + csub + 2 < p.N) dst[2] = t.z;
                            if (n0 + csub + 3 < p.N) dst[3] = t.w;
                        }
                    }
                }
                __syncwarp();
            }
            bl::tc_fence_before();
            __syncwarp();
            if (lane == 0) bl::mbar_arrive(&tempty[buf]);
        }
    }
    bl::tc_fence_before();
    __syncthreads();
    if (warp == 1) bl::tmem_dealloc(tmem_base, 512);
}
}  // namespace

extern "C" int bl_grouped_wgrad(const float* a, const float* b, float* out, int n_clients, int T, int M, int N,
                                long long lda, long long ldb, long long batch_stride, float alpha, int num_sms,
                                void* stream) {
    // a: [n*T rows][lda] (M valid columns), b: [n*T rows][ldb] (N valid columns)
    if (lda % 4 != 0 || ldb % 4 != 0 || T % 8 != 0 || lda < M || ldb < N) return -1;   // TMA stride / box constraints
    if (((uintptr_t)a) % 16 != 0 || ((uintptr_t)b) % 16 != 0) return -1;
    WgradParams p;
    memset(&p, 0, sizeof(p));
    p.n_clients = n_clients; p.T = T; p.M = M; p.N = N;
    p.KT = (T % 32 == 0) ? 32 : (T % 16 == 0 ? 16 : 8);
    int bn = (N + 31) / 32 * 32;
    if (bn > 256) bn = 256;
    p.BN = bn;
    p.m_tiles = (M + 127) / 128;
    p.n_tiles = (N + bn - 1) / bn;
    p.alpha = alpha;
    p.out = out;
    p.batch_stride = batch_stride;
    p.vec_ok = (((uintptr_t)out) % 16 == 0) && (batch_stride % 4 == 0) && (N % 4 == 0);
    const uint64_t rows = (uint64_t)n_clients * T;
    {
        uint64_t dims[2] = {(uint64_t)M, rows};
        uint64_t strides[1] = {(uint64_t)lda * 4};
        uint32_t box[2] = {32, (uint32_t)p.KT};
        int r = bl::make_tmap_f32(&p.map_a, a, 2, dims, strides, box, CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B);
        if (r != 0) return 1000 + r;
    }
    {
        uint64_t dims[2] = {(uint64_t)N, rows};
        uint64_t strides[1] = {(uint64_t)ldb * 4};
        uint32_t box[2] = {32, (uint32_t)p.KT};
        int r = bl::make_tmap_f32(&p.map_b, b, 2, dims, strides, box, CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B);
        if (r != 0) return 1000 + r;
    }
    const size_t stage_bytes = (size_t)(4 + bn / 32) * p.KT * 128;
    int stages = (int)((180 * 1024) / stage_bytes);
    if (stages > 8) stages = 8;
    if (stages < 2) return -2;
    p.stages = stages;
    const size_t smem = stages * stage_bytes + 256 + 8 * 32 * kEpiLd * sizeof(float);
    static bool attr_done = false;       // opt in to the full 227 KB once (not a stream op: keep it out of graph capture)
    if (!attr_done) {
        cudaError_t e = cudaFuncSetAttribute(wgrad_tcgen05_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
        if (e != cudaSuccess) return (int)e;
        attr_done = true;
    }
    const long long total = (long long)n_clients * p.m_tiles * p.n_tiles;
    int grid = num_sms > 0 ? num_sms : 148;
    if ((long long)grid > total) grid = (int)total;
    wgrad_tcgen05_kernel<<<grid, kWThreads, smem, (cudaStream_t)stream>>>(p);
    return (int)cudaGetLastError();
}


// Implicit-GEMM per-client conv weight gradient: gy [NB*Ho*Wo, Cout] (NHWC rows), x [NB, H, W, Cin] (NHWC),
// out[c][Cout][kh*kw*Cin] -- the im2col matrix is never materialised.
extern "C" int bl_conv_wgrad_implicit(const float* gy, const float* x, float* out, int n_clients, int Bc, int H,
                                      int W, int Cin, int Ho, int Wo, int Cout, int kh, int kw, int cs, int cp,
                                      long long batch_stride, float alpha, int num_sms, void* stream) {
    if (Cin % 32 != 0 || Cout % 4 != 0 || cs < 1 || cs > 8 || Wo > 32 || Wo * cs > 256) return -1;
    if (((uintptr_t)gy) % 16 != 0 || ((uintptr_t)x) % 16 != 0) return -1;
    // K chunk of ~32 output positions: whole output rows, then whole images
    int bh = 32 / Wo; if (bh < 1) bh = 1; if (bh > Ho) bh = Ho;
    while (Ho % bh != 0) --bh;
    int bb = 1;
    if (bh == Ho) { bb = 32 / (Wo * Ho); if (bb < 1) bb = 1; if (bb > Bc) bb = Bc; while (Bc % bb != 0) --bb; }
    const int KT = Wo * bh * bb;
    if (KT % 8 != 0 || KT > 256 || bh * cs > 256) return -1;
    WgradParams p;
    memset(&p, 0, sizeof(p));
    const int T = Bc * Ho * Wo, M = Cout, N = kh * kw * Cin;
    p.n_clients = n_clients; p.T = T; p.M = M; p.N = N; p.KT = KT;
    int bn = (N + 31) / 32 * 32; if (bn > 256) bn = 256;
    p.BN = bn; p.m_tiles = (M + 127) / 128; p.n_tiles = (N + bn - 1) / bn;
    p.alpha = alpha; p.out = out; p.batch_stride = batch_stride;
    p.vec_ok = (((uintptr_t)out) % 16 == 0) && (batch_stride % 4 == 0) && (N % 4 == 0);
    p.implicit = 1; p.Cin = Cin; p.kw = kw; p.taps = kh * kw; p.cs = cs; p.cp = cp; p.Ho = Ho; p.Wo = Wo; p.Bc = Bc;
    p.bh = bh; p.bb = bb;
    const uint64_t rows = (uint64_t)n_clients * T;
    {
        uint64_t dims[2] = {(uint64_t)M, rows};
        uint64_t strides[1] = {(uint64_t)M * 4};
        uint32_t box[2] = {32, (uint32_t)KT};
        int r = bl::make_tmap_f32(&p.map_a, gy, 2, dims, strides, box, CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B);
        if (r != 0) return 1000 + r;
    }
    {
        uint64_t dims[4] = {(uint64_t)Cin, (uint64_t)W, (uint64_t)H, (uint64_t)n_clients * Bc};
        uint64_t strides[3] = {(uint64_t)Cin * 4, (uint64_t)W * Cin * 4, (uint64_t)H * W * Cin * 4};
        uint32_t box[4] = {32, (uint32_t)(Wo * cs), (uint32_t)(bh * cs), (uint32_t)bb};
        uint32_t es[4] = {1, (uint32_t)cs, (uint32_t)cs, 1};
        int r = bl::make_tmap_f32(&p.map_x, x, 4, dims, strides, box, CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B, es);
        if (r != 0) return 2000 + r;
    }
    const size_t stage_bytes = (size_t)(4 + bn / 32) * p.KT * 128;
    int stages = (int)((180 * 1024) / stage_bytes);
    if (stages > 8) stages = 8;
    if (stages < 2) return -2;
    p.stages = stages;
    const size_t smem = stages * stage_bytes + 256 + 8 * 32 * kEpiLd * sizeof(float);
    static bool attr_done = false;
    if (!attr_done) {
        cudaError_t e = cudaFuncSetAttribute(wgrad_tcgen05_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
        if (e != cudaSuccess) return (int)e;
        attr_done = true;
    }
    const long long total = (long long)n_clients * p.m_tiles * p.n_tiles;
    int grid = num_sms > 0 ? num_sms : 148;
    if ((long long)grid > total) grid = (int)total;
    wgrad_tcgen05_kernel<<<grid, kWThreads, smem, (cudaStream_t)stream>>>(p);
    return (int)cudaGetLastError();
}
