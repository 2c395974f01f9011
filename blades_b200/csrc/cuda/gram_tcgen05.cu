// Split-K Gram matrix  G = U U^T  on the 5th-gen tensor cores -- K5 of SURVEY 2.7.
// Replaces the reference's O(N^2) Python loops of (v1 - v2).norm() / scipy cosine over CPU vectors
// (krum.py:85-90, clustering.py:28-33, clippedclustering.py:52-57, geomed.py:62,74): every pairwise
// distance / cosine / norm is derived from G on the host (aggregators/_gramops.py).
//
// Shape of the problem: M = N = #clients (<= 512), K = d = 11-24 M coordinates -> memory bound
// (each element of U must stream from HBM / NVLink exactly once).  Design:
//   * persistent split-K: CTA (ks, grp) owns a contiguous range of 32-float K chunks and
//     `mb_per_cta` 128-row M blocks; its fp32 accumulators (128 x NP per M block) live in TMEM for
//     the whole kernel and are flushed once with red.global.add at the end.
//   * operands: one smem tile [tile_rows x 128 B] per stage holds the K chunk of ALL rows (A blocks
//     are sub-ranges of the B tile, so every byte is loaded once per CTA).  Row blocks may live on
//     different GPUs: one TMA descriptor per block (peer memory is TMA-addressable through the
//     NVLink mapping); SWIZZLE_128B, K-major; out-of-range rows/columns are zero-filled by TMA.
//   * warp roles: warp 0 = TMA producer, warp 1 = MMA issuer (one elected thread issues
//     tcgen05.mma.kind::tf32), warps 2-9 = converter (ncu showed the 4-warp converter, not HBM or the tensor
//     pipe, pacing the pipeline), warps 2-5 also run the epilogue.  The converter scans for NaN/inf
//     (nan_to_num in place, rare) and, for 3xTF32, writes lo2 = 2 (x - trunc_tf32(x)) into a second smem tile:
//     the TMA tile itself is the hi operand (the tensor core truncates), and  hi*hi^T + hi*lo2^T  symmetrised by
//     the consumers equals hi*hi + hi*lo + lo*hi (~fp32 accuracy; distances suffer cancellation) with 2 MMAs and
//     2 instead of 3 shared-memory passes per element -- shared-memory bandwidth was the limit of the 3-MMA form.
//     NOTE: G is therefore NOT symmetric as written; every reader applies 0.5 (G + G^T).
//   * pipeline: full[s] (TMA -> converter), ready[s] (converter -> MMA), empty[s] (tcgen05.commit ->
//     TMA), done (last commit -> epilogue).
#include "common.cuh"
#include "tc_common.cuh"
#include <cstring>
#include <cstdlib>

#define GRAM_MAX_BLOCKS 9          // 8 peers + 1 extra row block
#define GRAM_CHUNK 32              // floats per K chunk = 128 B = one swizzle row

struct GramParams {
    CUtensorMap maps[GRAM_MAX_BLOCKS];
    int n_blocks;
    int blk_rows_pad[GRAM_MAX_BLOCKS];   // TMA box rows (multiple of 8)
    int blk_smem_row[GRAM_MAX_BLOCKS];   // first row of the block inside the stage tile
    int rows_covered;                    // sum of blk_rows_pad
    int tile_rows;                       // n_mblk * 128
    int np_n;                            // MMA N extent (multiple of 16, >= rows_covered, <= 512)
    int n_mblk;
    int mb_per_cta;
    int stages;
    int split3;                          // 1 = 3xTF32
    int slabs;                           // column slabs per stage (K chunk = slabs * row_bytes/4 floats per row)
    int dbg;                             // debug bisect: 1 = converter does nothing, 2 = no MMA issue
    int row_bytes;                       // 128 (SWIZZLE_128B) or 64 (SWIZZLE_64B: halves the stage so that 3xTF32
                                         // staging of up to 512 rows still gets a 3-deep pipeline)
    long long chunk0, chunk1;            // K chunk range of this launch
    float* gram;                         // [tile_rows][ld_gram] fp32, accumulated with red.add
    int ld_gram;
};

namespace {

constexpr int kThreads = 320;            // 10 warps: TMA, MMA, 8 converter (4 of them also epilogue)
constexpr int kConvThreads = 256;

__global__ void __launch_bounds__(kThreads, 1)
gram_tcgen05_kernel(const __grid_constant__ GramParams p) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    // carve: [stages][hi tile][lo tile?] then barriers
    const uint32_t rb = (uint32_t)p.row_bytes;
    const uint32_t slab_bytes = (uint32_t)p.tile_rows * rb;              // one [tile_rows x row_bytes] swizzled tile
    const uint32_t tile_bytes = slab_bytes * (uint32_t)p.slabs;          // hi (or lo) part of a stage
    const uint32_t stage_bytes = tile_bytes * (p.split3 ? 2u : 1u);
    uint8_t* tiles = smem_raw;                       // dynamic smem base is 1024-aligned (checked on host)
    uint64_t* bars = reinterpret_cast<uint64_t*>(tiles + (size_t)p.stages * stage_bytes);
    uint64_t* full = bars;
    uint64_t* ready = bars + p.stages;
    uint64_t* empty = bars + 2 * p.stages;
    uint64_t* done = bars + 3 * p.stages;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 3 * p.stages + 1);

    const int warp = bl::uniform_warp_idx();      // warp-uniform role index (tc_common.cuh "warp-uniform issue")
    const int lane = threadIdx.x & 31;

    // ---- work assignment
    // K chunks are dealt round-robin (chunk = chunk0 + blockIdx.x + it * gridDim.x): at any moment the 148 CTAs
    // read one contiguous ~37 KB window of every row, so DRAM pages are shared between neighbouring CTAs.
    // (A blocked split -- each CTA walking its own far-apart K range -- topped out at 61 % of the copy
    // bandwidth even with the MMA and converter disabled: 148 x 100 independent 256 B streams.)
    const int ksplits = gridDim.x;
    const long long nchunks = p.chunk1 - p.chunk0;
    const long long kc0 = p.chunk0 + blockIdx.x;
    const int mb0 = blockIdx.y * p.mb_per_cta;
    const int nmb = min(p.mb_per_cta, p.n_mblk - mb0);
    const int iters = (int)((nchunks > (long long)blockIdx.x) ? (nchunks - blockIdx.x + ksplits - 1) / ksplits : 0);
    const uint32_t tmem_cols_needed = (uint32_t)(p.mb_per_cta * p.np_n);
    uint32_t tmem_cols = 32;
    while (tmem_cols < tmem_cols_needed) tmem_cols <<= 1;

    // ---- one-time setup
    if (warp == 0 && lane == 0) {
        for (int b = 0; b < p.n_blocks; ++b) bl::tma_prefetch_desc(&p.maps[b]);
        for (int s = 0; s < p.stages; ++s) {
            bl::mbar_init(&full[s], 1);
            bl::mbar_init(&ready[s], kConvThreads / 32);     // one arrive per converter warp
            bl::mbar_init(&empty[s], 1);
        }
        bl::mbar_init(done, 1);
        bl::fence_barrier_init();
    }
    if (warp == 1) bl::tmem_alloc_dyn(tmem_slot, tmem_cols);
    // rows of the tile that no TMA box covers must read as zero (hi and lo tiles, every stage)
    if (p.rows_covered < p.tile_rows) {
        const uint32_t beg = (uint32_t)p.rows_covered * rb;
        const int n_tiles = p.stages * (p.split3 ? 2 : 1) * p.slabs;     // consecutive slabs
        for (int t = 0; t < n_tiles; ++t) {
            uint8_t* base = tiles + (size_t)t * slab_bytes;
            for (uint32_t off = beg + threadIdx.x * 16u; off < slab_bytes; off += kThreads * 16u)
                *reinterpret_cast<float4*>(base + off) = make_float4(0.f, 0.f, 0.f, 0.f);
        }
        bl::fence_proxy_async_smem();
    }
    bl::tc_fence_before();
    __syncthreads();
    bl::tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (iters > 0) {
        if (warp == 0) {
            // ================= TMA producer (whole warp converged, election inside the asm) =================
            {
                const uint32_t tx = (uint32_t)p.rows_covered * rb * (uint32_t)p.slabs;
                const uint32_t tiles0 = bl::smem_u32(tiles), full0 = bl::smem_u32(full), empty0 = bl::smem_u32(empty);
                int s = 0;
                uint32_t ph = 0;
                const int cf = p.row_bytes / 4;             // floats per slab row
                for (int it = 0; it < iters; ++it) {
                    bl::mbar_wait_u32(empty0 + 8u * s, ph ^ 1u);
                    const uint32_t fb = full0 + 8u * s;
                    bl::mbar_arrive_expect_tx_e(fb, tx);
                    const uint32_t dst = tiles0 + (uint32_t)s * stage_bytes;
                    const int c0 = (int)((kc0 + (long long)it * ksplits) * cf * p.slabs);
                    for (int sl = 0; sl < p.slabs; ++sl)
                        for (int b = 0; b < p.n_blocks; ++b)
                            bl::tma_load_2d_e(dst + (uint32_t)sl * slab_bytes + (uint32_t)p.blk_smem_row[b] * rb,
                                              &p.maps[b], fb, c0 + sl * cf, 0);
                    if (++s == p.stages) { s = 0; ph ^= 1u; }
                }
            }
        } else if (warp == 1) {
            // ================= MMA issuer =================
            const int n_halves = (p.np_n + 255) / 256;
            const uint32_t tiles0 = bl::smem_u32(tiles), ready0 = bl::smem_u32(ready), empty0 = bl::smem_u32(empty);
            const uint32_t sbo = 8u * rb;            // 8-row swizzle atom
            const uint32_t lay = (rb == 128u) ? bl::kLayoutSw128 : bl::kLayoutSw64;
            const uint64_t d0 = bl::umma_smem_desc(tiles0, 16, sbo, lay);     // start-address field advances by adds
            int s = 0;
            uint32_t ph = 0;
            for (int it = 0; it < iters; ++it) {
                bl::mbar_wait_u32(ready0 + 8u * s, ph);
                bl::tc_fence_after();
                const uint32_t hi0 = (uint32_t)s * stage_bytes;
                for (int sl = 0; sl < ((p.dbg & 2) ? 0 : p.slabs); ++sl) {
                    const uint32_t hi = hi0 + (uint32_t)sl * slab_bytes;
                    const uint32_t lo = hi + tile_bytes;
                    for (int m = 0; m < nmb; ++m) {
                        const uint32_t a_off = (uint32_t)(mb0 + m) * 128u * rb;
                        for (int h = 0; h < n_halves; ++h) {
                            const int ncols = min(256, p.np_n - h * 256);
                            const uint32_t idesc = bl::umma_idesc_tf32(128, (uint32_t)ncols, 0, 0);
                            const uint32_t b_off = (uint32_t)h * 256u * rb;
                            const uint32_t d_tmem = tmem_base + (uint32_t)(m * p.np_n + h * 256);
                            for (int k = 0; k < (int)(rb / 32u); ++k) {   // K = 8 tf32 = 32 B per MMA
                                const uint32_t koff = (uint32_t)k * 32u;
                                const uint64_t a_hi = d0 + (uint64_t)((hi + a_off + koff) >> 4);
                                const uint64_t b_hi = d0 + (uint64_t)((hi + b_off + koff) >> 4);
                                bl::umma_tf32_e(d_tmem, a_hi, b_hi, idesc, (it > 0 || sl > 0 || k > 0) ? 1u : 0u);
                                if (p.split3) {      // + hi * (2 lo)^T; its transpose comes from the symmetrisation
                                    const uint64_t b_lo = d0 + (uint64_t)((lo + b_off + koff) >> 4);
                                    bl::umma_tf32_e(d_tmem, a_hi, b_lo, idesc, 1u);
                                }
                            }
                        }
                    }
                }
                bl::umma_commit_e(empty0 + 8u * s);                    // smem slot reusable when MMAs retire
                if (it == iters - 1) bl::umma_commit_e(bl::smem_u32(done));    // accumulators final
                if (++s == p.stages) { s = 0; ph ^= 1u; }
            }
        } else {
            // ================= converter (warps 2..5) =================
            const int ct = threadIdx.x - 64;                       // 0..255
            const uint32_t n16 = (uint32_t)p.rows_covered * (rb / 16u);   // 16 B granules in the covered part of a slab
            int s = 0;
            uint32_t ph = 0;
            for (int it = 0; it < iters; ++it, s = (s + 1 == p.stages) ? 0 : s + 1, ph ^= (s == 0) ? 1u : 0u) {
                bl::mbar_wait(&full[s], ph);
                for (int sl = 0; sl < ((p.dbg & 1) ? 0 : p.slabs); ++sl) {
                uint8_t* hi = tiles + (size_t)s * stage_bytes + (size_t)sl * slab_bytes;
                uint8_t* lo = hi + tile_bytes;
#pragma unroll 4
                for (uint32_t g = ct; g < n16; g += kConvThreads) {
                    float4 x = *reinterpret_cast<float4*>(hi + g * 16u);
                    // The tile TMA wrote IS the hi operand: kind::tf32 reads the top 19 bits of each fp32 word, i.e. it
                    // truncates exactly like `& 0xFFFFE000` -- no rewrite.  Only a non-finite value (rows are sanitised
                    // where they are produced, so this is the rare path) is replaced in place by nan_to_num.
                    const bool bad = !(fabsf(x.x) <= FLT_MAX) || !(fabsf(x.y) <= FLT_MAX) || !(fabsf(x.z) <= FLT_MAX) ||
                                     !(fabsf(x.w) <= FLT_MAX);
                    if (bad) {
                        x.x = bl_sanitize(x.x); x.y = bl_sanitize(x.y); x.z = bl_sanitize(x.z); x.w = bl_sanitize(x.w);
                        *reinterpret_cast<float4*>(hi + g * 16u) = x;
                    }
                    if (p.split3) {
                        // lo tile = 2 * (x - trunc_tf32(x)): the MMA issuer adds hi*hi^T + hi*(2 lo)^T, and every consumer
                        // of G symmetrises it (0.5 (G + G^T) = hi hi^T + hi lo^T + lo hi^T) -- 2 MMAs instead of 3
                        float4 l;
                        l.x = 2.f * (x.x - __uint_as_float(__float_as_uint(x.x) & 0xFFFFE000u));
                        l.y = 2.f * (x.y - __uint_as_float(__float_as_uint(x.y) & 0xFFFFE000u));
                        l.z = 2.f * (x.z - __uint_as_float(__float_as_uint(x.z) & 0xFFFFE000u));
                        l.w = 2.f * (x.w - __uint_as_float(__float_as_uint(x.w) & 0xFFFFE000u));
                        *reinterpret_cast<float4*>(lo + g * 16u) = l;
                    }
                }
                }
                bl::fence_proxy_async_smem();                      // generic writes -> visible to UMMA
                __syncwarp();
                if (lane == 0) bl::mbar_arrive(&ready[s]);
            }
        }

        // ================= epilogue (warps 2..5): TMEM -> red.global.add =================
        if (warp >= 2 && warp < 6) {
            bl::mbar_wait(done, 0);
            bl::tc_fence_after();
            const int q = warp & 3;                                // TMEM lane quarter of this warp
            for (int m = 0; m < nmb; ++m) {
                const int row = (mb0 + m) * 128 + q * 32 + lane;
                for (int c = 0; c < p.np_n; c += 32) {
                    float v[32];
                    const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(m * p.np_n + c);
                    bl::tmem_ld_32x32(taddr, v);
                    if (row < p.rows_covered) {
                        float* dst = p.gram + (size_t)row * p.ld_gram + c;
                        // rows_covered is a multiple of 8 and dst is 128 B aligned: whole 16 B groups, one vector
                        // reduction per 4 accumulators (red.global.add.v4.f32, sm_90+)
#pragma unroll
                        for (int j = 0; j < 32; j += 4)
                            if (c + j < p.rows_covered)
                                asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};"
                                             :: "l"(dst + j), "f"(v[j]), "f"(v[j + 1]), "f"(v[j + 2]), "f"(v[j + 3]) : "memory");
                    }
                }
            }
        }
    }
    bl::tc_fence_before();
    __syncthreads();
    if (warp == 1) bl::tmem_dealloc(tmem_base, tmem_cols);
}

}  // namespace

// Host launcher.  blocks: base pointers of row blocks (each [rows_b x d], row stride ld_b floats).
struct GramBlockDesc {
    const float* base;
    long long ld;        // row stride in floats (multiple of 4)
    int rows;
};

extern "C" int bl_gram_tcgen05(const GramBlockDesc* blocks, int n_blocks, long long d, long long col0,
                               long long col1, float* gram, int ld_gram, int split3, int num_sms,
                               void* stream) {
    if (n_blocks < 1 || n_blocks > GRAM_MAX_BLOCKS) return -1;
    GramParams p;
    memset(&p, 0, sizeof(p));
    p.n_blocks = n_blocks;
    int row = 0;
    for (int b = 0; b < n_blocks; ++b) {
        const int pad = (blocks[b].rows + 7) / 8 * 8;
        if (pad > 256) return -3;                      // TMA box limit; callers split larger blocks
        if (blocks[b].ld % 4 != 0 || ((uintptr_t)blocks[b].base) % 16 != 0) return -4;
        p.blk_rows_pad[b] = pad;
        p.blk_smem_row[b] = row;
        row += pad;
    }
    p.rows_covered = row;
    if (row > 512) return -5;
    p.n_mblk = (row + 127) / 128;
    p.tile_rows = p.n_mblk * 128;
    p.np_n = (row + 15) / 16 * 16;
    if (p.np_n < 16) p.np_n = 16;
    // TMEM budget: mb_per_cta * np_n <= 512 columns; epilogue reads 32-column groups
    p.np_n = (p.np_n + 31) / 32 * 32;
    p.mb_per_cta = 512 / p.np_n;
    if (p.mb_per_cta < 1) return -6;
    if (p.mb_per_cta > p.n_mblk) p.mb_per_cta = p.n_mblk;
    const int groups = (p.n_mblk + p.mb_per_cta - 1) / p.mb_per_cta;
    p.split3 = split3 ? 1 : 0;
    const size_t budget = 227 * 1024 - 1024 - 256;
    // 128 B rows unless the (hi + lo) staging of one stage would leave fewer than 2 pipeline stages
    p.row_bytes = 128;
    if ((size_t)p.tile_rows * 128 * (p.split3 ? 2 : 1) * 2 > budget) p.row_bytes = 64;
    for (int b = 0; b < n_blocks; ++b) {
        uint64_t dims[2] = {(uint64_t)d, (uint64_t)blocks[b].rows};
        uint64_t strides[1] = {(uint64_t)blocks[b].ld * 4};
        uint32_t box[2] = {(uint32_t)p.row_bytes / 4, (uint32_t)p.blk_rows_pad[b]};
        int r = bl::make_tmap_f32(&p.maps[b], blocks[b].base, 2, dims, strides, box,
                                  p.row_bytes == 128 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_64B);
        if (r != 0) return 1000 + r;
    }
    // wider K chunks per row (2 x 128 B) = longer DRAM bursts per row visit and half the barrier round trips
    // per byte (measured: N=100 tf32 1.85 -> 1.55 ms, 3xTF32 2.41 -> 1.93 ms; 4 slabs gave nothing more);
    // needs at least 3 pipeline stages to pay off
    int slabs = 2;
    if ((size_t)p.tile_rows * p.row_bytes * slabs * (p.split3 ? 2 : 1) * 3 > budget) slabs = 1;
    {
        const char* e = getenv("BLADES_GRAM_SLABS");
        if (e) { int v = atoi(e); if (v == 1 || v == 2 || v == 4) slabs = v; }
    }
    p.slabs = slabs;
    {
        const char* e = getenv("BLADES_GRAM_DBG");
        p.dbg = e ? atoi(e) : 0;
    }
    const size_t stage_bytes = (size_t)p.tile_rows * p.row_bytes * slabs * (p.split3 ? 2 : 1);
    int stages = (int)(budget / stage_bytes);
    if (stages > 8) stages = 8;
    if (stages < 2) return -7;
    p.stages = stages;
    const long long chunk = (long long)(p.row_bytes / 4) * slabs;
    if (col0 % chunk != 0) return -2;
    p.chunk0 = col0 / chunk;
    p.chunk1 = (col1 + chunk - 1) / chunk;
    p.gram = gram;
    p.ld_gram = ld_gram;
    const long long nchunks = p.chunk1 - p.chunk0;
    if (nchunks <= 0) return 0;
    int ksplits = (num_sms > 0 ? num_sms : 148) / groups;
    if (ksplits < 1) ksplits = 1;
    if ((long long)ksplits > nchunks) ksplits = (int)nchunks;
    const size_t smem = (size_t)stages * stage_bytes + (3 * stages + 2) * sizeof(uint64_t) + 16;
    static bool attr_done = false;
    if (!attr_done) {
        cudaError_t e = cudaFuncSetAttribute(gram_tcgen05_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
        if (e != cudaSuccess) return (int)e;
        attr_done = true;
    }
    dim3 grid(ksplits, groups);
    gram_tcgen05_kernel<<<grid, kThreads, smem, (cudaStream_t)stream>>>(p);
    return (int)cudaGetLastError();
}
