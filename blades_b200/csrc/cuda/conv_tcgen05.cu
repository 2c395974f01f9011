// Implicit-GEMM convolution on the 5th-gen tensor cores: forward (fprop) and input-gradient (dgrad) of conv2d /
// linear layers for the client-batched training pass -- K9 of SURVEY 2.7 (reference call site: the autograd
// forward/backward of client.py:178-193, which runs cuDNN / cuBLAS there).
//
//     out[pix(m), n] = sum_{tap} sum_{c}  src[pix(m) * cs + off(tap), c] * Wt[tap][n][c]        (+ add, + bias)
//
// One kernel serves every case; the host describes the problem as a list of PHASES, each with a TAP list:
//   * fprop, stride cs: one phase, taps (r, s) with source offset (r - pad, s - pad); the A operand (activations,
//     K-major: one row = 32 channels of one input pixel = 128 B) is gathered by ONE strided 4-D TMA box per tap and
//     32-channel block straight from the NHWC tensor -- no im2col matrix; conv padding = TMA out-of-bounds zero fill.
//     B = the channels_last weight [Cout][kh*kw*Cin], K-major 2-D boxes.
//   * dgrad, stride 1: one phase, flipped taps over the output gradient; B = the SAME weight matrix read "MN-major"
//     (for tap t the [Cout x Cin] slice is K x N with N contiguous), so no transposed copy of the weights exists.
//   * dgrad, stride s > 1: s*s phases, one per input-pixel parity class; each is a unit-stride gather over the output
//     gradient with the taps whose (pixel + pad - r) is divisible by s, written to the strided pixel positions.
//   * linear layers are 1x1 convolutions over 1x1 images.
// Taps that can only ever see padding (3x3 convs on 1x1 / 2x2 maps) are dropped on the host.
//
// Tile = 128 output pixels (a box of bw x bh x bb pixels: full rows, then whole images) x BN <= 256 channels.
//   warp 0    TMA producer (per K step of 32 channels: A box 16 KB + B tile BN*128 B, SWIZZLE_128B)
//   warp 1    MMA issuer: tcgen05.mma.kind::tf32, M = 128, N = BN, four K = 8 atoms per stage; fp32 accumulators
//             double-buffered in TMEM (2 x 256 columns) so the next tile's MMAs overlap this tile's epilogue
//   warps 2-9 epilogue: tcgen05.ld -> (+ bias, + residual/accumulate) -> smem transpose -> coalesced 128 B row stores
// Persistent CTAs walk the tiles round-robin (m fastest: concurrently running CTAs share the weight tile in L2).
#include "common.cuh"
#include "tc_common.cuh"
#include <cstring>
#include <cstdlib>

#define CONV_MAX_TAPS 25
#define CONV_MAX_PHASES 4

struct ConvPhase {
    int ntaps;
    int oh_off, ow_off;                 // output pixel of grid point (h, w) = (ostep*h + oh_off, ostep*w + ow_off)
    int pad_;
    short dy[CONV_MAX_TAPS], dx[CONV_MAX_TAPS];     // source offset of the tap (conv padding already subtracted)
    short widx[CONV_MAX_TAPS];                      // tap index r*kw + s into the weight matrix
    short pad2_;
};

// Host-side problem description (mirrored by ops/conv.py with ctypes).
struct ConvDesc {
    const float* src;        // NHWC source [NB][Hs][Ws][lds] (Cs valid channels)
    const float* w;          // weight matrix [w_rows][ldw]
    float* out;              // NHWC output [NB][Hout][Wout][ldc] (N valid channels)
    const float* add;        // optional: out = acc + add (same indexing as out; may alias out)
    const float* bias;       // optional [N]
    int NB, Hs, Ws, Cs, lds;
    int w_rows, w_cols, ldw;
    int mode;                // 0 = fprop (B K-major), 1 = dgrad (B MN-major)
    int N;                   // output channels
    int wtap_stride;         // weight columns per tap (= Cin)
    int Hout, Wout, ldc;
    int Ht, Wt;              // output grid per phase
    int cs;                  // source stride
    int ostep;               // output pixel step
    int n_phases;
    int accumulate_only;     // phases without taps are skipped instead of zero-filled
    int num_sms;
    ConvPhase ph[CONV_MAX_PHASES];
};

struct ConvParams {
    CUtensorMap map_a, map_b;
    int b_mn_major, cblocks, wtap_stride;
    int N, BN, n_tiles;
    int bw, bh, bb, rows;
    int Ht, NB, h_tiles, b_tiles;
    int cs, n_phases, ostep;
    int Hout, Wout, ldc;
    float* out;
    const float* add;
    const float* bias;
    int accumulate_only, stages, vec_ok;
    int kpack;               // 32-channel K blocks per pipeline stage (fewer barrier round trips per byte)
    int dbg;                 // timing bisect (BLADES_CONV_DBG): 1 no B loads, 2 no A loads, 4 no MMAs, 8 no stores
    ConvPhase ph[CONV_MAX_PHASES];
};

namespace {
constexpr int kCThreads = 320;
constexpr int kCEpiLd = 36;
constexpr uint32_t kABytes = 128u * 128u;        // A slot: 128 rows x 128 B

__global__ void __launch_bounds__(kCThreads, 1)
conv_tcgen05_kernel(const __grid_constant__ ConvParams p) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    const uint32_t b_bytes = (uint32_t)p.BN * 128u;
    const uint32_t a_stage = kABytes * (uint32_t)p.kpack;
    const uint32_t stage_bytes = a_stage + b_bytes * (uint32_t)p.kpack;
    uint8_t* tiles = smem_raw;
    uint64_t* bars = reinterpret_cast<uint64_t*>(tiles + (size_t)p.stages * stage_bytes);
    uint64_t* full = bars;
    uint64_t* empty = bars + p.stages;
    uint64_t* tfull = bars + 2 * p.stages;
    uint64_t* tempty = bars + 2 * p.stages + 2;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * p.stages + 4);
    float* epi = reinterpret_cast<float*>(tiles + (size_t)p.stages * stage_bytes + 256);

    const int warp = bl::uniform_warp_idx(), lane = threadIdx.x & 31;
    const int m_tiles = p.h_tiles * p.b_tiles;
    const long long total_tiles = (long long)p.n_phases * p.n_tiles * m_tiles;

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

    // Warps 0 and 1 run their loops CONVERGED (all 32 lanes, uniform values, election inside the asm statements): see
    // tc_common.cuh "warp-uniform issue".  Stage / parity are running counters (no integer divisions).
    if (warp == 0) {
        // ================= TMA producer =================
        const uint32_t tx = (((p.dbg & 2) ? 0u : (uint32_t)p.rows * 128u) + ((p.dbg & 1) ? 0u : b_bytes)) *
                            (uint32_t)p.kpack;
        int s = 0;
        uint32_t par = 0;
        for (long long tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
            const int m = (int)(tile % m_tiles);
            const long long rest = tile / m_tiles;
            const int nt = (int)(rest % p.n_tiles);
            const ConvPhase& ph = p.ph[(int)(rest / p.n_tiles)];
            const int h0 = (m % p.h_tiles) * p.bh, b0 = (m / p.h_tiles) * p.bb;
            for (int t = 0; t < ph.ntaps; ++t) {
                const int xw = ph.dx[t], xh = h0 * p.cs + ph.dy[t];
                const int wcol = (int)ph.widx[t] * p.wtap_stride;
                for (int cb = 0; cb < p.cblocks; cb += p.kpack) {
                    bl::mbar_wait_u32(empty0 + 8u * s, par ^ 1u);
                    const uint32_t fb = full0 + 8u * s;
                    bl::mbar_arrive_expect_tx_e(fb, tx);
                    const uint32_t dst = tiles0 + (uint32_t)s * stage_bytes;
                    for (int kp = 0; kp < p.kpack; ++kp) {
                        const int c0 = (cb + kp) * 32;
                        if (!(p.dbg & 2)) bl::tma_load_4d_e(dst + kp * kABytes, &p.map_a, fb, c0, xw, xh, b0);
                        const uint32_t db = dst + a_stage + kp * b_bytes;
                        if (p.dbg & 1) {
                        } else if (!p.b_mn_major) {
                            bl::tma_load_2d_e(db, &p.map_b, fb, wcol + c0, nt * p.BN);
                        } else {
                            for (int j = 0; j < p.BN / 32; ++j)
                                bl::tma_load_2d_e(db + j * 4096, &p.map_b, fb, wcol + nt * p.BN + j * 32, c0);
                        }
                    }
                    if (++s == p.stages) { s = 0; par ^= 1u; }
                }
            }
        }
    } else if (warp == 1) {
        // ================= MMA issuer =================
        const uint32_t idesc = bl::umma_idesc_tf32(128, (uint32_t)p.BN, 0, (uint32_t)p.b_mn_major);
        uint32_t tcount = 0;
        int s = 0;
        uint32_t par = 0;
        // descriptors of stage 0 / K atom 0; the start-address field (bits 0-13, address >> 4) advances by plain adds
        const uint64_t ad0 = bl::umma_smem_desc(tiles0, 16, 1024, bl::kLayoutSw128);
        const uint64_t bd0 = p.b_mn_major ? bl::umma_smem_desc(tiles0 + a_stage, 4096, 512, bl::kLayoutSw128Base32B)
                                          : bl::umma_smem_desc(tiles0 + a_stage, 16, 1024, bl::kLayoutSw128);
        const uint32_t bk = p.b_mn_major ? (1024u >> 4) : (32u >> 4);
        for (long long tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
            const ConvPhase& ph = p.ph[(int)((tile / m_tiles) / p.n_tiles)];
            const int ksteps = ph.ntaps * (p.cblocks / p.kpack);
            if (ksteps == 0) continue;
            const uint32_t buf = tcount & 1u;
            const uint32_t tph = (tcount >> 1) & 1u;
            ++tcount;
            bl::mbar_wait_u32(tempty0 + 8u * buf, tph ^ 1u);
            bl::tc_fence_after();
            const uint32_t d_tmem = tmem_base + buf * 256u;
            for (int ks = 0; ks < ksteps; ++ks) {
                bl::mbar_wait_u32(full0 + 8u * s, par);
                bl::tc_fence_after();
                const uint64_t so = (uint64_t)(((uint32_t)s * stage_bytes) >> 4);
                for (int kp = 0; kp < p.kpack; ++kp) {
                    const uint64_t ao = so + (uint64_t)((kp * kABytes) >> 4), bo = so + (uint64_t)((kp * b_bytes) >> 4);
                    if (!(p.dbg & 4)) {
#pragma unroll
                        for (int k = 0; k < 4; ++k)         // K = 8 tf32 (32 B of the 128 B row) per MMA
                            bl::umma_tf32_e(d_tmem, ad0 + ao + (uint64_t)(k * 2), bd0 + bo + (uint64_t)(k * bk), idesc,
                                            (ks > 0 || kp > 0 || k > 0) ? 1u : 0u);
                    }
                }
                bl::umma_commit_e(empty0 + 8u * s);
                if (ks == ksteps - 1) bl::umma_commit_e(tfull0 + 8u * buf);
                if (++s == p.stages) { s = 0; par ^= 1u; }
            }
        }
    } else {
        // ================= epilogue =================
        const int q = warp & 3;
        const int half = (warp - 2) >> 2;
        float* stg = epi + (warp - 2) * (32 * kCEpiLd);
        const int rsub = lane >> 3, csub = (lane & 7) * 4;
        uint32_t tcount = 0;
        for (long long tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
            const int m = (int)(tile % m_tiles);
            const long long rest = tile / m_tiles;
            const int nt = (int)(rest % p.n_tiles);
            const ConvPhase& ph = p.ph[(int)(rest / p.n_tiles)];
            const bool has_k = ph.ntaps > 0;
            if (!has_k && p.accumulate_only) continue;
            const int h0 = (m % p.h_tiles) * p.bh, b0 = (m / p.h_tiles) * p.bb;
            // the 8 tile rows this thread stores: r = q*32 + r0 + rsub, r0 = 0, 4, .., 28
            long long off[8];
            uint32_t okmask = 0;
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                const int r = q * 32 + i * 4 + rsub;
                const int w_ = r % p.bw, hb = r / p.bw;
                const int h_ = h0 + hb % p.bh, b_ = b0 + hb / p.bh;
                const int oh = p.ostep * h_ + ph.oh_off, ow = p.ostep * w_ + ph.ow_off;
                const bool ok = r < p.rows && h_ < p.Ht && b_ < p.NB && oh < p.Hout && ow < p.Wout;
                off[i] = (((long long)b_ * p.Hout + oh) * p.Wout + ow) * p.ldc;
                okmask |= (ok ? 1u : 0u) << i;
            }
            uint32_t buf = 0;
            if (has_k) {
                buf = tcount & 1u;
                const uint32_t tph = (tcount >> 1) & 1u;
                ++tcount;
                bl::mbar_wait(&tfull[buf], tph);
                bl::tc_fence_after();
            }
            for (int cb = half * 32; cb < p.BN; cb += 64) {
                const int n0 = nt * p.BN + cb;
                if (n0 >= p.N) break;
                float v[32];
                if (has_k) {
                    bl::tmem_ld_32x32(tmem_base + ((uint32_t)(q * 32) << 16) + buf * 256u + (uint32_t)cb, v);
                } else {
#pragma unroll
                    for (int j = 0; j < 32; ++j) v[j] = 0.f;
                }
#pragma unroll
                for (int j = 0; j < 32; j += 4)
                    *reinterpret_cast<float4*>(stg + lane * kCEpiLd + j) = make_float4(v[j], v[j + 1], v[j + 2], v[j + 3]);
                __syncwarp();
                float4 bs = make_float4(0.f, 0.f, 0.f, 0.f);
                if (p.bias != nullptr) {
                    if (n0 + csub + 0 < p.N) bs.x = p.bias[n0 + csub + 0];
                    if (n0 + csub + 1 < p.N) bs.y = p.bias[n0 + csub + 1];
                    if (n0 + csub + 2 < p.N) bs.z = p.bias[n0 + csub + 2];
                    if (n0 + csub + 3 < p.N) bs.w = p.bias[n0 + csub + 3];
                }
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    const int r = i * 4 + rsub;
                    float4 t = *reinterpret_cast<const float4*>(stg + r * kCEpiLd + csub);
                    t.x += bs.x; t.y += bs.y; t.z += bs.z; t.w += bs.w;
                    if (((okmask >> i) & 1u) && !(p.dbg & 8)) {
                        const long long o = off[i] + n0 + csub;
                        if (p.vec_ok && n0 + csub + 4 <= p.N) {
                            if (p.add != nullptr) {
                                const float4 a = *reinterpret_cast<const float4*>(p.add + o);
                                t.x += a.x; t.y += a.y; t.z += a.z; t.w += a.w;
                            }
                            *reinterpret_cast<float4*>(p.out + o) = t;
                        } else {
                            const float tv[4] = {t.x, t.y, t.z, t.w};
#pragma unroll
                            for (int e = 0; e < 4; ++e)
                                if (n0 + csub + e < p.N)
                                    p.out[o + e] = tv[e] + (p.add != nullptr ? p.add[o + e] : 0.f);
                        }
                    }
                }
                __syncwarp();
            }
            if (has_k) {
                bl::tc_fence_before();
                __syncwarp();
                if (lane == 0) bl::mbar_arrive(&tempty[buf]);
            }
        }
    }
    bl::tc_fence_before();
    __syncthreads();
    if (warp == 1) bl::tmem_dealloc(tmem_base, 512);
}
}  // namespace

extern "C" int bl_sizeof_conv_desc() { return (int)sizeof(ConvDesc); }

// Tile box of the per-phase output grid [NB][Ht][Wt]: full rows, then whole images, at most 128 pixels.
// Returns 0 when the grid is not supported (a row longer than 128 pixels).
extern "C" int bl_conv_tile_box(int Wt, int Ht, int NB, int cs, int* bw, int* bh, int* bb) {
    if (Wt < 1 || Ht < 1 || NB < 1 || Wt > 128 || Wt * cs > 256) return 0;
    int h = 128 / Wt; if (h > Ht) h = Ht;
    while (h * cs > 256) --h;
    int b = 1;
    if (h == Ht) { b = 128 / (Wt * Ht); if (b < 1) b = 1; if (b > 256) b = 256; }
    *bw = Wt; *bh = h; *bb = b;
    return Wt * h * b;
}

extern "C" int bl_conv_tc(const ConvDesc* d, void* stream) {
    if (d->n_phases < 1 || d->n_phases > CONV_MAX_PHASES) return -1;
    if (((uintptr_t)d->src) % 16 != 0 || ((uintptr_t)d->w) % 16 != 0 || d->lds % 4 != 0 || d->ldw % 4 != 0) return -1;
    if (d->cs < 1 || d->N < 1 || d->Cs < 1) return -1;
    int max_taps = 0;
    for (int i = 0; i < d->n_phases; ++i) {
        if (d->ph[i].ntaps < 0 || d->ph[i].ntaps > CONV_MAX_TAPS) return -1;
        if (d->ph[i].ntaps > max_taps) max_taps = d->ph[i].ntaps;
    }
    // fprop: a partial last channel block would multiply zero-filled A columns with the NEXT tap's weights (0 * w, but
    // NaN for a non-finite w): only allowed when the block ends the weight row (single tap, e.g. linear layers).
    // dgrad: the reduction runs over weight ROWS, out-of-range rows are zero-filled on both operands.
    if (d->mode == 0 && d->Cs % 32 != 0 && max_taps > 1) return -1;
    ConvParams p;
    memset(&p, 0, sizeof(p));
    int bw, bh, bb;
    const int rows = bl_conv_tile_box(d->Wt, d->Ht, d->NB, d->cs, &bw, &bh, &bb);
    if (rows == 0) return -1;
    p.bw = bw; p.bh = bh; p.bb = bb; p.rows = rows;
    p.Ht = d->Ht; p.NB = d->NB;
    p.h_tiles = (d->Ht + bh - 1) / bh;
    p.b_tiles = (d->NB + bb - 1) / bb;
    p.b_mn_major = d->mode ? 1 : 0;
    p.cblocks = (d->Cs + 31) / 32;
    p.wtap_stride = d->wtap_stride;
    p.N = d->N;
    const int gran = p.b_mn_major ? 32 : 16;
    int bn = (d->N + gran - 1) / gran * gran;
    if (bn > 256) bn = 256;
    const long long m_tiles = (long long)p.h_tiles * p.b_tiles * d->n_phases;
    const int sms = d->num_sms > 0 ? d->num_sms : 148;
    while (bn >= 64 && bn % 64 == 0 && m_tiles * ((d->N + bn - 1) / bn) < sms) bn /= 2;
    p.BN = bn;
    p.n_tiles = (d->N + bn - 1) / bn;
    p.cs = d->cs; p.n_phases = d->n_phases; p.ostep = d->ostep;
    p.Hout = d->Hout; p.Wout = d->Wout; p.ldc = d->ldc;
    p.out = d->out; p.add = d->add; p.bias = d->bias;
    p.accumulate_only = d->accumulate_only;
    p.vec_ok = (((uintptr_t)d->out) % 16 == 0) && (d->ldc % 4 == 0) &&
               (d->add == nullptr || ((uintptr_t)d->add) % 16 == 0);
    memcpy(p.ph, d->ph, sizeof(p.ph));
    { const char* e = getenv("BLADES_CONV_DBG"); p.dbg = e ? atoi(e) : 0; }
    {
        uint64_t dims[4] = {(uint64_t)d->Cs, (uint64_t)d->Ws, (uint64_t)d->Hs, (uint64_t)d->NB};
        uint64_t strides[3] = {(uint64_t)d->lds * 4, (uint64_t)d->Ws * d->lds * 4, (uint64_t)d->Hs * d->Ws * d->lds * 4};
        uint32_t box[4] = {32, (uint32_t)(bw * d->cs), (uint32_t)(bh * d->cs), (uint32_t)bb};
        uint32_t es[4] = {1, (uint32_t)d->cs, (uint32_t)d->cs, 1};
        int r = bl::make_tmap_f32(&p.map_a, d->src, 4, dims, strides, box, CU_TENSOR_MAP_SWIZZLE_128B, es);
        if (r != 0) return 1000 + r;
    }
    {
        uint64_t dims[2] = {(uint64_t)d->w_cols, (uint64_t)d->w_rows};
        uint64_t strides[1] = {(uint64_t)d->ldw * 4};
        uint32_t box[2] = {32, (uint32_t)(p.b_mn_major ? 32 : bn)};
        int r = bl::make_tmap_f32(&p.map_b, d->w, 2, dims, strides, box,
                                  p.b_mn_major ? CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B : CU_TENSOR_MAP_SWIZZLE_128B);
        if (r != 0) return 2000 + r;
    }
    const size_t epi_bytes = 8 * 32 * kCEpiLd * sizeof(float);
    const size_t budget = 227 * 1024 - 1024 - 256 - epi_bytes;
    int kpack = 1;
    {
        const char* e = getenv("BLADES_CONV_KPACK");
        const int want = e ? atoi(e) : 2;
        // K blocks per stage: as many as leave a 4-deep pipeline (measured: 64-channel layers 64 -> 58 us with 2 blocks
        // per stage; a 3-deep pipeline of 64 KB stages halves the speed of the stride-2 gathers)
        for (int kp = 4; kp >= 2; kp /= 2)
            if (kp <= want && p.cblocks % kp == 0 && budget / ((kABytes + (size_t)bn * 128) * kp) >= 4) { kpack = kp; break; }
    }
    p.kpack = kpack;
    const size_t stage_bytes = (kABytes + (size_t)bn * 128) * kpack;
    int stages = (int)(budget / stage_bytes);
    if (stages > 8) stages = 8;
    if (stages < 2) return -2;
    p.stages = stages;
    // always ask for > half of the SM's shared memory: one CTA per SM (each CTA allocates all 512 TMEM columns)
    size_t smem = stages * stage_bytes + 256 + epi_bytes;
    if (smem < 120 * 1024) smem = 120 * 1024;
    static bool attr_done = false;
    if (!attr_done) {
        cudaError_t e = cudaFuncSetAttribute(conv_tcgen05_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
        if (e != cudaSuccess) return (int)e;
        attr_done = true;
    }
    const long long total = m_tiles * p.n_tiles;
    int grid = sms;
    if ((long long)grid > total) grid = (int)total;
    if (grid < 1) return 0;
    conv_tcgen05_kernel<<<grid, kCThreads, smem, (cudaStream_t)stream>>>(p);
    return (int)cudaGetLastError();
}
