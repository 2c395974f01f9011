// Coordinate-wise robust selection over the client dimension -- K3/K4 (+K7 prologue, K8 epilogue)
// of SURVEY 2.7: trimmed mean (reference trimmedmean.py:38-41: two strided topk + 3 temporaries)
// and median (reference median.py:23-24: two kthvalue passes) in ONE streaming pass.
//
// One thread owns one coordinate: it loads that coordinate from every client row (rows may live
// on peer GPUs -- plain global loads on NVLink-mapped pointers, coalesced 128 B per warp per row),
// sorts the <=128 values in registers with a static pruned Batcher network, and reduces the ranks
// it needs.  ALIE / IPM attackers are *virtual rows*: their common value (mean - z*std, or
// -eps*mean, over the honest rows) is computed from the same registers and merged analytically
// with multiplicity f -- f identical malicious rows are never stored or sorted.
// The result is written to every replica and theta += lr*agg is applied in the same kernel.
#include "common.cuh"
#include <cstdlib>

#define CE(a, b) { float lo_ = fminf(v[a], v[b]); v[b] = fmaxf(v[a], v[b]); v[a] = lo_; }
#include "gen/sortnet_gen.cuh"
#include "select_part_core.cuh"
#undef CE

struct SelectParams {
    const float* rows[128];   // real rows: honest first (stat rows), then other real rows
    int n_real;               // number of real rows (<= NP)
    int n_stat;               // first n_stat rows enter the attack statistics
    int n_virtual;            // multiplicity f of the virtual row
    int virt_kind;            // 0 none, 1 ALIE (mean - p*std_unbiased), 2 IPM (-p*mean)
    float virt_param;
    int mode;                 // 0 trimmed mean, 1 median
    int trim_b;
    long long c0, c1;         // coordinate range owned by this launch
    BlEpilogue ep;
};

// Pipe balance (ncu: the ALU pipe -- FMNMX/ISETP/SEL, 16 lanes/clk/SMSP -- is the limiter): the sorting
// network has to live on the ALU pipe, so everything else is written as FFMA / FADD.SAT arithmetic for
// the otherwise idle FMA pipe: masks are 0/1 floats, rank tests are saturating adds.
// Block size: the straight-line network is ~30-50 KB of SASS, more than the 32 KB L1.5 instruction cache, and ncu
// showed "no_instruction" as the top stall with 128-thread blocks (20 independent warps per SM each streaming
// the code at a different position).  Large blocks keep the warps of an SM roughly in lockstep so they share
// instruction-cache lines; two resident blocks per SM still overlap one block's load phase with the other's sort.
template <int NP> struct SelectBlock { static constexpr int kMax = NP <= 80 ? 640 : (NP <= 104 ? 512 : 384); };

template <int NP, int MODE>
__global__ void __launch_bounds__(SelectBlock<NP>::kMax)
coord_select_kernel(const __grid_constant__ SelectParams p) {
    const long long c = p.c0 + (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= p.c1) return;
    float v[NP];
    const int n = p.n_real;
    // Issue ALL row loads back to back before the first use (rows[i >= n] alias row 0 on the host
    // side, so no load is predicated): one DRAM/NVLink round trip per thread instead of NP.
    // 32-bit element offset from the (uniform) row base: no per-load 64-bit address arithmetic on the ALU pipe.
    const unsigned cu = (unsigned)c;
#pragma unroll
    for (int i = 0; i < NP; ++i) v[i] = __ldcs(p.rows[i] + cu);      // ld.global.cs: streaming, evict-first
#pragma unroll
    for (int i = 0; i < NP; ++i) v[i] = bl_sanitize(v[i]);

    // ---- attack prologue (K7): statistics of the honest rows (= the first n_stat slots), load order
    float m = 0.f;
    const int f = p.n_virtual;
    const float fstat = (float)p.n_stat;
    if (f > 0) {
        float s = 0.f;
#pragma unroll
        for (int i = 0; i < NP; ++i) s = fmaf(v[i], __saturatef(fstat - (float)i), s);     // mask = [i < n_stat]
        const float mu = s / fstat;
        if (p.virt_kind == 1) {
            float q = 0.f;
#pragma unroll
            for (int i = 0; i < NP; ++i) {
                const float d = (v[i] - mu) * __saturatef(fstat - (float)i);
                q = fmaf(d, d, q);
            }
            m = mu - p.virt_param * sqrtf(q / (fstat - 1.f));
        } else {
            m = -p.virt_param * mu;
        }
    }
    // padding slots sort to the top: FLT_MAX (finite, so 0-weight products stay 0); n > NP - 8 by dispatch,
    // so only the last 7 slots can be padding
#pragma unroll
    for (int i = (NP >= 8 ? NP - 7 : 0); i < NP; ++i) v[i] = (i < n) ? v[i] : FLT_MAX;

    SortNet<NP>::run(v);

    // ---- merge the virtual value with multiplicity f: r = #real values below m
    const int N = n + f;
    float rf = 0.f;
    if (f > 0) {
#pragma unroll
        for (int i = 0; i < NP; ++i) rf += (v[i] < m) ? 1.f : 0.f;
        rf = fminf(rf, (float)n);               // padding (FLT_MAX) never counts as a real value
    }
    const float ff = (float)f;
    float agg;
    if (MODE == 0) {
        const float lo = (float)p.trim_b, hi = (float)(N - p.trim_b);   // keep merged ranks [lo, hi)
        float s = 0.f;
#pragma unroll
        for (int i = 0; i < NP; ++i) {
            const float pos = fmaf(ff, __saturatef((float)(i + 1) - rf), (float)i);       // i + f*[i >= r]
            const float keep = __saturatef(pos - lo + 1.f) * __saturatef(hi - pos);      // [lo <= pos < hi]
            s = fmaf(keep, v[i], s);
        }
        if (f > 0) {
            const float a = fmaxf(rf, lo), b = fminf(rf + ff, hi);
            s = fmaf(m, fmaxf(b - a, 0.f), s);
        }
        agg = s / (hi - lo);
    } else {
        const float k0 = (float)((N - 1) >> 1), k1 = (float)(N >> 1);
        float a0 = 0.f, a1 = 0.f;
        if (f > 0) {
            a0 = m * __saturatef(k0 - rf + 1.f) * __saturatef(rf + ff - k0);
            a1 = m * __saturatef(k1 - rf + 1.f) * __saturatef(rf + ff - k1);
        }
#pragma unroll
        for (int i = 0; i < NP; ++i) {
            const float pos = fmaf(ff, __saturatef((float)(i + 1) - rf), (float)i);
            a0 = fmaf(__saturatef(pos - k0 + 1.f) * __saturatef(k0 + 1.f - pos), v[i], a0);
            a1 = fmaf(__saturatef(pos - k1 + 1.f) * __saturatef(k1 + 1.f - pos), v[i], a1);
        }
        agg = 0.5f * (a0 + a1);
    }
    bl_epilogue_store(p.ep, c, agg);
}

// ---------------------------------------------------------------------------------------------
// Partition-only trimmed mean (select_part_core.cuh): n_real == NP == 4 * trim_b and (no virtual rows or f >= b)
// -- the "20 % attackers, Trimmedmean(nb = f)" family (N = 10k clients: 8k honest rows, b = 2k).  Two half-size
// sorts + two bitonic splits instead of one full network: ~21 % fewer FMNMX on the pipe that bounds this kernel.
template <int NP>
__global__ void __launch_bounds__(SelectBlock<NP>::kMax)
coord_select_part_kernel(const __grid_constant__ SelectParams p) {
    const long long c = p.c0 + (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= p.c1) return;
    constexpr int H = NP / 2;
    float a[H], b[H];
    const unsigned cu = (unsigned)c;
#pragma unroll
    for (int i = 0; i < H; ++i) a[i] = __ldcs(p.rows[i] + cu);
#pragma unroll
    for (int i = 0; i < H; ++i) b[i] = __ldcs(p.rows[H + i] + cu);
#pragma unroll
    for (int i = 0; i < H; ++i) { a[i] = bl_sanitize(a[i]); b[i] = bl_sanitize(b[i]); }
    const int f = p.n_virtual;
    float m = 0.f;
    if (f > 0) m = bl_virtual_value<NP>(a, b, p.n_stat, p.virt_kind, p.virt_param);
    bl_epilogue_store(p.ep, c, bl_trimmed_partition<NP>(a, b, m, f));
}

// ---------------------------------------------------------------------------------------------
// Large-N fallback (128 < N <= 512): a block sorts a [NP x 32-coordinate] tile in shared memory
// with a bitonic network (one __syncthreads per stage).  Virtual rows are materialised into the
// tile (value computed per coordinate first).
struct SelectLargeParams {
    const float* rows[BL_MAX_ROWS];
    int n_real, n_stat, n_virtual, virt_kind;
    float virt_param;
    int mode, trim_b;
    long long c0, c1;
    BlEpilogue ep;
};

template <int NP>
__global__ void __launch_bounds__(256)
coord_select_large_kernel(const __grid_constant__ SelectLargeParams p) {
    extern __shared__ float tile[];                 // [NP][33]
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nw = blockDim.x >> 5;
    const long long c = p.c0 + (long long)blockIdx.x * 32 + lane;
    const bool live = c < p.c1;
    const int n = p.n_real, f = p.n_virtual, N = n + f;
    for (int i = wid; i < NP; i += nw) {
        float x = INFINITY;
        if (i < n && live) x = bl_sanitize(bl_ldg_stream(p.rows[i] + c));
        tile[i * 33 + lane] = x;
    }
    __syncthreads();
    if (f > 0) {
        // per-coordinate statistics by warp 0 (n_stat <= 512 values, sequential per lane)
        if (wid == 0) {
            float s = 0.f;
            for (int i = 0; i < p.n_stat; ++i) s += tile[i * 33 + lane];
            const float mu = s / (float)p.n_stat;
            float m;
            if (p.virt_kind == 1) {
                float q = 0.f;
                for (int i = 0; i < p.n_stat; ++i) { float d = tile[i * 33 + lane] - mu; q += d * d; }
                m = mu - p.virt_param * sqrtf(q / (float)(p.n_stat - 1));
            } else m = -p.virt_param * mu;
            for (int i = n; i < N; ++i) tile[i * 33 + lane] = live ? m : INFINITY;
        }
        __syncthreads();
    }
    for (int k = 2; k <= NP; k <<= 1)
        for (int j = k >> 1; j > 0; j >>= 1) {
            for (int t = wid; t < NP / 2; t += nw) {
                const int i = 2 * t - (t & (j - 1));      // index with bit j clear
                const int ixj = i + j;
                const bool up = ((i & k) == 0);
                float a = tile[i * 33 + lane], b = tile[ixj * 33 + lane];
                const float lo = fminf(a, b), hi = fmaxf(a, b);
                tile[i * 33 + lane] = up ? lo : hi;
                tile[ixj * 33 + lane] = up ? hi : lo;
            }
            __syncthreads();
        }
    if (wid == 0 && live) {
        float agg;
        if (p.mode == 0) {
            float s = 0.f;
            for (int i = p.trim_b; i < N - p.trim_b; ++i) s += tile[i * 33 + lane];
            agg = s / (float)(N - 2 * p.trim_b);
        } else {
            agg = 0.5f * (tile[((N - 1) >> 1) * 33 + lane] + tile[(N >> 1) * 33 + lane]);
        }
        bl_epilogue_store(p.ep, c, agg);
    }
}

static int select_block_size(int kmax) {
    static int forced = -1;
    if (forced < 0) {
        const char* e = getenv("BLADES_SELECT_BLOCK");
        forced = e ? atoi(e) : 0;
    }
    int b = forced > 0 ? forced : 256;       // measured best of {128, 256, 320, 640}
    if (b > kmax) b = kmax;
    return (b / 32) * 32;
}

static bool select_partition_enabled() {
    static const bool on = [] { const char* e = getenv("BLADES_SELECT_PARTITION"); return !(e && e[0] == '0'); }();
    return on;
}

// The partition-only kernel applies when the real rows fill the padded size exactly, the trim count is a quarter of
// them and a virtual row (if any) has multiplicity >= the trim count (select_part_core.cuh).
static bool partition_applies(const SelectParams& p, int NP) {
    return p.mode == 0 && p.n_real == NP && p.trim_b * 4 == NP && (p.n_virtual == 0 || p.n_virtual >= p.trim_b)
           && (p.n_virtual == 0 || p.n_stat >= 2 || p.virt_kind != 1) && select_partition_enabled();
}

template <int NP>
static bool launch_partition(const SelectParams& p, unsigned grid, int block, cudaStream_t st) {
    if constexpr (NP % 8 == 0) {
        if (partition_applies(p, NP)) {
            coord_select_part_kernel<NP><<<grid, block, 0, st>>>(p);
            return true;
        }
    }
    return false;
}

// Which kernel bl_coord_select would launch for these parameters (no CUDA call: usable without a GPU by the tests):
// 0 = none (n_real outside 1..128: the large-N kernel is a different entry point), 1 = full sorting network,
// 2 = partition-only trimmed mean.
extern "C" int bl_coord_select_choice(const SelectParams* p) {
    if (p->n_real < 1 || p->n_real > 128) return 0;
    return partition_applies(*p, (p->n_real + 7) / 8 * 8) ? 2 : 1;
}

template <int NP>
static cudaError_t launch_small(const SelectParams& p, cudaStream_t st) {
    const long long cols = p.c1 - p.c0;
    if (cols <= 0) return cudaSuccess;
    if (p.c1 > 0xFFFFFFFFLL) return cudaErrorInvalidValue;      // 32-bit element offsets
    const int block = select_block_size(SelectBlock<NP>::kMax);
    const unsigned grid = (unsigned)((cols + block - 1) / block);
    if (launch_partition<NP>(p, grid, block, st)) return cudaGetLastError();
    if (p.mode == 0) coord_select_kernel<NP, 0><<<grid, block, 0, st>>>(p);
    else coord_select_kernel<NP, 1><<<grid, block, 0, st>>>(p);
    return cudaGetLastError();
}

extern "C" int bl_coord_select(const SelectParams* p, void* stream) {
    cudaStream_t st = (cudaStream_t)stream;
    const int n = p->n_real;
    if (n < 1 || n > 128) return -1;
    switch ((n + 7) / 8) {
#define CASE(K) case K: return (int)launch_small<8 * K>(*p, st);
        CASE(1) CASE(2) CASE(3) CASE(4) CASE(5) CASE(6) CASE(7) CASE(8)
        CASE(9) CASE(10) CASE(11) CASE(12) CASE(13) CASE(14) CASE(15) CASE(16)
#undef CASE
    }
    return -1;
}

extern "C" int bl_coord_select_large(const SelectLargeParams* p, void* stream) {
    cudaStream_t st = (cudaStream_t)stream;
    const int N = p->n_real + p->n_virtual;
    if (N < 1 || N > BL_MAX_ROWS) return -1;
    const long long cols = p->c1 - p->c0;
    if (cols <= 0) return 0;
    const unsigned grid = (unsigned)((cols + 31) / 32);
    int np = 64;
    while (np < N) np <<= 1;
    const size_t smem = (size_t)np * 33 * sizeof(float);
    cudaError_t e;
#define LAUNCH(NPV)                                                                              \
    e = cudaFuncSetAttribute(coord_select_large_kernel<NPV>,                                       \
                             cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);              \
    if (e != cudaSuccess) return (int)e;                                                           \
    coord_select_large_kernel<NPV><<<grid, 256, smem, st>>>(*p);
    if (np == 64) { LAUNCH(64) } else if (np == 128) { LAUNCH(128) }
    else if (np == 256) { LAUNCH(256) } else { LAUNCH(512) }
#undef LAUNCH
    return (int)cudaGetLastError();
}

extern "C" int bl_sizeof_select_params() { return (int)sizeof(SelectParams); }
extern "C" int bl_sizeof_select_large_params() { return (int)sizeof(SelectLargeParams); }
