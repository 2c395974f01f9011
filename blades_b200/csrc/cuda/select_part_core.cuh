// Partition-only trimmed mean (per coordinate, in registers) -- the core shared by coord_select_part_kernel and by
// the host-side checker (csrc/host/select_core_check.cu compiles this very file for the CPU).
//
// The trimmed mean does not need the NP values sorted, only partitioned into bottom-b | middle | top-b.  For
// NP = 4Q real values held as two halves a[2Q], b[2Q] and trim count b = Q:
//   1. sort each half (2 x SortNet<2Q>: 610 compare-exchanges for NP = 80, vs 849 for SortNet<80>);
//   2. one bitonic split per end: the Q smallest of the union are {min(a[i], b[Q-1-i])}, the Q largest are
//      {max(a[Q+i], b[2Q-1-i])}; the 2Q losers of the two splits ARE the middle.
// A virtual value m of multiplicity f >= Q (ALIE / IPM rows) merges analytically: the lowest Q of the
// merged multiset are the bottom-set members below m plus copies of m, so every bottom-set member x >= m stays
// (contributing x) while a copy of m is trimmed instead -- and symmetrically at the top:
//   kept = sum(middle) + sum_{x in Bot} max(x - m, 0) + sum_{x in Top} min(x - m, 0) + f*m,   N - 2Q values kept.
// No large-minus-large cancellation: the middle is summed directly and the corrections are differences to m.
#pragma once

#ifndef BL_CORE_FN
#define BL_CORE_FN __device__ __forceinline__
#endif

BL_CORE_FN float bl_sat01(float x) {
#ifdef __CUDA_ARCH__
    return __saturatef(x);
#else
    return fminf(fmaxf(x, 0.f), 1.f);
#endif
}

// Virtual value from the statistics of the first n_stat values in LOAD order (a[0..H) are rows 0..H-1, b[0..H) rows
// H..NP-1): kind 1 = ALIE mean - p * std_unbiased, kind 2 = IPM -p * mean (same arithmetic as coord_select_kernel).
template <int NP>
BL_CORE_FN float bl_virtual_value(const float (&a)[NP / 2], const float (&b)[NP / 2], int n_stat, int kind, float param) {
    constexpr int H = NP / 2;
    const float fstat = (float)n_stat;
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < H; ++i) s = fmaf(a[i], bl_sat01(fstat - (float)i), s);
#pragma unroll
    for (int i = 0; i < H; ++i) s = fmaf(b[i], bl_sat01(fstat - (float)(H + i)), s);
    const float mu = s / fstat;
    if (kind != 1) return -param * mu;
    float q = 0.f;
#pragma unroll
    for (int i = 0; i < H; ++i) { const float d = (a[i] - mu) * bl_sat01(fstat - (float)i); q = fmaf(d, d, q); }
#pragma unroll
    for (int i = 0; i < H; ++i) { const float d = (b[i] - mu) * bl_sat01(fstat - (float)(H + i)); q = fmaf(d, d, q); }
    return mu - param * sqrtf(q / (fstat - 1.f));
}

// Trimmed mean of the NP real values (+ f copies of m), trimming Q = NP/4 from each end.  Requires f == 0 or f >= Q.
// Destroys a and b (sorted in place).
template <int NP>
BL_CORE_FN float bl_trimmed_partition(float (&a)[NP / 2], float (&b)[NP / 2], float m, int f) {
    static_assert(NP % 8 == 0 && NP >= 8 && NP <= 128, "halves must be SortNet sizes (multiples of 4)");
    constexpr int H = NP / 2, Q = NP / 4;
    SortNet<H>::run(a);
    SortNet<H>::run(b);
    float mid0 = 0.f, mid1 = 0.f, ext = 0.f;
    const float use = f > 0 ? 1.f : 0.f;
#pragma unroll
    for (int i = 0; i < Q; ++i) {
        const float x = a[i], y = b[Q - 1 - i];
        mid0 += fmaxf(x, y);                                   // loser of the bottom split -> middle
        ext = fmaf(use, fmaxf(fminf(x, y) - m, 0.f), ext);     // bottom-set member above m stays (as x - m + m)
    }
#pragma unroll
    for (int i = 0; i < Q; ++i) {
        const float x = a[Q + i], y = b[H - 1 - i];
        mid1 += fminf(x, y);                                   // loser of the top split -> middle
        ext = fmaf(use, fminf(fmaxf(x, y) - m, 0.f), ext);     // top-set member below m stays
    }
    const float kept = (mid0 + mid1) + ext + (float)f * (f > 0 ? m : 0.f);
    return kept / (float)(NP + f - 2 * Q);
}
