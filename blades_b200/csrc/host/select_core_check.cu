// Host-side checker of the partition-only trimmed mean: compiles the SAME sources the GPU kernel uses
// (gen/sortnet_gen.cuh, select_part_core.cuh) for the CPU and compares them with a sort-based double-precision
// reference over random, tied, outlier-laden and constant inputs.  Built and run by tests/test_select_core.py
// (nvcc, no GPU needed).  Exit code 0 = all cases agree.
#include <algorithm>
#include <cfloat>
#include <cmath>
#include <cstdio>
#include <random>
#include <vector>

#define BL_SORTNET_FN static __host__ __device__ inline
#define BL_CORE_FN __host__ __device__ inline
#define CE(a, b) { float lo_ = fminf(v[a], v[b]); v[b] = fmaxf(v[a], v[b]); v[a] = lo_; }
#include "../cuda/gen/sortnet_gen.cuh"
#include "../cuda/select_part_core.cuh"
#undef CE

static int g_fail = 0, g_cases = 0;

template <int NP>
static void run(std::mt19937& rng) {
    constexpr int H = NP / 2, Q = NP / 4;
    std::normal_distribution<float> gauss(0.f, 1.f);
    std::uniform_int_distribution<int> small(-3, 3);
    const int fs[] = {0, Q, Q + 1, 2 * Q + 5};
    for (int regime = 0; regime < 5; ++regime)
        for (int fi = 0; fi < 4; ++fi)
            for (int kind = 1; kind <= 2; ++kind)
                for (int rep = 0; rep < 40; ++rep) {
                    const int f = fs[fi];
                    const int n_stat = (rep % 3 == 0) ? NP - 3 : NP;
                    const float param = kind == 1 ? 0.2858f + 0.5f * (rep % 4) : 0.5f + 10.f * (rep % 3);
                    std::vector<float> x(NP);
                    for (int i = 0; i < NP; ++i) {
                        switch (regime) {
                            case 0: x[i] = gauss(rng) * 0.01f; break;                       // update-sized values
                            case 1: x[i] = (float)small(rng); break;                        // heavy ties
                            case 2: x[i] = gauss(rng) + ((i % 7 == 0) ? 1e6f * (i % 2 ? 1.f : -1.f) : 0.f); break;   // outliers
                            case 3: x[i] = 0.25f; break;                                    // constant
                            default: x[i] = gauss(rng) * (1.f + (float)(i % 5)); break;
                        }
                    }
                    float a[H], b[H];
                    for (int i = 0; i < H; ++i) { a[i] = x[i]; b[i] = x[H + i]; }
                    float m = 0.f;
                    if (f > 0) {
                        m = bl_virtual_value<NP>(a, b, n_stat, kind, param);
                        double mu = 0; for (int i = 0; i < n_stat; ++i) mu += x[i];
                        mu /= n_stat;
                        double q = 0; for (int i = 0; i < n_stat; ++i) q += (x[i] - mu) * (x[i] - mu);
                        const double want = kind == 1 ? mu - param * std::sqrt(q / (n_stat - 1)) : -param * mu;
                        double sc = 0; for (int i = 0; i < n_stat; ++i) sc = std::max(sc, (double)std::fabs(x[i]));
                        if (std::fabs(m - want) > 2e-5 * (sc * (1 + param) + 1e-12)) {
                            if (g_fail++ < 10) std::printf("VIRTUAL NP=%d regime=%d kind=%d: got %g want %g\n", NP, regime, kind, m, want);
                        }
                    }
                    const float got = bl_trimmed_partition<NP>(a, b, m, f);
                    std::vector<double> all(x.begin(), x.end());
                    for (int i = 0; i < f; ++i) all.push_back((double)m);
                    std::sort(all.begin(), all.end());
                    double s = 0, scale = 0;
                    for (size_t i = Q; i < all.size() - Q; ++i) { s += all[i]; scale = std::max(scale, std::fabs(all[i])); }
                    const double want = s / (double)(all.size() - 2 * Q);
                    ++g_cases;
                    if (!(std::fabs(got - want) <= 4e-6 * scale + 1e-30)) {
                        if (g_fail++ < 10)
                            std::printf("TRIM NP=%d regime=%d f=%d kind=%d: got %.9g want %.9g (scale %g)\n", NP, regime, f, kind, got, want, scale);
                    }
                }
}

int main() {
    std::mt19937 rng(12345);
    run<8>(rng); run<16>(rng); run<24>(rng); run<32>(rng); run<40>(rng); run<48>(rng); run<56>(rng); run<64>(rng);
    run<72>(rng); run<80>(rng); run<88>(rng); run<96>(rng); run<104>(rng); run<112>(rng); run<120>(rng); run<128>(rng);
    std::printf("%d cases, %d failures\n", g_cases, g_fail);
    return g_fail ? 1 : 0;
}
