// Host-side selectors on the N x N Gram / distance matrix -- "K6" of SURVEY 2.7.
// The reference runs these as Python loops or sklearn calls on the driver (krum.py:21-66,
// geomed.py:71-82, autogm.py:44-65, centeredclipping.py:37-44, clustering.py:39-41).  Here they are
// O(N^2..N^3) double-precision C++ on matrices of at most 513 x 513, called through ctypes from
// aggregators/_gramops.py (numpy twins exist there and are the test oracle).
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <limits>
#include <numeric>
#include <thread>
#include <vector>

namespace {

inline void dist_to_combo(const double* G, int n, const double* w, double* dist) {
    std::vector<double> Gw(n, 0.0);
    for (int i = 0; i < n; ++i) {
        const double* row = G + (size_t)i * n;
        double s = 0.0;
        for (int j = 0; j < n; ++j) s += row[j] * w[j];
        Gw[i] = s;
    }
    double wGw = 0.0;
    for (int i = 0; i < n; ++i) wGw += w[i] * Gw[i];
    for (int i = 0; i < n; ++i) {
        double sq = wGw - 2.0 * Gw[i] + G[(size_t)i * n + i];
        dist[i] = std::sqrt(std::max(sq, 0.0));
    }
}

int weiszfeld(const double* G, int n, const double* alphas, int maxiter, double eps, double ftol,
              int compounding, double* w) {
    std::vector<double> run(alphas, alphas + n), dist(n), nw(n);
    std::fill(w, w + n, 1.0 / n);
    dist_to_combo(G, n, w, dist.data());
    double obj = 0.0;
    for (int i = 0; i < n; ++i) obj += run[i] * dist[i];
    int it = 0;
    for (it = 1; it <= maxiter; ++it) {
        const double prev = obj;
        double sum = 0.0;
        for (int i = 0; i < n; ++i) {
            const double base = compounding ? run[i] : alphas[i];
            nw[i] = std::max(eps, base / std::max(eps, dist[i]));
            sum += nw[i];
        }
        for (int i = 0; i < n; ++i) { nw[i] /= sum; run[i] = nw[i]; w[i] = nw[i]; }
        dist_to_combo(G, n, w, dist.data());
        obj = 0.0;
        for (int i = 0; i < n; ++i) obj += run[i] * dist[i];
        if (std::fabs(prev - obj) < ftol * obj) break;
    }
    return std::min(it, maxiter);
}

}  // namespace

extern "C" {

// score_i = sum of the (n - f - 2) smallest off-diagonal entries of row i (optionally squared again).
void bl_krum_scores(const double* D, int n, int f, int squared_twice, double* scores) {
    const int k = std::max(0, n - f - 2);
    std::vector<double> row(n > 0 ? n - 1 : 0);
    for (int i = 0; i < n; ++i) {
        int t = 0;
        for (int j = 0; j < n; ++j) {
            if (j == i) continue;
            double v = D[(size_t)i * n + j];
            row[t++] = squared_twice ? v * v : v;
        }
        const int kk = std::min<int>(k, (int)row.size());
        std::partial_sort(row.begin(), row.begin() + kk, row.end());
        double s = 0.0;
        for (int j = 0; j < kk; ++j) s += row[j];
        scores[i] = s;
    }
}

int bl_weiszfeld(const double* G, int n, const double* alphas, int maxiter, double eps, double ftol,
                 int compounding, double* w) {
    return weiszfeld(G, n, alphas, maxiter, eps, ftol, compounding, w);
}

void bl_autogm(const double* G, int n, double lamb, int maxiter, double eps, double ftol, int sort_by_index,
               int compounding, double* w) {
    std::vector<double> alpha(n, 1.0 / n), dist(n);
    weiszfeld(G, n, alpha.data(), maxiter, eps, ftol, compounding, w);
    dist_to_combo(G, n, w, dist.data());
    auto objective = [&]() {
        double o = 0.0, a2 = 0.0;
        for (int i = 0; i < n; ++i) { o += alpha[i] * dist[i]; a2 += alpha[i] * alpha[i]; }
        return o + lamb * a2 / 2.0;
    };
    double glob = objective();
    std::vector<int> order(n);
    for (int iter = 0; iter < maxiter; ++iter) {
        const double prev = glob;
        std::iota(order.begin(), order.end(), 0);
        if (!sort_by_index)
            std::stable_sort(order.begin(), order.end(), [&](int a, int b) { return dist[a] < dist[b]; });
        double eta_opt = 1e16, csum = 0.0;
        for (int p = 0; p < n; ++p) {
            csum += dist[order[p]];
            const double eta = (csum + lamb) / (p + 1);
            if (eta - dist[order[p]] < 0) break;
            eta_opt = eta;
        }
        for (int i = 0; i < n; ++i) alpha[i] = std::max(eta_opt - dist[i], 0.0) / lamb;
        weiszfeld(G, n, alpha.data(), maxiter, eps, ftol, compounding, w);
        dist_to_combo(G, n, w, dist.data());
        glob = objective();
        if (std::fabs(prev - glob) < ftol * glob) break;
    }
}

// Centered clipping on the Gram matrix of [u_0..u_{n-1}, m_prev]  (size (n+1)^2); c has n+1 entries.
void bl_centered_clip(const double* G_aug, int n, double tau, int n_iter, double* c) {
    const int m = n + 1;
    std::vector<double> dist(m), nc(m);
    std::fill(c, c + m, 0.0);
    c[n] = 1.0;
    for (int it = 0; it < n_iter; ++it) {
        dist_to_combo(G_aug, m, c, dist.data());
        double ssum = 0.0;
        std::vector<double> scale(n);
        for (int i = 0; i < n; ++i) {
            scale[i] = dist[i] > 0 ? std::min(1.0, tau / dist[i]) : 1.0;
            ssum += scale[i];
        }
        for (int i = 0; i < m; ++i) nc[i] = c[i] * (1.0 - ssum / n);
        for (int i = 0; i < n; ++i) nc[i] += scale[i] / n;
        std::memcpy(c, nc.data(), sizeof(double) * m);
    }
}

// Two-cluster complete-linkage agglomeration on a symmetric 'distance' matrix; labels[i] in {0,1},
// label 0 = cluster of row 0.
void bl_complete_linkage2(const double* dist_in, int n, int64_t* labels) {
    // Label for label what sklearn's AgglomerativeClustering(metric='precomputed', linkage='complete', n_clusters=2)
    // returns (see aggregators/_gramops.py::complete_linkage_2 for the why): upper triangle only, scipy's NN-chain
    // merge order, stable sort of the merges by height, union-find relabelling, label 0 = the root's child with the
    // larger node id.
    if (n <= 0) return;
    if (n == 1) { labels[0] = 0; return; }
    const double INF = std::numeric_limits<double>::infinity();
    std::vector<double> D((size_t)n * n);
    for (int i = 0; i < n; ++i)
        for (int j = 0; j < n; ++j)
            D[(size_t)i * n + j] = (i == j) ? 0.0 : (i < j ? dist_in[(size_t)i * n + j] : dist_in[(size_t)j * n + i]);
    std::vector<int> size(n, 1), chain(n, 0);
    int clen = 0;
    struct Merge { int x, y; double h; int order; };
    std::vector<Merge> merges(n - 1);
    for (int k = 0; k < n - 1; ++k) {
        if (clen == 0) {
            clen = 1;
            for (int i = 0; i < n; ++i) if (size[i] > 0) { chain[0] = i; break; }
        }
        int x, y;
        double cur;
        while (true) {
            x = chain[clen - 1];
            if (clen > 1) { y = chain[clen - 2]; cur = D[(size_t)x * n + y]; }
            else { y = -1; cur = INF; }
            for (int i = 0; i < n; ++i) {
                if (size[i] == 0 || i == x) continue;
                const double d = D[(size_t)x * n + i];
                if (d < cur) { cur = d; y = i; }
            }
            if (clen > 1 && y == chain[clen - 2]) break;
            chain[clen++] = y;
        }
        clen -= 2;
        if (x > y) std::swap(x, y);
        merges[k] = {x, y, cur, k};
        size[y] += size[x];
        size[x] = 0;
        for (int i = 0; i < n; ++i) {
            if (size[i] == 0 || i == y) continue;
            const double v = std::max(D[(size_t)i * n + x], D[(size_t)i * n + y]);
            D[(size_t)i * n + y] = v;
            D[(size_t)y * n + i] = v;
        }
    }
    std::stable_sort(merges.begin(), merges.end(), [](const Merge& a, const Merge& b) { return a.h < b.h; });
    std::vector<int> parent(2 * n - 1);
    std::iota(parent.begin(), parent.end(), 0);
    auto find = [&](int a) {
        int r = a;
        while (parent[r] != r) r = parent[r];
        while (parent[a] != r) { const int nx = parent[a]; parent[a] = r; a = nx; }
        return r;
    };
    std::vector<int> child0(n - 1), child1(n - 1);
    for (int i = 0; i < n - 1; ++i) {
        const int xr = find(merges[i].x), yr = find(merges[i].y);
        child0[i] = std::min(xr, yr);
        child1[i] = std::max(xr, yr);
        parent[xr] = parent[yr] = n + i;
    }
    for (int i = 0; i < n; ++i) labels[i] = 0;
    std::vector<int> stack{child0[n - 2]};          // the root's child with the smaller node id is cluster 1
    while (!stack.empty()) {
        const int a = stack.back();
        stack.pop_back();
        if (a < n) labels[a] = 1;
        else { stack.push_back(child0[a - n]); stack.push_back(child1[a - n]); }
    }
}

// Multi-threaded mini-batch assembly into (pinned) host buffers: for client c and slot j copy
// sample index idx[c*per + j] of that client's array (base pointer src[c]) -- the host side of the
// one-H2D-copy-per-round input path (datasets/dataset.py: get_train_batches).
void bl_gather_batches(const void* const* src_x, const int64_t* const* src_y, const int64_t* idx, int n_clients,
                       int per_client, int64_t sample_bytes, void* dst_x, int64_t* dst_y) {
    const int64_t total = (int64_t)n_clients * per_client;
    auto work = [&](int64_t lo, int64_t hi) {
        for (int64_t t = lo; t < hi; ++t) {
            const int c = (int)(t / per_client);
            const int64_t s = idx[t];
            std::memcpy((char*)dst_x + t * sample_bytes, (const char*)src_x[c] + s * sample_bytes,
                        (size_t)sample_bytes);
            dst_y[t] = src_y[c][s];
        }
    };
    unsigned nt = std::thread::hardware_concurrency();
    nt = std::max(1u, std::min(nt, 8u));
    if (total * sample_bytes < (1 << 20) || nt == 1) { work(0, total); return; }
    std::vector<std::thread> pool;
    const int64_t step = (total + nt - 1) / nt;
    for (unsigned i = 0; i < nt; ++i) {
        const int64_t lo = i * step, hi = std::min<int64_t>(total, lo + step);
        if (lo < hi) pool.emplace_back(work, lo, hi);
    }
    for (auto& th : pool) th.join();
}

}  // extern "C"
