// Host build of the per-lane Delaunay star builder (csrc/gthread.cuh, __host__ __device__): the same source the frame kernel runs,
// one star per lane, executed here one star after the other over a strip-sorted set built by a sequential restatement of
// build_grid (csrc/frame_kernel.cuh).  tests/test_star_thread_host.py compares every star it certifies with Qhull's.
//   g++ -O2 -std=c++17 -shared -fPIC -Wno-unknown-pragmas -DMVOSR_THREAD_COST -o libstar_thread_host.so star_thread_host.cpp
#include <stdint.h>
#include <math.h>
#include <vector>
#include <algorithm>
#include "../../mvoscalerecovery_b200/csrc/gthread.cuh"

using namespace mvosr;

namespace {
struct HostSet {
    std::vector<float> x, y; std::vector<uint16_t> orig, row_start, row_bin, bin_row, row_cell, cell_start; std::vector<float2> row_xi;
    SortedSet ps;
};

void build(HostSet &H, int n, const float *U, const float *V, int cap, float density, int win_m, float wfac) {
    float xmn = INFINITY, xmx = -INFINITY, ymn = INFINITY, ymx = -INFINITY;
    for (int i = 0; i < n; ++i) { xmn = fminf(xmn, U[i]); xmx = fmaxf(xmx, U[i]); ymn = fminf(ymn, V[i]); ymx = fmaxf(ymx, V[i]); }
    const int NB = std::max(1, std::min(128, (cap - 32) / 18));
    float bh = (ymx - ymn) / (float)NB; if (!(bh > 0.f)) bh = 1.f;
    SortedSet &ps = H.ps;
    ps.n = n; ps.NB = NB; ps.R = 0; ps.win_m = win_m; ps.kdens = density; ps.wfac = wfac;
    ps.xmin = xmn; ps.xmax = xmx; ps.ymin = ymn; ps.ymax = ymx; ps.bh = bh; ps.inv_bh = 1.f / bh;
    std::vector<unsigned> cnt(NB, 0); std::vector<float> bmin(NB, INFINITY), bmax(NB, -INFINITY);
    for (int i = 0; i < n; ++i) { const int b = bin_of(ps, V[i]); ++cnt[b]; bmin[b] = fminf(bmin[b], U[i]); bmax[b] = fmaxf(bmax[b], U[i]); }
    H.bin_row.assign(NB, 0); H.row_bin.clear(); H.row_cell.clear(); H.row_xi.clear();
    const float Lmin = 0.2f * (xmx - xmn);
    int R = 0, ncell = 0;
    for (int start = 0; start < NB; ) {
        unsigned c = 0; float lo = INFINITY, hi = -INFINITY; int end = start;
        for (int idx = start; idx < NB; ++idx) {
            c += cnt[idx]; lo = fminf(lo, bmin[idx]); hi = fmaxf(hi, bmax[idx]);
            const float L = c ? fmaxf(hi - lo, Lmin) : Lmin;
            end = idx;
            if (idx == NB - 1 || !(density * L - (float)c * ((float)(idx + 1 - start) * bh) > 0.f)) break;
        }
        for (int q = start; q <= end; ++q) H.bin_row[q] = (uint16_t)R;
        const int nc = (int)(c >> 1) + 1;
        const float x0 = c ? lo : 0.f, ext = c ? hi - lo : 0.f;
        float2 xi; xi.x = x0; xi.y = ext > 0.f ? (float)nc / ext : 0.f;
        H.row_bin.push_back((uint16_t)start); H.row_cell.push_back((uint16_t)ncell); H.row_xi.push_back(xi);
        ncell += nc; ++R; start = end + 1;
    }
    H.row_bin.push_back((uint16_t)NB); H.row_cell.push_back((uint16_t)ncell);
    ps.R = R; ps.bin_row = H.bin_row.data(); ps.row_bin = H.row_bin.data(); ps.row_cell = H.row_cell.data(); ps.row_xi = H.row_xi.data();
    auto subcell = [&](float x, float y) {
        const int r = H.bin_row[bin_of(ps, y)];
        return H.row_cell[r] + strip_cell(H.row_xi[r], H.row_cell[r + 1] - H.row_cell[r], x);
    };
    std::vector<int> sc(n), order(n);
    for (int i = 0; i < n; ++i) { sc[i] = subcell(U[i], V[i]); order[i] = i; }
    std::sort(order.begin(), order.end(), [&](int a, int b) {
        if (sc[a] != sc[b]) return sc[a] < sc[b];
        if (U[a] != U[b]) return U[a] < U[b];
        return a < b; });
    H.x.resize(n); H.y.resize(n); H.orig.resize(n);
    H.cell_start.assign(ncell + 1, 0);
    for (int i = 0; i < n; ++i) ++H.cell_start[sc[i] + 1];
    for (int c = 0; c < ncell; ++c) H.cell_start[c + 1] += H.cell_start[c];
    for (int a = 0; a < n; ++a) { const int i = order[a]; H.x[a] = U[i]; H.y[a] = V[i]; H.orig[a] = (uint16_t)i; }
    H.row_start.resize(R + 1);
    for (int r = 0; r <= R; ++r) H.row_start[r] = H.cell_start[H.row_cell[r]];
    ps.x = H.x.data(); ps.y = H.y.data(); ps.orig = H.orig.data(); ps.row_start = H.row_start.data(); ps.cell_start = H.cell_start.data();
    // exact duplicates: all but the lowest index become holes
    std::vector<char> dup(n, 0);
    for (int a = 0; a < n; ++a) {
        const int b0 = H.row_start[H.bin_row[bin_of(ps, H.y[a])]];
        for (int c = a - 1; c >= b0 && H.x[c] == H.x[a]; --c) if (H.y[c] == H.y[a]) { dup[a] = 1; break; }
    }
    for (int a = 0; a < n; ++a) if (dup[a]) H.orig[a] = INF16;
}

// The index of a SUBSET by filtering (filter_grid of csrc/frame_kernel.cuh, restated sequentially): strips, bins and sub-cells stay,
// the sorted copy is compacted in order, row_start / cell_start become survivor counts.  newidx[old feature] = new index or INF16.
void filter(HostSet &H, const uint16_t *newidx, int n_new) {
    const int n_old = H.ps.n;
    std::vector<uint16_t> before(n_old + 1, 0);
    int run = 0;
    for (int i = 0; i < n_old; ++i) { before[i] = (uint16_t)run; const int o = H.orig[i]; if (o != INF16 && newidx[o] != INF16) ++run; }
    before[n_old] = (uint16_t)run;
    for (auto &c : H.cell_start) c = before[c];
    for (auto &r : H.row_start) r = before[r];
    for (int i = 0; i < n_old; ++i) {
        const int o = H.orig[i];
        if (o == INF16 || newidx[o] == INF16) continue;
        const int pos = before[i];
        H.x[pos] = H.x[i]; H.y[pos] = H.y[i]; H.orig[pos] = newidx[o];
    }
    H.ps.n = n_new;
}
}  // namespace

// status / deg / ring by FEATURE index (ring: 16 feature indices per star, counter-clockwise from the nearest neighbour).
// cost[0..2]: candidate evaluations, steps, stars; cost[3]: sum over groups of 32 consecutive sorted positions of 32 x (max evaluations
// of a lane) -- what a warp executes when its lanes run in lock step; cost[4]: the same for steps; cost[5]: strips R; cost[6]: warp iterations of the step loops in the lock-step model (sum over groups and steps of the largest candidate count among the lanes still walking); cost[7]: largest candidate count of a star.

// keep != nullptr: the index is built over all n points and then FILTERED to the points with keep[i] != 0 (new feature index = rank
// among the kept ones); status / deg / ring are then by NEW feature index.
extern "C" int star_thread_run(int n, const float *u, const float *v, int cap, float density, int win_m, float wfac,
                               int32_t *status, int32_t *deg, int32_t *ring, uint64_t *cost, const uint8_t *keep) {
    HostSet H;
    build(H, n, u, v, cap, density, win_m, wfac);
    if (keep) {
        std::vector<uint16_t> newidx(n, INF16);
        int m = 0;
        for (int i = 0; i < n; ++i) if (keep[i]) newidx[i] = (uint16_t)m++;
        filter(H, newidx.data(), m);
        n = m;
    }
    for (int k = 0; k < 8; ++k) cost[k] = 0;
    cost[5] = (uint64_t)H.ps.R;
    unsigned gmax_e = 0, gmax_s = 0;
    std::vector<unsigned> gM, gS;                      // candidates per scan and steps of the lanes of the current group
    auto flush = [&]() {
        // lock-step model: at step k the warp runs max(M) over the lanes that are still walking
        for (unsigned k = 0;; ++k) {
            unsigned m = 0; for (size_t l = 0; l < gM.size(); ++l) if (gS[l] > k) m = std::max(m, gM[l]);
            if (!m) break;
            cost[6] += m;
        }
        unsigned mm = 0; for (unsigned v : gM) mm = std::max(mm, v);
        cost[7] = std::max<uint64_t>(cost[7], mm);
        gM.clear(); gS.clear();
    };
    for (int p = 0; p < n; ++p) {
        if ((p & 31) == 0) { cost[3] += 32ull * gmax_e; cost[4] += 32ull * gmax_s; gmax_e = gmax_s = 0; flush(); }
        const int o = H.orig[p];
        if (o == INF16) continue;
        TRing r; int d = 0; TCost c = {0, 0, 0};
        const int st = thread_star(H.ps, p, r, d, &c);
        status[o] = st; deg[o] = st == TS_OK ? d : 0;
        if (st == TS_OK) for (int k = 0; k < d; ++k) ring[16 * o + k] = H.orig[tring_get(r, k)];
        cost[0] += c.evals; cost[1] += c.steps; cost[2] += 1;
        gM.push_back(c.steps ? c.evals / c.steps : 0); gS.push_back(c.steps);
        gmax_e = std::max(gmax_e, c.evals); gmax_s = std::max(gmax_s, c.steps);
    }
    cost[3] += 32ull * gmax_e; cost[4] += 32ull * gmax_s; flush();
    return 0;
}
