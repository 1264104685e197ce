// Host-side execution of the product's thread-path star code (csrc/star.cuh, csrc/predicates.cuh):
// the very same source the CUDA kernel compiles, run sequentially on the CPU so the Delaunay logic
// can be checked where no GPU exists.  TEST INFRASTRUCTURE -- built by tests/host_sim/build.sh.
#include <stdint.h>
#include <stdlib.h>
#include <math.h>
#include <vector>
#include <algorithm>
#include "../../mvoscalerecovery_b200/csrc/star.cuh"

using namespace mvosr;

struct HostStar {
    uint16_t v[64];
    int get(int i) const { return v[i]; }
    void set(int i, int x) { v[i] = (uint16_t)x; }
};

// mirrors build_grid() of frame_kernel.cuh (sequential)
static void host_grid(int n, const double *px, const double *py, int cap, Grid &g, std::vector<uint16_t> &cell_start,
                      std::vector<uint16_t> &cell_n, std::vector<uint16_t> &cell_pts, std::vector<uint8_t> &dup) {
    float xmn = INFINITY, xmx = -INFINITY, ymn = INFINITY, ymx = -INFINITY;
    for (int i = 0; i < n; ++i) { xmn = fminf(xmn, (float)px[i]); xmx = fmaxf(xmx, (float)px[i]); ymn = fminf(ymn, (float)py[i]); ymx = fmaxf(ymx, (float)py[i]); }
    double w = (double)xmx - xmn, hgt = (double)ymx - ymn, h;
    if (w > 0 && hgt > 0) h = sqrt(2.0 * w * hgt / n); else h = fmax(w, hgt) * 2.0 / n;
    h = fmax(h, fmax(sqrt(w * hgt / cap), fmax(w, hgt) / cap));
    if (!(h > 0)) h = 1.0;
    int gx, gy;
    for (;;) { double a = floor(w / h) + 1, b = floor(hgt / h) + 1; if (a * b <= cap) { gx = (int)a; gy = (int)b; break; } h *= 1.25; }
    g.xmin = xmn; g.ymin = ymn; g.h = h; g.inv_h = 1.0 / h; g.gx = gx; g.gy = gy;
    int nc = gx * gy;
    std::vector<std::vector<int>> cells(nc);
    for (int i = 0; i < n; ++i) cells[cell_coord(px[i], xmn, g.inv_h, gx) + gx * cell_coord(py[i], ymn, g.inv_h, gy)].push_back(i);
    cell_start.assign(nc + 1, 0); cell_n.assign(nc, 0); cell_pts.assign(n, INF16); dup.assign(n, 0);
    int o = 0;
    for (int c = 0; c < nc; ++c) {
        cell_start[c] = (uint16_t)o;
        int m = 0;
        for (int s : cells[c]) {
            bool d = false;
            for (int j = 0; j < m; ++j) { int q = cell_pts[o + j]; if (px[q] == px[s] && py[q] == py[s]) { d = true; break; } }
            if (d) dup[s] = 1; else cell_pts[o + m++] = (uint16_t)s;
        }
        cell_n[c] = (uint16_t)m;
        o += (int)cells[c].size();
    }
    cell_start[nc] = (uint16_t)o;
}

// Canonical Delaunay triangles via the product's thread-path stars. defer_cells<0 disables deferral.
extern "C" int host_sim_delaunay(const float *fx, const float *fy, int n, int cap, int32_t *tri_out, int *n_defer_out, int *n_exact_out) {
    std::vector<double> vx(fx, fx + n), vy(fy, fy + n);
    const double *px = vx.data(), *py = vy.data();
    Grid g; std::vector<uint16_t> cs, cn, cp; std::vector<uint8_t> dup;
    host_grid(n, px, py, cap, g, cs, cn, cp, dup);
    PointSet ps; ps.px = px; ps.py = py; ps.cell_start = cs.data(); ps.cell_n = cn.data(); ps.cell_pts = cp.data(); ps.g = g;
    int T = 0, n_exact = 0, n_defer = 0;
    for (int p = 0; p < n; ++p) {
        if (dup[p]) continue;
        HostStar st; int d = 0;
        int r = build_star_thread(st, d, p, ps, n_exact);
        if (r == STAR_DEFER) {
            // the device hands these to the warp path; here: count them and finish with an unbounded thread build
            ++n_defer;
            r = build_star_thread_t<64, false>(st, d, p, ps, n_exact);
        }
        if (r == STAR_NONE) continue;
        if (r != STAR_OK) return -100 - r;
        std::vector<std::pair<int, int>> loc;
        for (int i = 0; i < d; ++i) {
            int qa = st.get(i), qb = st.get(i + 1 < d ? i + 1 : 0);
            if (qa == INF16 || qb == INF16) continue;
            if (qa > p && qb > p) loc.push_back({ std::min(qa, qb), std::max(qa, qb) });
        }
        std::sort(loc.begin(), loc.end());
        for (auto &e : loc) { tri_out[3 * T] = p; tri_out[3 * T + 1] = e.first; tri_out[3 * T + 2] = e.second; ++T; }
    }
    if (n_defer_out) *n_defer_out = n_defer;
    if (n_exact_out) *n_exact_out = n_exact;
    return T;
}
