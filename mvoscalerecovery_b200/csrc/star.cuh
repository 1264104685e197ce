// Per-site Delaunay stars over a shared-memory point set (device code, sm_100a).
//
// Replaces scipy.spatial.Delaunay(feature2d).simplices (Qhull) at reference
// src/rescale.py:124-125,136-137, per frame, for a point set staged in shared memory.
//
// Method (B200-first, no shared mutable mesh): the Delaunay star of every point p -- the cyclic
// counter-clockwise list of its neighbours -- is built independently by incremental insertion:
// candidates come from a uniform grid in a rectangle that grows around p's cell, each candidate is
// tested for conflict against the triangles of the current star (exact in-circle with symbolic
// tie-break; ghost triangles (p,q,inf)/(p,inf,q) on the hull use exact orientation), the arc of
// conflicting triangles is removed and the candidate inserted (Bowyer-Watson restricted to one
// star).  The star is final once every unexamined point is farther from p than twice the largest
// circumradius of the star (it cannot cut p's Voronoi cell any more); hull points examine all.
// Because predicates are exact and ties are broken symbolically the triangulation is unique, so
// independent stars agree with each other; triangle (a<b<c) is emitted by the star of a.
//
// Two execution shapes share the same rules:
//   thread path: one thread per point, star ids in shared memory (interleaved by thread);
//   warp path  : one warp per point for the stars the thread path gives up on (hull / near-hull
//                points whose search rectangle keeps growing): star cached with coordinates in
//                per-warp shared memory, lanes pre-filter 32 candidates at a time and evaluate the
//                conflict of one inserted candidate against all star triangles in parallel.
#pragma once
#include <stdint.h>
#include "predicates.cuh"

namespace mvosr {

constexpr int NT = 512;                 // threads per CTA of the fused frame kernel
constexpr int NWARP = NT / 32;
constexpr int MAXDEG_T = 24;            // star capacity, thread path (overflow -> warp path)
constexpr int MAXDEG_W = 32;            // star capacity, warp path (overflow -> frame status OVERFLOW)
constexpr int DEFER_CELLS = 30;         // thread path gives up after examining this many cells
constexpr uint16_t INF16 = 0xFFFF;

struct Grid {
    float xmin, ymin;
    double inv_h, h;
    int gx, gy;
};

struct PointSet {
    const float *px, *py;               // shared memory
    const uint16_t *cell_start;         // [gx*gy+1]
    const uint16_t *cell_n;             // [gx*gy] non-duplicate points in the cell
    const uint16_t *cell_pts;           // ids sorted by cell (duplicates moved to the cell tail as INF16)
    Grid g;
};

MVOSR_HD int cell_coord(double x, double xmin, double inv_h, int gmax) {
    int c = (int)((x - xmin) * inv_h);
    return c < 0 ? 0 : (c >= gmax ? gmax - 1 : c);
}

// result codes of a star build
enum { STAR_OK = 0, STAR_DEFER = 1, STAR_OVERFLOW = 2, STAR_INCONSISTENT = 3, STAR_NONE = 4 };

// -------------------------------------------------------------------------------------------------
// thread path
// -------------------------------------------------------------------------------------------------
struct ThreadStar {
    uint16_t *base;                      // star ids: slot i at base[i*NT]
    MVOSR_HD int get(int i) const { return base[i * NT]; }
    MVOSR_HD void set(int i, int v) { base[i * NT] = (uint16_t)v; }
};

struct Bootstrap { int qpos, qneg; };

// Candidate s against an empty star (collinear bootstrap). Returns true when the star got its
// first real triangle.
template <class Star>
MVOSR_HD bool star_bootstrap(Star &st, int &d, Bootstrap &bs, int s, double ppx, double ppy,
                                               const float *px, const float *py, int &n_exact) {
    if (bs.qpos < 0) { bs.qpos = s; return false; }
    double qx = (double)px[bs.qpos] - ppx, qy = (double)py[bs.qpos] - ppy;
    double sx = (double)px[s] - ppx, sy = (double)py[s] - ppy;
    int o = cross_sign(qx, qy, sx, sy, n_exact);
    if (o == 0) {
        // same line through p: keep only the nearest point on each side (exactly collinear => compare along the dominant axis)
        bool same_side = (fabs(qx) >= fabs(qy)) ? ((sx > 0) == (qx > 0)) : ((sy > 0) == (qy > 0));
        if (same_side) {
            if (fabs(sx) + fabs(sy) < fabs(qx) + fabs(qy)) bs.qpos = s;
        } else if (bs.qneg < 0) {
            bs.qneg = s;
        } else {
            double nx = (double)px[bs.qneg] - ppx, ny = (double)py[bs.qneg] - ppy;
            if (fabs(sx) + fabs(sy) < fabs(nx) + fabs(ny)) bs.qneg = s;
        }
        return false;
    }
    d = 0;
    if (o > 0) {            // s left of p->qpos
        st.set(d++, bs.qpos); st.set(d++, s); if (bs.qneg >= 0) st.set(d++, bs.qneg); st.set(d++, INF16);
    } else {
        if (bs.qneg >= 0) st.set(d++, bs.qneg);
        st.set(d++, s); st.set(d++, bs.qpos); st.set(d++, INF16);
    }
    return true;
}

// Arc bookkeeping shared by both paths: given the conflict mask over d triangles returns the
// start i0 and length len of the (single, cyclic) run of set bits. false if not a single run.
MVOSR_HD bool conflict_arc(uint32_t cf, int d, int &i0, int &len) {
    uint32_t full = d >= 32 ? 0xFFFFFFFFu : ((1u << d) - 1u);
    uint32_t prev = ((cf << 1) | (cf >> (d - 1))) & full;        // bit i = cf[i-1]
    uint32_t starts = cf & ~prev;
    if (mvosr_popc(starts) != 1) return false;
    i0 = mvosr_ffs(starts) - 1;
    len = mvosr_popc(cf);
    // the run must be contiguous: rotating cf right by i0 gives the low `len` bits set
    uint32_t rot = i0 ? (((cf >> i0) | (cf << (d - i0))) & full) : cf;
    return rot == (len >= 32 ? 0xFFFFFFFFu : ((1u << len) - 1u));
}

// Insert candidate s into a non-empty star (thread path).
template <int MAXD, class Star>
MVOSR_HD int star_insert(Star &st, int &d, int p, double ppx, double ppy, int s,
                                           const float *px, const float *py, int &n_exact) {
    double sx = (double)px[s] - ppx, sy = (double)py[s] - ppy, sl = sx * sx + sy * sy;
    uint32_t cf = 0;
    int q_first = st.get(0);
    double fx_ = 0, fy_ = 0, fl_ = 0;
    if (q_first != INF16) { fx_ = (double)px[q_first] - ppx; fy_ = (double)py[q_first] - ppy; fl_ = fx_ * fx_ + fy_ * fy_; }
    int qa = q_first; double ax = fx_, ay = fy_, al = fl_;
    for (int i = 0; i < d; ++i) {
        int qb; double bx, by, bl;
        if (i + 1 < d) {
            qb = st.get(i + 1);
            if (qb != INF16) { bx = (double)px[qb] - ppx; by = (double)py[qb] - ppy; bl = bx * bx + by * by; }
            else { bx = by = bl = 0; }
        } else { qb = q_first; bx = fx_; by = fy_; bl = fl_; }
        bool c;
        if (qb == INF16) {               // ghost (p, qa, inf): outside lies LEFT of p->qa
            int o = cross_sign(ax, ay, sx, sy, n_exact);
            c = o > 0 || (o == 0 && strictly_between(ax, ay, sx, sy));
        } else if (qa == INF16) {        // ghost (p, inf, qb): outside lies RIGHT of p->qb
            int o = cross_sign(bx, by, sx, sy, n_exact);
            c = o < 0 || (o == 0 && strictly_between(bx, by, sx, sy));
        } else {
            c = incircle_sos(ax, ay, al, bx, by, bl, sx, sy, sl, p, qa, qb, s, n_exact);
        }
        cf |= (uint32_t)c << i;
        qa = qb; ax = bx; ay = by; al = bl;
    }
    if (!cf) return STAR_OK;
    int i0, len;
    if (!conflict_arc(cf, d, i0, len) || len >= d) return STAR_INCONSISTENT;
    int nd = d - len + 2;
    if (nd > MAXD) return STAR_OVERFLOW;
    // neighbours i0+1 .. i0+len-1 (cyclic) disappear, s goes in right after slot i0
    int e = i0 + len - d;                 // > 0: the removed run wraps past the end by e slots
    if (e > 0) {
        int keep = i0 - e + 1;            // old[e..i0] -> new[0..keep-1]
        for (int k = 0; k < keep; ++k) st.set(k, st.get(k + e));
        st.set(keep, s);
    } else if (len == 1) {
        for (int k = d; k > i0 + 1; --k) st.set(k, st.get(k - 1));
        st.set(i0 + 1, s);
    } else {
        st.set(i0 + 1, s);
        int shift = len - 2;
        if (shift > 0) for (int k = i0 + 2; k + shift < d; ++k) st.set(k, st.get(k + shift));
    }
    d = nd;
    return STAR_OK;
}

// Is the star closed (no ghost) and every unexamined point provably unable to cut it?
// margin = lower bound on the distance from p to any unexamined point.
template <class Star>
MVOSR_HD bool star_final(const Star &st, int d, double ppx, double ppy, double margin,
                                           const float *px, const float *py) {
    if (d < 3 || !(margin > 0)) return false;
    double m2 = margin * margin;
    int q_first = st.get(0);
    if (q_first == INF16) return false;
    double fx_ = (double)px[q_first] - ppx, fy_ = (double)py[q_first] - ppy;
    double ax = fx_, ay = fy_;
    for (int i = 0; i < d; ++i) {
        double bx, by;
        if (i + 1 < d) {
            int qb = st.get(i + 1);
            if (qb == INF16) return false;
            bx = (double)px[qb] - ppx; by = (double)py[qb] - ppy;
        } else { bx = fx_; by = fy_; }
        // 4 R^2 = |a|^2 |b|^2 |a-b|^2 / w^2 ; need margin^2 >= 4 R^2 with slack for the rounding of w
        double l = ax * by, r = ay * bx;
        double w = (l - r) - 4.0e-16 * (fabs(l) + fabs(r));
        if (!(w > 0)) return false;
        double ex = ax - bx, ey = ay - by;
        double rhs = (ax * ax + ay * ay) * (bx * bx + by * by) * (ex * ex + ey * ey);
        if (!(m2 * w * w >= rhs * (1.0 + 1.0e-9))) return false;
        ax = bx; ay = by;
    }
    return true;
}

// Rectangle of examined cells around p and the distance bound it implies.
struct Rect {
    int x0, x1, y0, y1;                  // inclusive cell ranges
    MVOSR_HD double margin(const Grid &g, double ppx, double ppy, int &side) const {
        const double INFD = 1.0e300;
        double ml = x0 > 0 ? ppx - ((double)g.xmin + x0 * g.h) : INFD;
        double mr = x1 < g.gx - 1 ? ((double)g.xmin + (x1 + 1) * g.h) - ppx : INFD;
        double mb = y0 > 0 ? ppy - ((double)g.ymin + y0 * g.h) : INFD;
        double mt = y1 < g.gy - 1 ? ((double)g.ymin + (y1 + 1) * g.h) - ppy : INFD;
        double m = ml; side = 0;
        if (mr < m) { m = mr; side = 1; }
        if (mb < m) { m = mb; side = 2; }
        if (mt < m) { m = mt; side = 3; }
        if (m >= INFD) side = -1;        // the rectangle covers the whole grid
        return m - 1.0e-6;               // slack for the rounding in cell assignment
    }
};

// Build the star of p, thread path.  On STAR_OK the star is in st[0..d).
template <int MAXD, bool CAN_DEFER, class Star>
MVOSR_HD int build_star_thread_t(Star &st, int &d, int p, const PointSet &ps, int &n_exact) {
    const float *px = ps.px, *py = ps.py;
    const Grid &g = ps.g;
    double ppx = px[p], ppy = py[p];
    Rect rc;
    rc.x0 = rc.x1 = cell_coord(ppx, g.xmin, g.inv_h, g.gx);
    rc.y0 = rc.y1 = cell_coord(ppy, g.ymin, g.inv_h, g.gy);
    d = 0;
    Bootstrap bs; bs.qpos = bs.qneg = -1;
    int cells = 0;
    int ax0 = rc.x0, ax1 = rc.x1, ay0 = rc.y0, ay1 = rc.y1;     // cells to examine this round
    for (;;) {
        for (int cy = ay0; cy <= ay1; ++cy) {
            for (int cx = ax0; cx <= ax1; ++cx) {
                int c = cy * g.gx + cx;
                int beg = ps.cell_start[c], n = ps.cell_n[c];
                for (int k = 0; k < n; ++k) {
                    int s = ps.cell_pts[beg + k];
                    if (s == p) continue;
                    if (d == 0) { star_bootstrap(st, d, bs, s, ppx, ppy, px, py, n_exact); continue; }
                    int r = star_insert<MAXD>(st, d, p, ppx, ppy, s, px, py, n_exact);
                    if (r == STAR_OVERFLOW && CAN_DEFER) return STAR_DEFER;
                    if (r != STAR_OK) return r;
                }
            }
            cells += ax1 - ax0 + 1;
        }
        int side;
        double m = rc.margin(g, ppx, ppy, side);
        if (side < 0) return d > 0 ? STAR_OK : STAR_NONE;
        if (star_final(st, d, ppx, ppy, m, px, py)) return STAR_OK;
        if (CAN_DEFER && cells > DEFER_CELLS) return STAR_DEFER;
        if (side == 0) { --rc.x0; ax0 = ax1 = rc.x0; ay0 = rc.y0; ay1 = rc.y1; }
        else if (side == 1) { ++rc.x1; ax0 = ax1 = rc.x1; ay0 = rc.y0; ay1 = rc.y1; }
        else if (side == 2) { --rc.y0; ay0 = ay1 = rc.y0; ax0 = rc.x0; ax1 = rc.x1; }
        else { ++rc.y1; ay0 = ay1 = rc.y1; ax0 = rc.x0; ax1 = rc.x1; }
    }
}

template <class Star>
MVOSR_HD int build_star_thread(Star &st, int &d, int p, const PointSet &ps, int &n_exact) {
    return build_star_thread_t<MAXDEG_T, true>(st, d, p, ps, n_exact);
}

#if defined(__CUDACC__)
// -------------------------------------------------------------------------------------------------
// warp path
// -------------------------------------------------------------------------------------------------
struct WarpStar {                        // per-warp shared memory, MAXDEG_W slots
    uint16_t *id; double *qx, *qy, *ql;
    __device__ __forceinline__ int get(int i) const { return id[i]; }
    __device__ __forceinline__ void set(int i, int v) { id[i] = (uint16_t)v; }
};

// Lane-parallel: might candidate (sx,sy) conflict with any triangle of the cached star? (float64 filter;
// "undecided" counts as yes).  Every lane tests its own candidate against all d triangles.
__device__ __forceinline__ bool warp_prefilter(const WarpStar &ws, int d, double sx, double sy) {
    double sl = sx * sx + sy * sy;
    for (int i = 0; i < d; ++i) {
        int j = i + 1 < d ? i + 1 : 0;
        int qa = ws.id[i], qb = ws.id[j];
        if (qb == INF16) {
            double l = ws.qx[i] * sy, r = ws.qy[i] * sx;
            if (l - r >= -3.4e-16 * (fabs(l) + fabs(r))) return true;
        } else if (qa == INF16) {
            double l = ws.qx[j] * sy, r = ws.qy[j] * sx;
            if (l - r <= 3.4e-16 * (fabs(l) + fabs(r))) return true;
        } else {
            int f = det3_lift_sign_filter(ws.qx[i], ws.qy[i], ws.ql[i], ws.qx[j], ws.qy[j], ws.ql[j], sx, sy, sl);
            if (f != 1) return true;     // det3 < 0 (inside) or undecided
        }
    }
    return false;
}

// Warp-uniform insertion of candidate s: lane i evaluates triangle i, the arc is removed cooperatively.
__device__ __forceinline__ int warp_insert(WarpStar &ws, int &d, int p, double ppx, double ppy, int s,
                                           const float *px, const float *py, int lane, int &n_exact) {
    double sx = (double)px[s] - ppx, sy = (double)py[s] - ppy, sl = sx * sx + sy * sy;
    bool c = false;
    if (lane < d) {
        int j = lane + 1 < d ? lane + 1 : 0;
        int qa = ws.id[lane], qb = ws.id[j];
        if (qb == INF16) {
            int o = cross_sign(ws.qx[lane], ws.qy[lane], sx, sy, n_exact);
            c = o > 0 || (o == 0 && strictly_between(ws.qx[lane], ws.qy[lane], sx, sy));
        } else if (qa == INF16) {
            int o = cross_sign(ws.qx[j], ws.qy[j], sx, sy, n_exact);
            c = o < 0 || (o == 0 && strictly_between(ws.qx[j], ws.qy[j], sx, sy));
        } else {
            c = incircle_sos(ws.qx[lane], ws.qy[lane], ws.ql[lane], ws.qx[j], ws.qy[j], ws.ql[j], sx, sy, sl,
                             p, qa, qb, s, n_exact);
        }
    }
    uint32_t cf = __ballot_sync(0xFFFFFFFFu, c);
    if (!cf) return STAR_OK;
    int i0, len;
    if (!conflict_arc(cf, d, i0, len) || len >= d) return STAR_INCONSISTENT;
    int nd = d - len + 2;
    if (nd > MAXDEG_W) return STAR_OVERFLOW;
    // new[0] = s ; new[k] = old[(i0+len+k-1) % d], k = 1..nd-1
    int vid = 0; double vx = 0, vy = 0, vl = 0;
    if (lane == 0) { vid = s; vx = sx; vy = sy; vl = sl; }
    else if (lane < nd) {
        int src = (i0 + len + lane - 1) % d;
        vid = ws.id[src]; vx = ws.qx[src]; vy = ws.qy[src]; vl = ws.ql[src];
    }
    __syncwarp();
    if (lane < nd) { ws.id[lane] = (uint16_t)vid; ws.qx[lane] = vx; ws.qy[lane] = vy; ws.ql[lane] = vl; }
    __syncwarp();
    d = nd;
    return STAR_OK;
}

// lane-parallel version of star_final
__device__ __forceinline__ bool warp_star_final(const WarpStar &ws, int d, double margin, int lane) {
    if (d < 3 || !(margin > 0)) return false;
    bool ok = true;
    if (lane < d) {
        int j = lane + 1 < d ? lane + 1 : 0;
        if (ws.id[lane] == INF16) ok = false;
        else if (ws.id[j] == INF16) ok = false;
        else {
            double ax = ws.qx[lane], ay = ws.qy[lane], bx = ws.qx[j], by = ws.qy[j];
            double l = ax * by, r = ay * bx;
            double w = (l - r) - 4.0e-16 * (fabs(l) + fabs(r));
            double ex = ax - bx, ey = ay - by;
            double rhs = ws.ql[lane] * ws.ql[j] * (ex * ex + ey * ey);
            ok = (w > 0) && (margin * margin * w * w >= rhs * (1.0 + 1.0e-9));
        }
    }
    return __all_sync(0xFFFFFFFFu, ok);
}

// Process one batch of candidates held one per lane (s < 0: none) against the warp star.
__device__ __forceinline__ int warp_candidates(WarpStar &ws, int &d, Bootstrap &bs, int p, double ppx, double ppy, int s_lane,
                                               const float *px, const float *py, int lane, int &n_exact) {
    bool flag = false;
    if (s_lane >= 0 && s_lane != p) {
        if (d == 0) flag = true;
        else flag = warp_prefilter(ws, d, (double)px[s_lane] - ppx, (double)py[s_lane] - ppy);
    }
    uint32_t m = __ballot_sync(0xFFFFFFFFu, flag);
    while (m) {
        int l = __ffs(m) - 1; m &= m - 1;
        int s = __shfl_sync(0xFFFFFFFFu, s_lane, l);
        if (d == 0) {
            // warp-uniform bootstrap (every lane computes the same thing; lane 0 owns the writes)
            int dd = 0; Bootstrap b2 = bs;
            struct Tmp { uint16_t v[4]; __device__ void set(int i, int x) { v[i] = (uint16_t)x; } } tmp;
            bool got = star_bootstrap(tmp, dd, b2, s, ppx, ppy, px, py, n_exact);
            bs = b2;
            if (got) {
                if (lane < dd) {
                    int q = tmp.v[lane];
                    ws.id[lane] = (uint16_t)q;
                    if (q != INF16) {
                        double qx = (double)px[q] - ppx, qy = (double)py[q] - ppy;
                        ws.qx[lane] = qx; ws.qy[lane] = qy; ws.ql[lane] = qx * qx + qy * qy;
                    } else { ws.qx[lane] = ws.qy[lane] = ws.ql[lane] = 0; }
                }
                __syncwarp();
                d = dd;
            }
            continue;
        }
        int r = warp_insert(ws, d, p, ppx, ppy, s, px, py, lane, n_exact);
        if (r != STAR_OK) return r;
    }
    return STAR_OK;
}

// Build the star of p with the whole warp.  Result in ws[0..d).
__device__ __forceinline__ int build_star_warp(WarpStar &ws, int &d, int p, const PointSet &ps, int lane, int &n_exact) {
    const float *px = ps.px, *py = ps.py;
    const Grid &g = ps.g;
    double ppx = px[p], ppy = py[p];
    Rect rc;
    rc.x0 = rc.x1 = cell_coord(ppx, g.xmin, g.inv_h, g.gx);
    rc.y0 = rc.y1 = cell_coord(ppy, g.ymin, g.inv_h, g.gy);
    d = 0;
    Bootstrap bs; bs.qpos = bs.qneg = -1;
    int ax0 = rc.x0, ax1 = rc.x1, ay0 = rc.y0, ay1 = rc.y1;
    for (;;) {
        if (ax0 == ax1) {
            // a column of cells (or the single start cell): lanes own rows, walk the cells in lockstep
            for (int yb = ay0; yb <= ay1; yb += 32) {
                int cy = yb + lane;
                int beg = 0, n = 0;
                if (cy <= ay1) { int c = cy * g.gx + ax0; beg = ps.cell_start[c]; n = ps.cell_n[c]; }
                int nmax = n;
                for (int o = 16; o; o >>= 1) nmax = max(nmax, __shfl_xor_sync(0xFFFFFFFFu, nmax, o));
                for (int k = 0; k < nmax; ++k) {
                    int s = k < n ? (int)ps.cell_pts[beg + k] : -1;
                    int r = warp_candidates(ws, d, bs, p, ppx, ppy, s, px, py, lane, n_exact);
                    if (r != STAR_OK) return r;
                }
            }
        } else {
            // a row of cells: their points are one contiguous run of cell_pts (duplicates are INF16 gaps)
            int c0 = ay0 * g.gx + ax0, c1 = ay0 * g.gx + ax1;
            int beg = ps.cell_start[c0], end = ps.cell_start[c1 + 1];
            for (int k = beg; k < end; k += 32) {
                int s = -1;
                if (k + lane < end) { int v = ps.cell_pts[k + lane]; if (v != INF16) s = v; }
                int r = warp_candidates(ws, d, bs, p, ppx, ppy, s, px, py, lane, n_exact);
                if (r != STAR_OK) return r;
            }
        }
        int side;
        double m = rc.margin(g, ppx, ppy, side);
        if (side < 0) return d > 0 ? STAR_OK : STAR_NONE;
        if (warp_star_final(ws, d, m, lane)) return STAR_OK;
        if (side == 0) { --rc.x0; ax0 = ax1 = rc.x0; ay0 = rc.y0; ay1 = rc.y1; }
        else if (side == 1) { ++rc.x1; ax0 = ax1 = rc.x1; ay0 = rc.y0; ay1 = rc.y1; }
        else if (side == 2) { --rc.y0; ay0 = ay1 = rc.y0; ax0 = rc.x0; ax1 = rc.x1; }
        else { ++rc.y1; ay0 = ay1 = rc.y1; ax0 = rc.x0; ax1 = rc.x1; }
    }
}

#endif  // __CUDACC__

}  // namespace mvosr
