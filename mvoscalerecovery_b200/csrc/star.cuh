// Per-site Delaunay stars over a shared-memory point set (device code, sm_100a).
//
// Replaces scipy.spatial.Delaunay(feature2d).simplices (Qhull) at reference
// src/rescale.py:124-125,136-137, per frame, for a point set staged in shared memory.
//
// Method (B200-first, no shared mutable mesh): the Delaunay star of every point p -- the cyclic
// counter-clockwise list of its neighbours -- is built independently by incremental insertion.
// Candidates come from a uniform grid in a rectangle that grows around p's cell.  For a candidate s
// the star triangle whose angular sector (seen from p) contains s is located with exact orientation
// tests; s cuts p's Voronoi cell iff it lies in the circumcircle of THAT triangle (the cell vertex of
// a sector is the extreme point of the cell in every direction of the sector), so one exact in-circle
// test (symbolic tie-break) decides; on conflict the arc of conflicting triangles is grown to both
// sides, removed, and s inserted (Bowyer-Watson restricted to one star).  Hull points carry one INF
// marker standing for the two ghost triangles (p,q,inf),(p,inf,q), whose conflict is an exact
// orientation test.  A star is final once every unexamined point is farther from p than twice its
// largest circumradius; hull points examine everything.  Exact predicates + symbolic ties make the
// triangulation unique, so independently built stars agree; triangle (a<b<c) is emitted by star a.
//
// Two execution shapes share the rules:
//   thread path: one thread per point, flattened candidate loop (every iteration of every lane is one
//                candidate), star ids in shared memory interleaved by thread, points fetched dynamically;
//   warp path  : one warp per point for the stars the thread path gives up on (hull / near-hull
//                points): star cached with coordinates and Voronoi vertices in per-warp shared memory,
//                the search rectangle doubles, lanes own grid cells (culled against the star's
//                circumdisks), pre-filter their candidates, and evaluate the exact conflict of an
//                inserted candidate against all star triangles in parallel.
#pragma once
#include <stdint.h>
#include "predicates.cuh"

namespace mvosr {

constexpr int NT = 512;                 // threads per CTA of the fused frame kernel
constexpr int NWARP = NT / 32;
constexpr int MAXDEG_T = 24;            // star capacity, thread path (overflow -> warp path)
constexpr int MAXDEG_W = 32;            // star capacity, warp path (overflow -> frame status OVERFLOW)
constexpr uint16_t INF16 = 0xFFFF;

struct Grid {
    double xmin, ymin;
    double inv_h, h;
    int gx, gy;
};

struct PointSet {
    const double *px, *py;              // shared memory; float32-exact values held as float64
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
                             const double *px, const double *py, int &n_exact) {
    if (bs.qpos < 0) { bs.qpos = s; return false; }
    double qx = px[bs.qpos] - ppx, qy = py[bs.qpos] - ppy;
    double sx = px[s] - ppx, sy = py[s] - ppy;
    int o = cross_sign(qx, qy, sx, sy, n_exact);
    if (o == 0) {
        // same line through p: keep only the nearest point on each side (exactly collinear => compare along the dominant axis)
        bool same_side = (fabs(qx) >= fabs(qy)) ? ((sx > 0) == (qx > 0)) : ((sy > 0) == (qy > 0));
        if (same_side) {
            if (fabs(sx) + fabs(sy) < fabs(qx) + fabs(qy)) bs.qpos = s;
        } else if (bs.qneg < 0) {
            bs.qneg = s;
        } else {
            double nx = px[bs.qneg] - ppx, ny = py[bs.qneg] - ppy;
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

// conflict of star triangle t = (p, slot t, slot t+1) with s, given the orientation masks of s w.r.t. the rays
template <class Star>
MVOSR_HD bool tri_conflict(const Star &st, int d, int t, int p, double ppx, double ppy, int s, double sx, double sy, double sl,
                           uint32_t pos, uint32_t neg, const double *px, const double *py, int &n_exact) {
    int t1 = t + 1 < d ? t + 1 : 0;
    int qa = st.get(t), qb = st.get(t1);
    if (qb == INF16) {               // ghost (p, qa, inf): outside lies LEFT of p->qa
        if ((pos >> t) & 1u) return true;
        if ((neg >> t) & 1u) return false;
        return strictly_between(px[qa] - ppx, py[qa] - ppy, sx, sy);
    }
    if (qa == INF16) {               // ghost (p, inf, qb): outside lies RIGHT of p->qb
        if ((neg >> t1) & 1u) return true;
        if ((pos >> t1) & 1u) return false;
        return strictly_between(px[qb] - ppx, py[qb] - ppy, sx, sy);
    }
    double ax = px[qa] - ppx, ay = py[qa] - ppy, bx = px[qb] - ppx, by = py[qb] - ppy;
    return incircle_sos(ax, ay, ax * ax + ay * ay, bx, by, bx * bx + by * by, sx, sy, sl, p, qa, qb, s, n_exact);
}

// Insert candidate s into a non-empty star.
template <int MAXD, class Star>
MVOSR_HD int star_insert(Star &st, int &d, int p, double ppx, double ppy, int s,
                         const double *px, const double *py, int &n_exact) {
    const double sx = px[s] - ppx, sy = py[s] - ppy, sl = sx * sx + sy * sy;
    // orientation of s w.r.t. every ray p->q_i
    uint32_t pos = 0, neg = 0;
    int inf_slot = -1;
    for (int i = 0; i < d; ++i) {
        int q = st.get(i);
        if (q == INF16) { inf_slot = i; continue; }
        int o = cross_sign(px[q] - ppx, py[q] - ppy, sx, sy, n_exact);
        pos |= (uint32_t)(o > 0) << i;
        neg |= (uint32_t)(o < 0) << i;
    }
    const uint32_t full = d >= 32 ? 0xFFFFFFFFu : ((1u << d) - 1u);
    // sector t (closed) contains s  <=>  o_t >= 0 and o_{t+1} <= 0   (finite triangles span < 180 degrees)
    uint32_t nonneg = ~neg & full, nonpos = ~pos & full;
    uint32_t nonpos_next = ((nonpos >> 1) | (nonpos << (d - 1))) & full;      // bit t = nonpos[t+1]
    uint32_t cand = nonneg & nonpos_next;
    if (inf_slot >= 0) {
        int k = inf_slot ? inf_slot - 1 : d - 1;
        cand |= (1u << inf_slot) | (1u << k);                                  // ghost triangles are tested directly (free)
    }
    // flood fill over the cyclic adjacency from the seed sectors: tests exactly the conflicting arc plus its two
    // bounding triangles, through ONE tri_conflict call site (keeps divergent lanes on the same instructions)
    uint32_t cf = 0, tested = 0, pending = cand;
    while (pending) {
        int t = mvosr_ffs(pending) - 1;
        uint32_t bit = 1u << t;
        pending &= ~bit; tested |= bit;
        if (tri_conflict(st, d, t, p, ppx, ppy, s, sx, sy, sl, pos, neg, px, py, n_exact)) {
            cf |= bit;
            uint32_t nb = (1u << (t + 1 < d ? t + 1 : 0)) | (1u << (t ? t - 1 : d - 1));
            pending |= nb & ~tested;
        }
    }
    if (!cf) return STAR_OK;
    int i0, len;
    {
        uint32_t prevm = ((cf << 1) | (cf >> (d - 1))) & full;
        uint32_t starts = cf & ~prevm;
        if (mvosr_popc(starts) != 1) return STAR_INCONSISTENT;             // all triangles, or more than one arc
        i0 = mvosr_ffs(starts) - 1; len = mvosr_popc(cf);
        uint32_t rot = i0 ? (((cf >> i0) | (cf << (d - i0))) & full) : cf;
        if (rot != (len >= 32 ? 0xFFFFFFFFu : ((1u << len) - 1u)) || len >= d) return STAR_INCONSISTENT;
    }
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

// Upper bound of (2 * largest circumradius)^2 over the star, +inf if the star is open (ghost) or a
// triangle is too flat for float64 to bound.  A point farther than that from p cannot cut the cell.
template <class Star>
MVOSR_HD double star_reach2(const Star &st, int d, double ppx, double ppy, const double *px, const double *py) {
    const double INFD = 1.0e300;
    if (d < 3) return INFD;
    int q_first = st.get(0);
    if (q_first == INF16) return INFD;
    double fx_ = px[q_first] - ppx, fy_ = py[q_first] - ppy;
    double ax = fx_, ay = fy_, worst = 0;
    for (int i = 0; i < d; ++i) {
        double bx, by;
        if (i + 1 < d) {
            int qb = st.get(i + 1);
            if (qb == INF16) return INFD;
            bx = px[qb] - ppx; by = py[qb] - ppy;
        } else { bx = fx_; by = fy_; }
        // 4 R^2 = |a|^2 |b|^2 |a-b|^2 / w^2, with w bounded from below against its rounding error
        double l = ax * by, r = ay * bx;
        double w = (l - r) - 4.0e-16 * (fabs(l) + fabs(r));
        if (!(w > 0)) return INFD;
        double ex = ax - bx, ey = ay - by;
        double v = (ax * ax + ay * ay) * (bx * bx + by * by) * (ex * ex + ey * ey) / (w * w);
        worst = v > worst ? v : worst;
        ax = bx; ay = by;
    }
    return worst * (1.0 + 1.0e-9);
}

// Rectangle of examined cells around p and the distance bound it implies.
struct Rect {
    int x0, x1, y0, y1;                  // inclusive cell ranges
    MVOSR_HD double margin(const Grid &g, double ppx, double ppy, int &side) const {
        const double INFD = 1.0e300;
        double ml = x0 > 0 ? ppx - (g.xmin + x0 * g.h) : INFD;
        double mr = x1 < g.gx - 1 ? (g.xmin + (x1 + 1) * g.h) - ppx : INFD;
        double mb = y0 > 0 ? ppy - (g.ymin + y0 * g.h) : INFD;
        double mt = y1 < g.gy - 1 ? (g.ymin + (y1 + 1) * g.h) - ppy : INFD;
        double m = ml; side = 0;
        if (mr < m) { m = mr; side = 1; }
        if (mb < m) { m = mb; side = 2; }
        if (mt < m) { m = mt; side = 3; }
        if (m >= INFD) side = -1;        // the rectangle covers the whole grid
        return m - 1.0e-6;               // slack for the rounding in cell assignment
    }
};

// -------------------------------------------------------------------------------------------------
// thread path: resumable per-point state so that one loop iteration == one candidate for every lane
// -------------------------------------------------------------------------------------------------
// Candidates are visited ring by ring around p's cell: level 1 = the 3x3 block, level r >= 2 = the
// perimeter of the (2r+1)^2 block.  The termination test runs once per completed level.
constexpr int RING_DEFER = 2;            // thread path hands the point to the warp path after this level (5x5 cells)
enum { WORK_HAVE = -1, WORK_RING_END = -2, WORK_CONTINUE = -3 };

struct StarWork {
    int p, d;
    double ppx, ppy, reach2;
    int pcx, pcy;                        // p's cell
    int r, j, jn;                        // ring level, index inside the level, cells in the level
    int k, beg, n;                       // cursor inside the current cell
    Bootstrap bs;
};

MVOSR_HD void work_begin(StarWork &w, int p, const PointSet &ps) {
    w.p = p; w.d = 0; w.reach2 = 1.0e300;
    w.ppx = ps.px[p]; w.ppy = ps.py[p];
    w.pcx = cell_coord(w.ppx, ps.g.xmin, ps.g.inv_h, ps.g.gx);
    w.pcy = cell_coord(w.ppy, ps.g.ymin, ps.g.inv_h, ps.g.gy);
    w.r = 1; w.j = -1; w.jn = 9; w.k = 0; w.n = 0; w.beg = 0;
    w.bs.qpos = w.bs.qneg = -1;
}

// cell j of ring level r around (pcx,pcy); false if outside the grid
MVOSR_HD bool ring_cell(int pcx, int pcy, int r, int j, int gx, int gy, int &cx, int &cy) {
    if (r == 1) { cx = pcx - 1 + j % 3; cy = pcy - 1 + j / 3; }
    else {
        int w = 2 * r + 1;
        if (j < w) { cx = pcx - r + j; cy = pcy - r; }
        else if (j < 2 * w) { cx = pcx - r + (j - w); cy = pcy + r; }
        else if (j < 2 * w + (w - 2)) { cx = pcx - r; cy = pcy - r + 1 + (j - 2 * w); }
        else { cx = pcx + r; cy = pcy - r + 1 + (j - 2 * w - (w - 2)); }
    }
    return cx >= 0 && cx < gx && cy >= 0 && cy < gy;
}

// Advance to the next candidate of the current level: WORK_HAVE (candidate in s) or WORK_RING_END.
MVOSR_HD int work_next(StarWork &w, const PointSet &ps, int &s) {
    for (;;) {
        if (w.k < w.n) { s = ps.cell_pts[w.beg + w.k++]; return WORK_HAVE; }
        if (++w.j >= w.jn) return WORK_RING_END;
        int cx, cy;
        if (!ring_cell(w.pcx, w.pcy, w.r, w.j, ps.g.gx, ps.g.gy, cx, cy)) continue;
        int c = cy * ps.g.gx + cx;
        w.beg = ps.cell_start[c]; w.n = ps.cell_n[c]; w.k = 0;
    }
}

// A level is exhausted: final?  Otherwise open the next level (WORK_CONTINUE) or give up (STAR_DEFER).
template <bool CAN_DEFER, class Star>
MVOSR_HD int work_ring_end(StarWork &w, const Star &st, const PointSet &ps) {
    Rect rc;
    rc.x0 = w.pcx - w.r < 0 ? 0 : w.pcx - w.r; rc.x1 = w.pcx + w.r > ps.g.gx - 1 ? ps.g.gx - 1 : w.pcx + w.r;
    rc.y0 = w.pcy - w.r < 0 ? 0 : w.pcy - w.r; rc.y1 = w.pcy + w.r > ps.g.gy - 1 ? ps.g.gy - 1 : w.pcy + w.r;
    int side;
    double m = rc.margin(ps.g, w.ppx, w.ppy, side);
    if (side < 0) return w.d > 0 ? STAR_OK : STAR_NONE;
    w.reach2 = star_reach2(st, w.d, w.ppx, w.ppy, ps.px, ps.py);
    if (m > 0 && m * m >= w.reach2) return STAR_OK;
    if (CAN_DEFER && w.r >= RING_DEFER) return STAR_DEFER;
    ++w.r; w.j = -1; w.jn = 8 * w.r; w.k = 0; w.n = 0;
    return WORK_CONTINUE;
}

// Feed one candidate.  Returns STAR_OK to continue, anything else ends the point.
template <int MAXD, bool CAN_DEFER, class Star>
MVOSR_HD int work_feed(StarWork &w, Star &st, const PointSet &ps, int s, int &n_exact) {
    if (s == w.p) return STAR_OK;
    if (w.d == 0) { star_bootstrap(st, w.d, w.bs, s, w.ppx, w.ppy, ps.px, ps.py, n_exact); return STAR_OK; }
    double sx = ps.px[s] - w.ppx, sy = ps.py[s] - w.ppy;
    if (sx * sx + sy * sy > w.reach2) return STAR_OK;          // beyond twice the largest circumradius: cannot cut the cell
    int r = star_insert<MAXD>(st, w.d, w.p, w.ppx, w.ppy, s, ps.px, ps.py, n_exact);
    if (r == STAR_OVERFLOW && CAN_DEFER) return STAR_DEFER;
    return r;
}

// Sequential build of one star (used by the host simulation and as the reference semantics of the
// vote-batched device loop in frame_kernel.cuh, which drives the same work_* functions).
template <int MAXD, bool CAN_DEFER, class Star>
MVOSR_HD int build_star_thread_t(Star &st, int &d, int p, const PointSet &ps, int &n_exact) {
    StarWork w;
    work_begin(w, p, ps);
    int rc;
    for (;;) {
        int s;
        rc = work_next(w, ps, s);
        if (rc == WORK_RING_END) {
            rc = work_ring_end<CAN_DEFER>(w, st, ps);
            if (rc == WORK_CONTINUE) continue;
            break;
        }
        rc = work_feed<MAXD, CAN_DEFER>(w, st, ps, s, n_exact);
        if (rc != STAR_OK) break;
    }
    d = w.d;
    return rc;
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
    double *qx, *qy, *ql;                // neighbour i relative to p, and its lift
    double *vx, *vy, *r2;                // Voronoi vertex of triangle i (circumcentre relative to p), radius^2 (inflated); inf if unknown
    uint16_t *id;
    __device__ __forceinline__ int get(int i) const { return id[i]; }
};
constexpr int WARPSTAR_BYTES = MAXDEG_W * (6 * 8 + 2);

// lane i recomputes the Voronoi vertex of triangle i
__device__ __forceinline__ void warp_refresh(WarpStar &ws, int d, int lane) {
    if (lane < d) {
        int j = lane + 1 < d ? lane + 1 : 0;
        double vx = 0, vy = 0, r2 = 1.0e300;
        if (ws.id[lane] != INF16 && ws.id[j] != INF16) {
            double ax = ws.qx[lane], ay = ws.qy[lane], al = ws.ql[lane], bx = ws.qx[j], by = ws.qy[j], bl = ws.ql[j];
            double l = ax * by, r = ay * bx, w = l - r;
            if (w > 1.0e-10 * (fabs(l) + fabs(r))) {
                double inv = 0.5 / w;
                vx = (al * by - bl * ay) * inv; vy = (bl * ax - al * bx) * inv;
                r2 = (vx * vx + vy * vy) * (1.0 + 1.0e-5);
            }
        }
        ws.vx[lane] = vx; ws.vy[lane] = vy; ws.r2[lane] = r2;
    }
    __syncwarp();
}

// Lane-parallel, conservative: might candidate (sx,sy) cut the cell?  ("undecided" counts as yes.)
__device__ __forceinline__ bool warp_prefilter(const WarpStar &ws, int d, double sx, double sy) {
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
            double r2 = ws.r2[i];
            if (r2 >= 1.0e299) return true;
            double dx = sx - ws.vx[i], dy = sy - ws.vy[i];
            if (dx * dx + dy * dy < r2) return true;       // inside the (inflated) circumdisk
        }
    }
    return false;
}

// Lane-parallel, conservative: can the cell rectangle [x0,x1]x[y0,y1] (relative to p) contain a cutting point?
__device__ __forceinline__ bool warp_cell_may_cut(const WarpStar &ws, int d, double x0, double y0, double x1, double y1) {
    for (int i = 0; i < d; ++i) {
        int j = i + 1 < d ? i + 1 : 0;
        int qa = ws.id[i], qb = ws.id[j];
        if (qb == INF16 || qa == INF16) {
            // ghost half-plane: left of p->qa (cross(qa, c) > 0) resp. right of p->qb (cross(qb, c) < 0) for some corner c
            double qx = qb == INF16 ? ws.qx[i] : ws.qx[j], qy = qb == INF16 ? ws.qy[i] : ws.qy[j];
            double sgn = qb == INF16 ? 1.0 : -1.0;
            double c0 = sgn * (qx * y0 - qy * x0), c1 = sgn * (qx * y0 - qy * x1), c2 = sgn * (qx * y1 - qy * x0), c3 = sgn * (qx * y1 - qy * x1);
            double m = fmax(fmax(c0, c1), fmax(c2, c3));
            double tol = 1.0e-9 * (fabs(qx) + fabs(qy)) * (fabs(x0) + fabs(x1) + fabs(y0) + fabs(y1) + 1.0);
            if (m >= -tol) return true;
        } else {
            double r2 = ws.r2[i];
            if (r2 >= 1.0e299) return true;
            double vx = ws.vx[i], vy = ws.vy[i];
            double dx = vx < x0 ? x0 - vx : (vx > x1 ? vx - x1 : 0.0), dy = vy < y0 ? y0 - vy : (vy > y1 ? vy - y1 : 0.0);
            if (dx * dx + dy * dy < r2) return true;
        }
    }
    return false;
}

// Warp-uniform insertion of candidate s: lane i evaluates triangle i exactly, the arc is removed cooperatively.
__device__ __forceinline__ int warp_insert(WarpStar &ws, int &d, int p, double ppx, double ppy, int s,
                                           const double *px, const double *py, int lane, int &n_exact) {
    double sx = px[s] - ppx, sy = py[s] - ppy, sl = sx * sx + sy * sy;
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
    const uint32_t full = d >= 32 ? 0xFFFFFFFFu : ((1u << d) - 1u);
    uint32_t prevm = ((cf << 1) | (cf >> (d - 1))) & full;
    uint32_t starts = cf & ~prevm;
    if (__popc(starts) != 1) return STAR_INCONSISTENT;
    int i0 = __ffs(starts) - 1, len = __popc(cf);
    uint32_t rot = i0 ? (((cf >> i0) | (cf << (d - i0))) & full) : cf;
    if (rot != (len >= 32 ? 0xFFFFFFFFu : ((1u << len) - 1u)) || len >= d) return STAR_INCONSISTENT;
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
    warp_refresh(ws, d, lane);
    return STAR_OK;
}

// lane-parallel (2 * largest circumradius)^2, +inf when open
__device__ __forceinline__ double warp_reach2(const WarpStar &ws, int d, int lane) {
    double v = 0;
    if (d < 3) v = 1.0e300;
    else if (lane < d) {
        int j = lane + 1 < d ? lane + 1 : 0;
        if (ws.id[lane] == INF16 || ws.id[j] == INF16) v = 1.0e300;
        else {
            double ax = ws.qx[lane], ay = ws.qy[lane], bx = ws.qx[j], by = ws.qy[j];
            double l = ax * by, r = ay * bx;
            double w = (l - r) - 4.0e-16 * (fabs(l) + fabs(r));
            double ex = ax - bx, ey = ay - by;
            v = w > 0 ? ws.ql[lane] * ws.ql[j] * (ex * ex + ey * ey) / (w * w) * (1.0 + 1.0e-9) : 1.0e300;
        }
    }
#pragma unroll
    for (int o = 16; o; o >>= 1) v = fmax(v, __shfl_xor_sync(0xFFFFFFFFu, v, o));
    return v;
}

struct WarpTmpStar { uint16_t v[4]; __device__ __forceinline__ void set(int i, int x) { v[i] = (uint16_t)x; } };

// Process one batch of candidates held one per lane (s < 0: none) against the warp star.
__device__ __forceinline__ int warp_candidates(WarpStar &ws, int &d, Bootstrap &bs, int p, double ppx, double ppy, int s_lane,
                                               const double *px, const double *py, int lane, int &n_exact) {
    bool flag = false;
    if (s_lane >= 0 && s_lane != p) {
        if (d == 0) flag = true;
        else flag = warp_prefilter(ws, d, px[s_lane] - ppx, py[s_lane] - ppy);
    }
    uint32_t m = __ballot_sync(0xFFFFFFFFu, flag);
    while (m) {
        int l = __ffs(m) - 1; m &= m - 1;
        int s = __shfl_sync(0xFFFFFFFFu, s_lane, l);
        if (d == 0) {
            // warp-uniform bootstrap (every lane computes the same thing)
            int dd = 0; Bootstrap b2 = bs; WarpTmpStar tmp;
            bool got = star_bootstrap(tmp, dd, b2, s, ppx, ppy, px, py, n_exact);
            bs = b2;
            if (got) {
                if (lane < dd) {
                    int q = tmp.v[lane];
                    ws.id[lane] = (uint16_t)q;
                    if (q != INF16) {
                        double qx = px[q] - ppx, qy = py[q] - ppy;
                        ws.qx[lane] = qx; ws.qy[lane] = qy; ws.ql[lane] = qx * qx + qy * qy;
                    } else { ws.qx[lane] = ws.qy[lane] = ws.ql[lane] = 0; }
                }
                __syncwarp();
                d = dd;
                warp_refresh(ws, d, lane);
            }
            continue;
        }
        int r = warp_insert(ws, d, p, ppx, ppy, s, px, py, lane, n_exact);
        if (r != STAR_OK) return r;
    }
    return STAR_OK;
}

// One chunk of up to 32 cells, one per lane (c < 0: none): cull against the star, then feed the cells' points.
__device__ __forceinline__ int warp_cells(WarpStar &ws, int &d, Bootstrap &bs, int p, double ppx, double ppy, int c,
                                          const PointSet &ps, int lane, int &n_exact) {
    const Grid &g = ps.g;
    int beg = 0, n = 0;
    if (c >= 0) {
        n = ps.cell_n[c];
        if (n && d > 0) {
            int cy = c / g.gx, cx = c - cy * g.gx;
            double x0 = g.xmin + cx * g.h - ppx - 1.0e-6, y0 = g.ymin + cy * g.h - ppy - 1.0e-6;
            if (!warp_cell_may_cut(ws, d, x0, y0, x0 + g.h + 2.0e-6, y0 + g.h + 2.0e-6)) n = 0;
        }
        beg = ps.cell_start[c];
    }
    int nmax = n;
#pragma unroll
    for (int o = 16; o; o >>= 1) nmax = max(nmax, __shfl_xor_sync(0xFFFFFFFFu, nmax, o));
    for (int k = 0; k < nmax; ++k) {
        int s = k < n ? (int)ps.cell_pts[beg + k] : -1;
        int r = warp_candidates(ws, d, bs, p, ppx, ppy, s, ps.px, ps.py, lane, n_exact);
        if (r != STAR_OK) return r;
    }
    return STAR_OK;
}

// Build the star of p with the whole warp.  Result in ws[0..d).
// Phase 1: the 3x3 and 5x5 blocks of cells around p (nearest first; box test for finality after each).
// Phase 2: a directed search.  The union of the star's circumdisks only shrinks as the star is clipped, so the
//          cells that can still matter lie inside R = union of the disks' bounding boxes taken now (the whole grid
//          when the star is open).  R is swept in 8x8-cell blocks: lanes cull blocks, then the cells of each
//          surviving block, against the disks and ghost half-planes; only surviving cells feed candidates.
//          Near-hull stars have huge but thin conflict regions: almost everything is rejected by the culls.
__device__ __forceinline__ int build_star_warp(WarpStar &ws, int &d, int p, const PointSet &ps, int lane, int &n_exact) {
    const Grid &g = ps.g;
    const double ppx = ps.px[p], ppy = ps.py[p];
    const int pcx = cell_coord(ppx, g.xmin, g.inv_h, g.gx), pcy = cell_coord(ppy, g.ymin, g.inv_h, g.gy);
    Rect old; old.x0 = 1; old.x1 = 0; old.y0 = 1; old.y1 = 0;          // empty
    Rect rc;
    d = 0;
    Bootstrap bs; bs.qpos = bs.qneg = -1;
    for (int half = 1; half <= 2; ++half) {
        rc.x0 = max(0, pcx - half); rc.x1 = min(g.gx - 1, pcx + half);
        rc.y0 = max(0, pcy - half); rc.y1 = min(g.gy - 1, pcy + half);
        const int W = rc.x1 - rc.x0 + 1, H = rc.y1 - rc.y0 + 1, total = W * H;      // <= 25
        int c = -1;
        if (lane < total) {
            int cy = rc.y0 + lane / W, cx = rc.x0 + lane % W;
            if (!(cx >= old.x0 && cx <= old.x1 && cy >= old.y0 && cy <= old.y1)) c = cy * g.gx + cx;
        }
        int r = warp_cells(ws, d, bs, p, ppx, ppy, c, ps, lane, n_exact);
        if (r != STAR_OK) return r;
        int side;
        double m = rc.margin(g, ppx, ppy, side);
        if (side < 0) return d > 0 ? STAR_OK : STAR_NONE;
        double reach2 = warp_reach2(ws, d, lane);
        if (m > 0 && m * m >= reach2) return STAR_OK;
        old = rc;
    }
    // R: cell range covering every circumdisk (lane i owns triangle i)
    int rx0 = g.gx, rx1 = -1, ry0 = g.gy, ry1 = -1;
    if (d == 0) { rx0 = 0; rx1 = g.gx - 1; ry0 = 0; ry1 = g.gy - 1; }
    else if (lane < d) {
        double r2 = ws.r2[lane];
        if (r2 >= 1.0e299) { rx0 = 0; rx1 = g.gx - 1; ry0 = 0; ry1 = g.gy - 1; }
        else {
            double rr = sqrt(r2) * (1.0 + 1.0e-9) + 1.0e-6, cxa = ppx + ws.vx[lane], cya = ppy + ws.vy[lane];
            rx0 = cell_coord(fmax(cxa - rr, g.xmin), g.xmin, g.inv_h, g.gx); rx1 = cell_coord(fmin(cxa + rr, g.xmin + g.gx * g.h), g.xmin, g.inv_h, g.gx);
            ry0 = cell_coord(fmax(cya - rr, g.ymin), g.ymin, g.inv_h, g.gy); ry1 = cell_coord(fmin(cya + rr, g.ymin + g.gy * g.h), g.ymin, g.inv_h, g.gy);
        }
    }
#pragma unroll
    for (int o = 16; o; o >>= 1) {
        rx0 = min(rx0, __shfl_xor_sync(0xFFFFFFFFu, rx0, o)); rx1 = max(rx1, __shfl_xor_sync(0xFFFFFFFFu, rx1, o));
        ry0 = min(ry0, __shfl_xor_sync(0xFFFFFFFFu, ry0, o)); ry1 = max(ry1, __shfl_xor_sync(0xFFFFFFFFu, ry1, o));
    }
    // sweep R in blocks of 8x8 cells
    const int bx0 = rx0 >> 3, bx1 = rx1 >> 3, by0 = ry0 >> 3, by1 = ry1 >> 3;
    const int BW = bx1 - bx0 + 1, nblk = BW * (by1 - by0 + 1);
    for (int b0 = 0; b0 < nblk; b0 += 32) {
        int bi = b0 + lane;
        bool alive = false;
        int bx = 0, by = 0;
        if (bi < nblk) {
            by = by0 + bi / BW; bx = bx0 + bi % BW;
            alive = true;
            if (d > 0) {
                int cxa = bx << 3, cya = by << 3, cxb = min(g.gx, cxa + 8), cyb = min(g.gy, cya + 8);
                double x0 = g.xmin + cxa * g.h - ppx - 1.0e-6, y0 = g.ymin + cya * g.h - ppy - 1.0e-6;
                double x1 = g.xmin + cxb * g.h - ppx + 1.0e-6, y1 = g.ymin + cyb * g.h - ppy + 1.0e-6;
                alive = warp_cell_may_cut(ws, d, x0, y0, x1, y1);
            }
        }
        uint32_t am = __ballot_sync(0xFFFFFFFFu, alive);
        while (am) {
            int l = __ffs(am) - 1; am &= am - 1;
            int cbx = __shfl_sync(0xFFFFFFFFu, bx, l), cby = __shfl_sync(0xFFFFFFFFu, by, l);
#pragma unroll
            for (int hlf = 0; hlf < 2; ++hlf) {
                int cx = (cbx << 3) + (lane & 7), cy = (cby << 3) + (lane >> 3) + 4 * hlf;
                int c = -1;
                if (cx < g.gx && cy < g.gy && cx >= rx0 && cx <= rx1 && cy >= ry0 && cy <= ry1 &&
                    !(cx >= old.x0 && cx <= old.x1 && cy >= old.y0 && cy <= old.y1)) c = cy * g.gx + cx;
                int r = warp_cells(ws, d, bs, p, ppx, ppy, c, ps, lane, n_exact);
                if (r != STAR_OK) return r;
            }
        }
    }
    return d > 0 ? STAR_OK : STAR_NONE;
}
#endif  // __CUDACC__

}  // namespace mvosr
