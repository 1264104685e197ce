// Half-warp ("group") star builder: the production path of the per-frame Delaunay stage.
//
// One group of 16 lanes builds the star of one point; two groups share a warp and run the same loop.
// The star lives in REGISTERS: lane i of the group holds neighbour slot i (id, coordinates relative to
// p, lift), a copy of slot i+1, and the cached circumdisk of star triangle i = (p, slot i, slot i+1).
//   * candidates are spread over the lanes (one each), sorted nearest-first with a 16-lane bitonic
//     network, pre-filtered in parallel against the cached circumdisks (conservative), and only the
//     flagged ones are inserted, one at a time;
//   * an insertion evaluates the EXACT conflict (in-circle with symbolic tie-break / ghost orientation,
//     predicates.cuh) of the candidate against all star triangles at once, lane i testing triangle i;
//     the conflicting arc is removed and the candidate spliced in with shuffles;
//   * levels: the 3x3 cell block, then ring 2 (5x5), each followed by the finality test "every
//     unexamined point is farther than twice the largest circumradius"; stars that are still not final
//     (hull / near-hull) run a directed search: the cell range R covering all current circumdisks (the
//     whole grid while the star is open) is swept in 8x8-cell blocks, blocks and then cells being culled
//     against the disks and ghost half-planes.  The union of the disks only shrinks as the star is
//     clipped, so one sweep of R suffices.
// Same rules, same predicates and therefore the same (unique) triangulation as star.cuh's sequential
// builder, which remains the host-simulated reference and the device fallback for stars of degree > 16.
#pragma once
#include "star.cuh"

namespace mvosr {

constexpr int GL = 16;                   // lanes per group
constexpr int NGROUP = NT / GL;
constexpr double GINF = 1.0e300;

struct GCtx {                            // group-uniform + per-lane registers
    unsigned gmask;                      // lanes of this group inside the warp
    int gl;                              // lane inside the group
    int p, d;
    double ppx, ppy, reach2;
    int sid; double sqx, sqy, sql;       // slot gl
    int nid; double nqx, nqy, nql;       // slot gl+1 (cyclic)
    double vx, vy, r2;                   // triangle gl: circumcentre rel. p and inflated radius^2 (r2 >= 0);
                                         // ghost (p,q,inf): (vx,vy)=q, r2=-1; ghost (p,inf,q): r2=-2; unused lane: r2=-3;
                                         // r2 = GINF: finite but too flat to bound -> always "may conflict"
    double tr4;                          // (2 * circumradius)^2 upper bound, GINF for ghost / unknown
    int n_exact;
};

template <class T> __device__ __forceinline__ T gshfl(const GCtx &c, T v, int src) { return __shfl_sync(c.gmask, v, src, GL); }
__device__ __forceinline__ unsigned gballot(const GCtx &c, bool pred) {
    return (__ballot_sync(c.gmask, pred) >> (c.gmask & 0x10000u ? 16 : 0)) & 0xFFFFu;
}
__device__ __forceinline__ double gmax(const GCtx &c, double v) {
#pragma unroll
    for (int o = 8; o; o >>= 1) v = fmax(v, __shfl_xor_sync(c.gmask, v, o, GL));
    return v;
}
__device__ __forceinline__ int gmin_i(const GCtx &c, int v) {
#pragma unroll
    for (int o = 8; o; o >>= 1) v = min(v, __shfl_xor_sync(c.gmask, v, o, GL));
    return v;
}
__device__ __forceinline__ int gmax_i(const GCtx &c, int v) {
#pragma unroll
    for (int o = 8; o; o >>= 1) v = max(v, __shfl_xor_sync(c.gmask, v, o, GL));
    return v;
}

// Refresh the copy of the next slot and the triangle cache after any change of the star.
__device__ __forceinline__ void g_refresh(GCtx &c) {
    int nx = c.gl + 1 < c.d ? c.gl + 1 : 0;
    c.nid = gshfl(c, c.sid, nx); c.nqx = gshfl(c, c.sqx, nx); c.nqy = gshfl(c, c.sqy, nx); c.nql = gshfl(c, c.sql, nx);
    double vx = 0, vy = 0, r2 = -3.0, tr4 = 0;
    if (c.gl < c.d) {
        if (c.nid == INF16) { vx = c.sqx; vy = c.sqy; r2 = -1.0; tr4 = GINF; }
        else if (c.sid == INF16) { vx = c.nqx; vy = c.nqy; r2 = -2.0; tr4 = GINF; }
        else {
            double l = c.sqx * c.nqy, r = c.sqy * c.nqx, w = l - r, aw = fabs(l) + fabs(r);
            double wl = w - 4.0e-16 * aw;
            r2 = GINF; tr4 = GINF;
            if (wl > 0) {
                double ex = c.sqx - c.nqx, ey = c.sqy - c.nqy;
                tr4 = c.sql * c.nql * (ex * ex + ey * ey) / (wl * wl) * (1.0 + 1.0e-9);
                if (w > 1.0e-6 * aw) {
                    double inv = 0.5 / w;
                    vx = (c.sql * c.nqy - c.nql * c.sqy) * inv; vy = (c.nql * c.sqx - c.sql * c.nqx) * inv;
                    r2 = (vx * vx + vy * vy) * (1.0 + 1.0e-5);
                }
            }
        }
    }
    c.vx = vx; c.vy = vy; c.r2 = r2; c.tr4 = tr4;
    c.reach2 = c.d >= 3 ? gmax(c, tr4) : GINF;
}

// Conservative, lane-parallel: may the lane's candidate (relative sx,sy) cut the cell of p?
__device__ __forceinline__ bool g_prefilter(const GCtx &c, double sx, double sy) {
    bool hit = false;
    for (int i = 0; i < c.d; ++i) {
        double vx = gshfl(c, c.vx, i), vy = gshfl(c, c.vy, i), r2 = gshfl(c, c.r2, i);
        if (r2 >= 0) {
            double dx = sx - vx, dy = sy - vy;
            hit |= (dx * dx + dy * dy < r2);
        } else {
            double l = vx * sy, r = vy * sx, tol = 3.4e-16 * (fabs(l) + fabs(r));
            hit |= (r2 == -1.0) ? (l - r >= -tol) : (l - r <= tol);
        }
    }
    return hit;
}

// Conservative, lane-parallel: can the rectangle [x0,x1]x[y0,y1] (relative to p) contain a cutting point?
__device__ __forceinline__ bool g_rect_may_cut(const GCtx &c, double x0, double y0, double x1, double y1) {
    bool hit = false;
    for (int i = 0; i < c.d; ++i) {
        double vx = gshfl(c, c.vx, i), vy = gshfl(c, c.vy, i), r2 = gshfl(c, c.r2, i);
        if (r2 >= 0) {
            double dx = vx < x0 ? x0 - vx : (vx > x1 ? vx - x1 : 0.0), dy = vy < y0 ? y0 - vy : (vy > y1 ? vy - y1 : 0.0);
            hit |= (dx * dx + dy * dy < r2);
        } else {
            double sgn = r2 == -1.0 ? 1.0 : -1.0;
            double c0 = sgn * (vx * y0 - vy * x0), c1 = sgn * (vx * y0 - vy * x1), c2 = sgn * (vx * y1 - vy * x0), c3 = sgn * (vx * y1 - vy * x1);
            double tol = 1.0e-9 * (fabs(vx) + fabs(vy)) * (fabs(x0) + fabs(x1) + fabs(y0) + fabs(y1) + 1.0);
            hit |= (fmax(fmax(c0, c1), fmax(c2, c3)) >= -tol);
        }
    }
    return hit;
}

// Exact insertion of candidate s (group-uniform).  Returns 1 inserted, 0 no conflict, <0 -STAR_* error.
__device__ __forceinline__ int g_insert(GCtx &c, int s, double sx, double sy, double sl) {
    bool cf_lane = false;
    if (c.gl < c.d) {
        if (c.nid == INF16) {                // ghost (p, slot, inf): outside lies LEFT of p->slot
            int o = cross_sign(c.sqx, c.sqy, sx, sy, c.n_exact);
            cf_lane = o > 0 || (o == 0 && strictly_between(c.sqx, c.sqy, sx, sy));
        } else if (c.sid == INF16) {         // ghost (p, inf, next): outside lies RIGHT of p->next
            int o = cross_sign(c.nqx, c.nqy, sx, sy, c.n_exact);
            cf_lane = o < 0 || (o == 0 && strictly_between(c.nqx, c.nqy, sx, sy));
        } else {
            cf_lane = incircle_sos(c.sqx, c.sqy, c.sql, c.nqx, c.nqy, c.nql, sx, sy, sl, c.p, c.sid, c.nid, s, c.n_exact);
        }
    }
    const unsigned cf = gballot(c, cf_lane);
    if (!cf) return 0;
    const int d = c.d;
    const unsigned full = (1u << d) - 1u;
    unsigned prevm = ((cf << 1) | (cf >> (d - 1))) & full;
    unsigned starts = cf & ~prevm;
    if (__popc(starts) != 1) return -STAR_INCONSISTENT;
    int i0 = __ffs(starts) - 1, len = __popc(cf);
    unsigned rot = i0 ? (((cf >> i0) | (cf << (d - i0))) & full) : cf;
    if (rot != ((1u << len) - 1u) || len >= d) return -STAR_INCONSISTENT;
    const int nd = d - len + 2;
    if (nd > GL) return -STAR_OVERFLOW;
    // new[0] = s ; new[k] = old[(i0+len+k-1) % d], k = 1..nd-1
    int src = (i0 + len + c.gl - 1) % d;
    if (c.gl == 0 || c.gl >= nd) src = 0;
    int id2 = gshfl(c, c.sid, src); double x2 = gshfl(c, c.sqx, src), y2 = gshfl(c, c.sqy, src), l2 = gshfl(c, c.sql, src);
    if (c.gl == 0) { id2 = s; x2 = sx; y2 = sy; l2 = sl; }
    c.sid = id2; c.sqx = x2; c.sqy = y2; c.sql = l2;
    c.d = nd;
    g_refresh(c);
    return 1;
}

struct GTmpStar { int v[4]; __device__ __forceinline__ void set(int i, int x) { v[i] = x; } };

// A batch of candidates, one per lane (id < 0: none).  Sorted nearest-first when `sort` is set.
__device__ __forceinline__ int g_batch(GCtx &c, Bootstrap &bs, int id, bool sort, const double *px, const double *py) {
    double sx = 0, sy = 0, sl = GINF;
    if (id >= 0 && id != c.p) { sx = px[id] - c.ppx; sy = py[id] - c.ppy; sl = sx * sx + sy * sy; if (sl > c.reach2) { id = -1; sl = GINF; } }
    else id = -1;
    if (!gballot(c, id >= 0)) return 0;
    if (sort) {
        // bitonic sort of (sl, id) over the 16 lanes, ascending
#pragma unroll
        for (int k = 2; k <= GL; k <<= 1) {
#pragma unroll
            for (int j = k >> 1; j > 0; j >>= 1) {
                double osl = __shfl_xor_sync(c.gmask, sl, j, GL); int oid = __shfl_xor_sync(c.gmask, id, j, GL);
                bool up = ((c.gl & k) == 0), lower = ((c.gl & j) == 0);
                bool take = (lower == up) ? (osl < sl) : (osl > sl);
                if (take) { sl = osl; id = oid; }
            }
        }
        if (id >= 0) { sx = px[id] - c.ppx; sy = py[id] - c.ppy; }
    }
    unsigned valid = gballot(c, id >= 0);
    if (c.d == 0) {
        // bootstrap: sequential over the candidates until the star owns a real triangle (group-uniform)
        while (valid && c.d == 0) {
            int j = __ffs(valid) - 1; valid &= valid - 1;
            int s = gshfl(c, id, j);
            int dd = 0; GTmpStar tmp;
            if (star_bootstrap(tmp, dd, bs, s, c.ppx, c.ppy, px, py, c.n_exact)) {
                int q = c.gl < dd ? tmp.v[c.gl] : (int)INF16;
                c.sid = q; c.sqx = 0; c.sqy = 0; c.sql = 0;
                if (c.gl < dd && q != INF16) { c.sqx = px[q] - c.ppx; c.sqy = py[q] - c.ppy; c.sql = c.sqx * c.sqx + c.sqy * c.sqy; }
                c.d = dd;
                g_refresh(c);
            }
        }
        if (c.d == 0) return 0;
    }
    bool mine = (valid >> c.gl) & 1u;
    bool pf = g_prefilter(c, sx, sy);                 // contains group shuffles: every lane must execute it
    unsigned F = gballot(c, mine && sl <= c.reach2 && pf);
    int cnt = 0;
    while (F) {
        int j = __ffs(F) - 1; F &= F - 1;
        int s = gshfl(c, id, j); double csx = gshfl(c, sx, j), csy = gshfl(c, sy, j), csl = gshfl(c, sl, j);
        int r = g_insert(c, s, csx, csy, csl);
        if (r < 0) return r;
        if (r == 1 && F && (++cnt & 1) == 0) {
            bool still = (F >> c.gl) & 1u;
            bool pf2 = g_prefilter(c, sx, sy);
            F &= gballot(c, still && sl <= c.reach2 && pf2);
        }
    }
    return 0;
}

// Finality test after the block of half-width `half` around p's cell has been examined.
// Returns 1 final, 0 not final, 2 the block covers the grid (final by exhaustion).
__device__ __forceinline__ int g_final(const GCtx &c, const Grid &g, int pcx, int pcy, int half) {
    Rect rc;
    rc.x0 = max(0, pcx - half); rc.x1 = min(g.gx - 1, pcx + half);
    rc.y0 = max(0, pcy - half); rc.y1 = min(g.gy - 1, pcy + half);
    int side;
    double m = rc.margin(g, c.ppx, c.ppy, side);
    if (side < 0) return 2;
    return (m > 0 && m * m >= c.reach2) ? 1 : 0;
}

// Build the star of p.  Returns STAR_OK (star in registers, c.d slots), STAR_NONE, or an error/overflow code.
// Written as one loop around a SINGLE g_batch call site (and a single culling site): the body is large, and
// every extra inlined copy costs instruction-cache misses on all 16 warps of the CTA.
//   stage 1: the 3x3 cell block   = three row runs of cell_pts (lanes 0..2 describe them)
//   stage 2: ring 2 of the 5x5    = two 5-cell row runs (lanes 0,1) + left/right cells of the middle rows (lanes 2..7)
//   stage 3: directed search over R = cell range of all circumdisks (whole grid while open / unknown), swept in
//            8x8-cell blocks; blocks, then cells, are culled against the disks and ghost half-planes
__device__ __forceinline__ int g_build(GCtx &c, int p, const PointSet &ps) {
    const Grid &g = ps.g;
    c.p = p; c.d = 0; c.reach2 = GINF; c.ppx = ps.px[p]; c.ppy = ps.py[p];
    c.sid = INF16; c.sqx = c.sqy = c.sql = 0; c.nid = INF16; c.nqx = c.nqy = c.nql = 0; c.vx = c.vy = 0; c.r2 = -3.0; c.tr4 = 0;
    const int pcx = cell_coord(c.ppx, g.xmin, g.inv_h, g.gx), pcy = cell_coord(c.ppy, g.ymin, g.inv_h, g.gy);
    Bootstrap bs; bs.qpos = bs.qneg = -1;
    int stage = 0;                       // bumped to 1 by the first "level end" below
    int run_beg = 0, pre = 0, total = 0, off = 0;
    // directed-search cursor (group-uniform unless noted)
    int rx0 = 0, rx1 = -1, ry0 = 0, ry1 = -1, bx0 = 0, by0 = 0, BW = 1, nblk = 0, b0 = 0, cbx = 0, cby = 0, q = 4, k = 0, nmax = 0;
    unsigned am = 0;
    int bx = 0, by = 0, n = 0, beg = 0;  // per lane
    const int ox0 = max(0, pcx - 2), ox1 = min(g.gx - 1, pcx + 2), oy0 = max(0, pcy - 2), oy1 = min(g.gy - 1, pcy + 2);
    for (;;) {
        int id = -1;
        bool sort = false;
        if (stage <= 2) {
            if (off >= total) {
                // ---- level end (or start)
                if (stage >= 1) {
                    int f = g_final(c, g, pcx, pcy, stage);
                    if (f) return c.d > 0 ? STAR_OK : STAR_NONE;
                }
                ++stage;
                if (stage <= 2) {
                    int len = 0; run_beg = 0;
                    if (stage == 1) {
                        int cy = pcy - 1 + c.gl;
                        if (c.gl < 3 && cy >= 0 && cy < g.gy) {
                            int x0 = max(0, pcx - 1), x1 = min(g.gx - 1, pcx + 1);
                            run_beg = ps.cell_start[cy * g.gx + x0]; len = ps.cell_start[cy * g.gx + x1 + 1] - run_beg;
                        }
                    } else if (c.gl < 2) {
                        int cy = c.gl == 0 ? pcy - 2 : pcy + 2;
                        if (cy >= 0 && cy < g.gy) { run_beg = ps.cell_start[cy * g.gx + ox0]; len = ps.cell_start[cy * g.gx + ox1 + 1] - run_beg; }
                    } else if (c.gl < 8) {
                        int kk = c.gl - 2, cy = pcy - 1 + (kk >> 1), cx = (kk & 1) ? pcx + 2 : pcx - 2;
                        if (cy >= 0 && cy < g.gy && cx >= 0 && cx < g.gx) { run_beg = ps.cell_start[cy * g.gx + cx]; len = ps.cell_start[cy * g.gx + cx + 1] - run_beg; }
                    }
                    pre = len;                           // inclusive prefix over the 8 run lanes
#pragma unroll
                    for (int o = 1; o < 8; o <<= 1) { int t = __shfl_up_sync(c.gmask, pre, o, GL); if (c.gl >= o) pre += t; }
                    total = gshfl(c, pre, 7); off = 0;
                } else {
                    // ---- enter the directed search: R from the circumdisks as they are now
                    int a0 = g.gx, a1 = -1, c0 = g.gy, c1 = -1;
                    if (c.d == 0) { a0 = 0; a1 = g.gx - 1; c0 = 0; c1 = g.gy - 1; }
                    else if (c.gl < c.d) {
                        if (c.r2 < 0 || c.r2 >= 1.0e299) { a0 = 0; a1 = g.gx - 1; c0 = 0; c1 = g.gy - 1; }
                        else {
                            double rr = sqrt(c.r2) * (1.0 + 1.0e-9) + 1.0e-6, cxa = c.ppx + c.vx, cya = c.ppy + c.vy;
                            a0 = cell_coord(fmax(cxa - rr, g.xmin), g.xmin, g.inv_h, g.gx); a1 = cell_coord(fmin(cxa + rr, g.xmin + g.gx * g.h), g.xmin, g.inv_h, g.gx);
                            c0 = cell_coord(fmax(cya - rr, g.ymin), g.ymin, g.inv_h, g.gy); c1 = cell_coord(fmin(cya + rr, g.ymin + g.gy * g.h), g.ymin, g.inv_h, g.gy);
                        }
                    }
                    rx0 = gmin_i(c, a0); rx1 = gmax_i(c, a1); ry0 = gmin_i(c, c0); ry1 = gmax_i(c, c1);
                    bx0 = rx0 >> 3; by0 = ry0 >> 3; BW = (rx1 >> 3) - bx0 + 1; nblk = BW * ((ry1 >> 3) - by0 + 1);
                    b0 = -GL; am = 0; q = 4; k = 0; nmax = 0;
                }
                continue;
            }
            // ---- next batch of the level: lane t takes element off+gl of the concatenated runs
            {
                int t = off + c.gl, r = 0, base = 0;
#pragma unroll
                for (int kk = 0; kk < 7; ++kk) { int pk = gshfl(c, pre, kk); if (t >= pk) { r = kk + 1; base = pk; } }
                int rb = gshfl(c, run_beg, r);
                if (t < total) { int v = ps.cell_pts[rb + (t - base)]; if (v != INF16) id = v; }
                off += GL; sort = true;
            }
        } else {
            // ---- directed search cursor: advance until a batch of candidates is available
            bool got = false;
            for (;;) {
                if (k < nmax) { id = k < n ? (int)ps.cell_pts[beg + k] : -1; ++k; got = true; break; }
                int kind;                                  // 0: cells of chunk q of block (cbx,cby); 1: next 16 blocks
                if (q < 4) kind = 0;
                else if (am) { int l = __ffs(am) - 1; am &= am - 1; cbx = gshfl(c, bx, l); cby = gshfl(c, by, l); q = 0; kind = 0; }
                else { b0 += GL; if (b0 >= nblk) break; kind = 1; }
                double x0 = 0, y0 = 0, x1 = 0, y1 = 0;
                bool valid = false;
                int cc = 0;
                if (kind == 0) {
                    int cx = (cbx << 3) + (c.gl & 7), cy = (cby << 3) + (c.gl >> 3) + 2 * q;
                    valid = cx < g.gx && cy < g.gy && cx >= rx0 && cx <= rx1 && cy >= ry0 && cy <= ry1 &&
                            !(cx >= ox0 && cx <= ox1 && cy >= oy0 && cy <= oy1);
                    if (valid) {
                        cc = cy * g.gx + cx;
                        x0 = g.xmin + cx * g.h - c.ppx - 1.0e-6; y0 = g.ymin + cy * g.h - c.ppy - 1.0e-6;
                        x1 = x0 + g.h + 2.0e-6; y1 = y0 + g.h + 2.0e-6;
                    }
                } else {
                    int bi = b0 + c.gl;
                    if (bi < nblk) {
                        by = by0 + bi / BW; bx = bx0 + bi % BW; valid = true;
                        int cxa = bx << 3, cya = by << 3, cxb = min(g.gx, cxa + 8), cyb = min(g.gy, cya + 8);
                        x0 = g.xmin + cxa * g.h - c.ppx - 1.0e-6; y0 = g.ymin + cya * g.h - c.ppy - 1.0e-6;
                        x1 = g.xmin + cxb * g.h - c.ppx + 1.0e-6; y1 = g.ymin + cyb * g.h - c.ppy + 1.0e-6;
                    }
                }
                bool cut = true;
                if (c.d > 0) cut = g_rect_may_cut(c, x0, y0, x1, y1);       // the single culling site (group shuffles inside)
                if (kind == 0) {
                    n = 0; beg = 0;
                    if (valid && cut) { n = ps.cell_n[cc]; beg = ps.cell_start[cc]; }
                    nmax = gmax_i(c, n); k = 0; ++q;
                } else {
                    am = gballot(c, valid && cut);
                }
            }
            if (!got) return c.d > 0 ? STAR_OK : STAR_NONE;
        }
        int r = g_batch(c, bs, id, sort, ps.px, ps.py);                      // the single batch site
        if (r < 0) return -r;
    }
}

}  // namespace mvosr
