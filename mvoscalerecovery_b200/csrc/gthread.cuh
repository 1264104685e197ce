// Thread path of the Delaunay stars: ONE STAR PER LANE (scalar code, __host__ __device__: tests/host_sim runs it on the CPU against
// Qhull).
//
// The warp-per-star and two-stars-per-warp paths of gstrip.cuh spread the CANDIDATES of a star over lanes: every step of the walk
// costs a fixed overhead of reductions, ballots and broadcasts (~70 warp instructions) on top of evaluating the circumcentre
// parameter t of every candidate with its error bound (~40 per slot), and most lanes idle in the per-star set-up and in the
// consumers.  Here a lane walks its own star sequentially over the candidate runs of its block in shared memory:
//   * nearest point q0 (certified as in stars_pair), then gift wrapping from q0 back to q0;
//   * a step makes two branch-free passes over the candidates: (1) propose the candidate on the left of p->cur with the smallest
//     circumcentre parameter t in plain float32; (2) certify it: the triangle (p, cur, b) is Delaunay iff its circle is empty, so
//     every other candidate must be certainly OUTSIDE -- the in-circle determinant with p at the origin is m0 |s|^2 + m1 sx + m2 sy
//     with three cofactors of (cur, b): 3 FMA + 3 FMA for the forward error bound (KERR, the filter of stars_fast), no division, no
//     interval per candidate, and no side test (a proposal spoilt by a rounded side fails the certificate);
//   * the step is final when the left cap of the winner's circle lies inside the block (w_cap_inside, as on the other paths).
// Anything float32 cannot certify (side or in-circle sign within the error bound, nearest point not unique, a hole = duplicate
// point among the candidates, more than TDEG neighbours) and anything that leaves the block (hull edges, far neighbours, large
// circles) returns a code and the star goes to the warp-per-star path, exactly as from stars_pair.
#pragma once
#include "gindex.cuh"

namespace mvosr {

constexpr int TDEG = 16;                 // ring capacity of the thread path (16-bit sorted positions, two per register)
enum { TS_OK = 0, TS_NOCAND = 1, TS_OUT = 2, TS_DEFER = 3 };
struct TRing { uint32_t w[TDEG / 2]; };
// (unrolled selects: a dynamically indexed register array would live in local memory)
MVOSR_HD void tring_set(TRing &r, int k, int v) {
#pragma unroll
    for (int j = 0; j < TDEG / 2; ++j)
        if (j == (k >> 1)) r.w[j] = (k & 1) ? ((r.w[j] & 0xFFFFu) | ((uint32_t)v << 16)) : ((r.w[j] & 0xFFFF0000u) | (uint32_t)v);
}
MVOSR_HD int tring_get(const TRing &r, int k) {
    uint32_t w = 0;
#pragma unroll
    for (int j = 0; j < TDEG / 2; ++j) if (j == (k >> 1)) w = r.w[j];
    return (int)((k & 1) ? (w >> 16) : (w & 0xFFFFu));
}

struct TCost { unsigned evals, steps, scans; };      // cost model of the host simulation (MVOSR_THREAD_COST)
#ifdef MVOSR_THREAD_COST
#define TCOST(x) (x)
#else
#define TCOST(x) ((void)0)
#endif

// The star of the point at sorted position p (not a hole).  TS_OK: ring[0..deg) = its neighbours, counter-clockwise from the nearest
// one, as sorted positions; the star is closed.
MVOSR_HD int thread_star(const SortedSet &ps, int p, TRing &ring, int &deg, TCost *cost = nullptr) {
    const float ppx = ps.x[p], ppy = ps.y[p];
    const Block bk = block_of(ps, p, ppx, ppy);
    int lo[BLOCK_ROWS], hi[BLOCK_ROWS];
#pragma unroll
    for (int r = 0; r < BLOCK_ROWS; ++r) {
        const int row = bk.r0 + r;
        lo[r] = 0; hi[r] = 0;
        if (row <= bk.r1) { lo[r] = row_lower(ps, row, bk.xlo); hi[r] = row_upper(ps, row, bk.xhi); }
    }
    const float slack = strip_slack(ps);
    const float BX0 = (bk.open & 1) ? -MVOSR_INFF : bk.xlo - ppx + slack, BX1 = (bk.open & 2) ? MVOSR_INFF : bk.xhi - ppx - slack;
    const float BY0 = (bk.open & 4) ? -MVOSR_INFF : row_ylo(ps, bk.r0) - ppy + slack, BY1 = (bk.open & 8) ? MVOSR_INFF : row_yhi(ps, bk.r1) - ppy - slack;
    // ---- q0: the nearest point, certified by the distance to the block's boundary and unique up to rounding.  (Every candidate loop
    // below is branch-free: the lanes of a warp run them in lock step, each over its own runs.)
    float l1 = MVOSR_INFF, l2 = MVOSR_INFF; int q0 = -1; bool hole = false;
#pragma unroll
    for (int r = 0; r < BLOCK_ROWS; ++r)
        for (int pos = lo[r]; pos < hi[r]; ++pos) {
            hole |= ps.orig[pos] == INF16;
            const float sx = ps.x[pos] - ppx, sy = ps.y[pos] - ppy;
            const float l = pos != p ? fmaf(sx, sx, sy * sy) : MVOSR_INFF;
            const bool better = l < l1;
            l2 = better ? l1 : fminf(l2, l); l1 = better ? l : l1; q0 = better ? pos : q0;
        }
    TCOST(cost->scans++);
    if (hole || q0 < 0) return TS_DEFER;
    const float mg = fminf(fminf(-BX0, BX1), fminf(-BY0, BY1));
    if (!(l1 * 1.000001f < mg * mg)) return TS_OUT;
    if (!(l1 * 1.000002f < l2)) return TS_DEFER;
    // ---- the walk
    float cx = ps.x[q0] - ppx, cy = ps.y[q0] - ppy; int cpos = q0;
    deg = 1;
#pragma unroll
    for (int j = 0; j < TDEG / 2; ++j) ring.w[j] = 0;
    tring_set(ring, 0, q0);
    for (;;) {
        // pass 1: the candidate on the left of p->cur with the smallest circumcentre parameter t = (|s|^2 - s.cur) / cross(cur, s),
        // plain float32 -- a proposal, nothing is decided here (p itself has cross = 0)
        float tb = MVOSR_INFF; int bpos = -1;
#pragma unroll
        for (int r = 0; r < BLOCK_ROWS; ++r)
            for (int pos = lo[r]; pos < hi[r]; ++pos) {
                TCOST(cost->evals++);
                const float sx = ps.x[pos] - ppx, sy = ps.y[pos] - ppy;
                const float cr = cx * sy - cy * sx;
                const float num = fmaf(sx, sx, sy * sy) - fmaf(sx, cx, sy * cy);
                const float t = num * rcp_approx(cr);
                const bool better = cr > 0.f && pos != cpos && t < tb;
                tb = better ? t : tb; bpos = better ? pos : bpos;
            }
        TCOST(cost->steps++);
        if (bpos < 0) return TS_NOCAND;                                            // nothing on the left inside the block: hull edge or far neighbour
        // the proposal's circle (p, cur, b): cofactors of the in-circle determinant with p at the origin and their error bounds
        const float bx = ps.x[bpos] - ppx, by = ps.y[bpos] - ppy;
        const float cl = fmaf(cx, cx, cy * cy), bl = fmaf(bx, bx, by * by);
        const float t0 = cx * by, t1 = cy * bx, t2 = cy * bl, t3 = cl * by, t4 = cl * bx, t5 = cx * bl;
        const float m0 = t0 - t1, m1 = t2 - t3, m2 = t4 - t5;
        const float e0 = KERR * (fabsf(t0) + fabsf(t1)), e1 = KERR * (fabsf(t2) + fabsf(t3)), e2 = KERR * (fabsf(t4) + fabsf(t5));
        if (!(m0 > 64.f * e0)) return TS_OUT;                                      // too flat to bound (or b not certainly on the left)
        // pass 2: the triangle (p, cur, b) is Delaunay iff its circle is empty -- every other candidate must be CERTAINLY outside
        // (det > err, the filter of stars_fast).  That also settles every side the proposal may have got wrong.
        bool bad = false;
#pragma unroll
        for (int r = 0; r < BLOCK_ROWS; ++r)
            for (int pos = lo[r]; pos < hi[r]; ++pos) {
                const float sx = ps.x[pos] - ppx, sy = ps.y[pos] - ppy, sl = fmaf(sx, sx, sy * sy);
                const float det = fmaf(m0, sl, fmaf(m1, sx, m2 * sy));
                const float err = fmaf(e0, sl, fmaf(e1, fabsf(sx), e2 * fabsf(sy))) + 1.0e-30f;
                bad |= !(det > err) && pos != p && pos != cpos && pos != bpos;
            }
        if (bad) return TS_DEFER;
        // ---- centre (-m1, -m2) / (2 m0) with a bound on its error (g_regions); the left cap of the circle must lie inside the block
        const float inv = 0.5f / m0, rho = 2.f * e0 * inv;
        const float vx = -m1 * inv, vy = -m2 * inv;
        const float dv = ((e1 + e2) + (fabsf(m1) + fabsf(m2)) * rho) * inv * 1.5f;
        const float rs = sqrtf(fmaf(vx, vx, vy * vy)) * 1.0001f + 2.f * dv + 1.0e-3f;
        const bool disk_in = vx - rs >= BX0 && vx + rs <= BX1 && vy - rs >= BY0 && vy + rs <= BY1;
        if (!disk_in && !w_cap_inside<false>(cx, cy, 1.f, vx, vy, rs, BX0, BX1, BY0, BY1, WBox())) return TS_OUT;
        if (bpos == q0) break;                                                     // closed
        if (deg >= TDEG) return TS_DEFER;
        tring_set(ring, deg, bpos); ++deg;
        cx = bx; cy = by; cpos = bpos;
    }
    return TS_OK;
}

}  // namespace mvosr
