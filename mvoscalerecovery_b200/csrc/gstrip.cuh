// Per-site Delaunay stars over a STRIP-sorted point set staged in shared memory (device code, sm_100a).
//
// Replaces scipy.spatial.Delaunay(feature2d).simplices (Qhull 8.0.2) at reference
// src/rescale.py:124-125,136-137, per frame.
//
// Spatial index (round 2; gstar.cuh holds the round-1 uniform grid, kept selectable with -DMVOSR_UNIFORM_GRID): the points are
// cut into horizontal STRIPS whose heights follow the local density -- a strip is grown over the bins of a y-histogram until
// (points in it) x (its height) >= k x (its x-extent), i.e. until a cell holding k points is about square -- and sorted by x
// inside each strip.  Every neighbourhood query is then "the points of strip r with x in [a, b]": two binary searches, a
// contiguous run of shared memory.  The candidate block of a point is its own strip and the two above and below, cut to the
// x-window spanned by its WIN_M-th neighbours on either side in its own strip: about the same ~35 candidates whatever the
// density, where the uniform grid (tuned to image-uniform features) sent most stars of a perspective or clustered feature
// distribution to the slow paths (bench.py --workload kitti00-ground: 40 k frames/s against 209 k uniform).
//
// Method (no shared mutable mesh, no inter-thread ordering): the Delaunay star of every point p -- the cyclic
// counter-clockwise list of its neighbours -- is built independently.  Exact predicates with a symbolic tie-break
// (predicates.cuh) make the triangulation unique, so independently built stars agree; triangle (a<b<c) is emitted by the
// star of a, into a block that a scan later orders: the canonical triangle order comes for free.  Hull points carry one INF
// slot standing for the two ghost triangles (p,q,inf),(p,inf,q).
//
// Four levels, each handing what it cannot certify to the next (run_stars):
//   1. stars_pair  two stars per warp in lock step (one per half-warp), lanes = the points of the candidate block held in
//                  registers.  Gift wrapping: from edge (p,cur) the next neighbour is the candidate on the left with the
//                  smallest circumcentre parameter t; float32 with forward error bounds, a step is final when the winner's
//                  interval is disjoint from every other candidate's and the left cap of its circle lies inside the block.
//                  Closed stars only.
//   2. stars_wrap  one warp per star, same algorithm plus streaming: when the cap leaves the block, or nothing lies on the
//                  left (hull edge / far neighbour), the strips the cap (or the half-plane) covers are gathered 32 strips at
//                  a time, nearest first, and their points evaluated in dense batches (w_stream).
//   3. stars_fast  exact half-warp path for whatever float32 could not certify (ties, collinearities, crowded blocks):
//                  Bowyer-Watson restricted to one star held in registers, float32 filter + exact float64/integer
//                  predicates, strip sweep bounded by the union of circumdisks.
//   4. fb_build    exact full-warp path, 32 slots, collinear bootstrap, scan in growing square windows.
//   5. hub_star    stars of more than 32 neighbours, kept in memory instead of on the lanes of a warp (up to 254 neighbours).
//
// Two passes per frame.  The vote pass (Delaunay #1, EMIT = false) feeds every finished star to the depth-order graph vote
// (consume_vote) and stores its ring of neighbours (FrameView::rpool).  The emit pass (Delaunay #2, EMIT = true) runs over
// the survivors of the graph check, a subset of the same points: stars that lost no neighbour are emitted straight from the
// stored ring by the caller (frame_kernel.cuh), the others are rebuilt here SEEDED -- surviving neighbours are still
// neighbours, consecutive survivors with nothing dropped between them still span a triangle, so only the gaps left by dropped
// neighbours are walked (stars_pair / stars_wrap, fv.oldof != nullptr).
#pragma once
#include <stdint.h>
#include <math_constants.h>
#include <stdio.h>
#include "predicates.cuh"
#include "gindex.cuh"
#include "gthread.cuh"

namespace mvosr {

#ifndef MVOSR_NT
#define MVOSR_NT 896                     // 28 warps: 73 registers per thread, no spills in the star paths (1024: 64 registers, ~2 % slower)
#endif
constexpr int NT = MVOSR_NT;             // threads per CTA of the fused frame kernel
constexpr int NWARP = NT / 32;
constexpr int GL = 16;                   // lanes per star on the fast path
enum { STAR_OK = 0, STAR_DEFER = 1, STAR_OVERFLOW = 2, STAR_INCONSISTENT = 3, STAR_NONE = 4 };

struct StarCtl {                         // shared-memory work queues of one Delaunay pass
    int next_pos, next_thread, n_defer, n_defer_hi, n_defer2, n_hub, n_wrap;    // n_defer_hi: stars queued from the END of defer[] (expected to be long: open / far-neighbour stars, taken first)
    unsigned long long cnt[8];           // profiling counters (MVOSR_STAR_COUNTERS): tests, splices, batches, rows, runs, exact, loop iterations, refill iterations
};

// ---------------------------------------------------------------------------------------------
// consumers of a finished star held one slot per lane (W lanes: 16 fast path, 32 fallback)
// ---------------------------------------------------------------------------------------------
struct FrameView {
    const float *Z;                      // depth by feature index (graph check)
    uint8_t *pflag;                      // by feature index: bit0 duplicate, bit1 keep
    uint16_t *tri;                       // [T][3] feature indices, ascending inside a row
    uint16_t *tbase; uint8_t *tcnt;      // by feature index: block of triangles owned (smallest vertex)
    int *T; int *status;                 // triangle counter, frame status (shared)
    uint32_t pass_mask;
    int tri_cap;
    // ring store (vote pass): the finished star of feature o as original feature indices, counter-clockwise, INF16 = the
    // hull gap; rinfo[o] = (first pool entry << 8) | degree, 0 = no ring.  Stars none of whose neighbours is dropped by
    // the graph check are stars of Delaunay #2 as well and are emitted from here instead of being rebuilt.
    uint16_t *rpool; uint32_t *rinfo; int *rcount; int rpool_cap;
    // seeds (emit pass of Delaunay #2): oldof[new feature index] = index in Delaunay #1; the rings of the stars to rebuild have
    // been rewritten as sorted positions of the new grid, RING_DROPPED marking the neighbours the graph check removed
    const uint16_t *oldof;
    // float64 inputs (SRC_F64 of the frame kernel): the votes compare the float32 roundings (rounding is monotone, so an order seen
    // in float32 is the order of the float64 values) and fall back to the float64 values of feature o -- f2d[2*(fbase+srcidx[o])+1],
    // f3d[3*(fbase+srcidx[o])+2] -- when two roundings tie.  NULL: float32 inputs.
    const double *f3d, *f2d; const uint32_t *srcidx; int fbase;
};
constexpr uint16_t RING_DROPPED = 0xFFFE;
// rinfo degree field, bit 7: a PARTIAL ring (vote pass) -- the counter-clockwise part of a star that stars_pair certified before it had to
// give the star up, as sorted positions from the nearest neighbour on; stars_wrap resumes the walk there and clears the word
constexpr uint32_t RING_PARTIAL = 0x80u;

template <typename T>
__device__ __forceinline__ bool edge_consistent(T va, T za, T vb, T zb) {
    // check_triangle (graph.py:124-129): (v_a - v_b) * (d_a - d_b) < 0 in float64 on float32-exact values --
    // the product of two exact non-zero differences cannot underflow, so the sign rule is exact
    return (va < vb && za > zb) || (va > vb && za < zb);
}

// graph vote (graph.py:131-145) of triangle (p,a,b) for vertex p.  The reference orders the vertices by feature index (i0 < i1 < i2)
// and looks up (consistent(0,1), consistent(1,2), consistent(0,2), position of p): with the rank of every vertex among the three, the
// edge whose ranks sum to 1 is (0,1), to 3 is (1,2), to 2 is (0,2) -- no data movement.
template <typename T>
__device__ __forceinline__ bool graph_vote(int op, T vp, T zp, int oa, T va, T za, int ob, T vb, T zb,
                                           uint32_t pass_mask) {
    const int pa = op > oa, pb = op > ob, ab = oa > ob;
    const int rp = pa + pb, ra = (1 - pa) + ab, rb = 3 - rp - ra;
    // weight of an edge in idx = a*4 + b*2 + c by its rank sum s: s = 1 -> 4, s = 2 -> 1, s = 3 -> 2
    const int wpa = (0x2140 >> (4 * (rp + ra))) & 7, wpb = (0x2140 >> (4 * (rp + rb))) & 7, wab = (0x2140 >> (4 * (ra + rb))) & 7;
    const int idx = (edge_consistent(vp, zp, va, za) ? wpa : 0) + (edge_consistent(vp, zp, vb, zb) ? wpb : 0) + (edge_consistent(va, za, vb, zb) ? wab : 0);
    return (pass_mask >> (idx * 3 + rp)) & 1u;
}

// the vote of star triangle (p, sid, nid) (sorted positions) for p
__device__ __forceinline__ bool graph_vote_at(const SortedSet &ps, const FrameView &fv, int p, int sid, int nid) {
    const int op = ps.orig[p], oa = ps.orig[sid], ob = ps.orig[nid];
    const float vp = ps.y[p], va = ps.y[sid], vb = ps.y[nid], zp = fv.Z[op], za = fv.Z[oa], zb = fv.Z[ob];
    if (fv.f3d && (vp == va || vp == vb || va == vb || zp == za || zp == zb || za == zb)) {
        const size_t ip = (size_t)fv.fbase + fv.srcidx[op], ia = (size_t)fv.fbase + fv.srcidx[oa], ib = (size_t)fv.fbase + fv.srcidx[ob];
        return graph_vote<double>(op, fv.f2d[2 * ip + 1], fv.f3d[3 * ip + 2], oa, fv.f2d[2 * ia + 1], fv.f3d[3 * ia + 2],
                                  ob, fv.f2d[2 * ib + 1], fv.f3d[3 * ib + 2], fv.pass_mask);
    }
    return graph_vote<float>(op, vp, zp, oa, va, za, ob, vb, zb, fv.pass_mask);
}

// the same vote by feature indices (V: the frame's v by feature index -- the values the sorted copy holds)
__device__ __forceinline__ bool graph_vote_orig(const FrameView &fv, const float *V, int op, int oa, int ob) {
    const float vp = V[op], va = V[oa], vb = V[ob], zp = fv.Z[op], za = fv.Z[oa], zb = fv.Z[ob];
    if (fv.f3d && (vp == va || vp == vb || va == vb || zp == za || zp == zb || za == zb)) {
        const size_t ip = (size_t)fv.fbase + fv.srcidx[op], ia = (size_t)fv.fbase + fv.srcidx[oa], ib = (size_t)fv.fbase + fv.srcidx[ob];
        return graph_vote<double>(op, fv.f2d[2 * ip + 1], fv.f3d[3 * ip + 2], oa, fv.f2d[2 * ia + 1], fv.f3d[3 * ia + 2],
                                  ob, fv.f2d[2 * ib + 1], fv.f3d[3 * ib + 2], fv.pass_mask);
    }
    return graph_vote<float>(op, vp, zp, oa, va, za, ob, vb, zb, fv.pass_mask);
}

// keep = (#incident triangles with p>0.6)/(#incident) > 0.5 (graph.py:33-35,131-132); 0/0 -> False.
// With the ring store on (vote pass of the full pipeline) a finished star only STORES its ring here: the votes of all stored rings
// are taken afterwards by votes_from_rings, one star per thread -- in the star builders the triangles of a star sit on 6 of the 16
// (or 32) lanes that built it, and the vote was 6.5 % of the kernel's instructions at 10-14 active lanes.  A star whose ring does
// not fit the pool votes here.
template <int W>
__device__ __forceinline__ void consume_vote(unsigned mask, int gl, int d, int p, int sid, int nid,
                                             const SortedSet &ps, const FrameView &fv) {
    const int op = ps.orig[p];
    if (fv.rpool && d > 0) {
        int rb = 0;
        if (gl == 0) rb = atomicAdd(fv.rcount, d);
        rb = __shfl_sync(mask, rb, 0, W);
        if (rb + d <= fv.rpool_cap) {
            if (gl < d) fv.rpool[rb + gl] = sid != INF16 ? ps.orig[sid] : INF16;
            if (gl == 0) fv.rinfo[op] = ((uint32_t)rb << 8) | (uint32_t)d;
            return;
        }
    }
    const bool tri = gl < d && sid != INF16 && nid != INF16;
    bool vote = false;
    if (tri) vote = graph_vote_at(ps, fv, p, sid, nid);
    unsigned bt = __ballot_sync(mask, tri) & mask, bv = __ballot_sync(mask, vote) & mask;
    if (gl == 0 && 2 * __popc(bv) > __popc(bt)) fv.pflag[op] |= 2;
}

// The graph votes of every star whose ring is in the store: thread o walks the ring of feature o (feature indices, counter-clockwise,
// INF16 = the hull gap) and sets its keep flag.  Block-wide; the caller synchronises.
__device__ __forceinline__ void votes_from_rings(int n, const float *V, const FrameView &fv, const uint16_t *rpool) {
    for (int o = threadIdx.x; o < n; o += NT) {
        const uint32_t info = fv.rinfo[o];
        if (!info || (info & RING_PARTIAL)) continue;
        const int d = (int)(info & 0xFFu);
        const uint16_t *r = rpool + (info >> 8);
        const int first = r[0];
        int prev = first, nt = 0, nv = 0;
        for (int k = 0; k < d; ++k) {
            const int nxt = k + 1 < d ? (int)r[k + 1] : first;
            if (prev != INF16 && nxt != INF16) { ++nt; nv += graph_vote_orig(fv, V, o, prev, nxt) ? 1 : 0; }
            prev = nxt;
        }
        if (2 * nv > nt) fv.pflag[o] |= 2;
    }
}

// triangles (op < oa, ob) are written as one block sorted by (min,max) of the other two vertices
template <int W>
__device__ __forceinline__ void consume_emit(unsigned mask, int gl, int d, int p, int sid, int nid,
                                             const SortedSet &ps, const FrameView &fv) {
    const bool tri = gl < d && sid != INF16 && nid != INF16;
    const int op = ps.orig[p];
    unsigned key = 0xFFFFFFFFu;
    if (tri) {
        int oa = ps.orig[sid], ob = ps.orig[nid];
        if (op < oa && op < ob) key = ((unsigned)min(oa, ob) << 16) | (unsigned)max(oa, ob);
    }
    const bool own = key != 0xFFFFFFFFu;
    const int k = __popc(__ballot_sync(mask, own) & mask);
    if (!k) return;
    int base = 0;
    if (gl == 0) base = atomicAdd(fv.T, k);
    base = __shfl_sync(mask, base, 0, W);
    if (base + k > fv.tri_cap) {
        if (gl == 0) atomicOr(fv.status, MVOSR_ST_OVERFLOW);
#ifdef MVOSR_DEBUG_PRINT
        if (gl == 0) printf("emit overflow: base=%d k=%d cap=%d\n", base, k, fv.tri_cap);
#endif
        return;
    }
    int r = 0;                            // rank of this lane's key (keys of owners are distinct)
#pragma unroll
    for (int j = 0; j < W; ++j) { unsigned kj = __shfl_sync(mask, key, j, W); r += kj < key; }
    if (own) { uint16_t *t = fv.tri + 3 * (base + r); t[0] = (uint16_t)op; t[1] = (uint16_t)(key >> 16); t[2] = (uint16_t)(key & 0xFFFFu); }
    if (gl == 0) { fv.tbase[op] = (uint16_t)base; fv.tcnt[op] = (uint8_t)k; }
}

// ---------------------------------------------------------------------------------------------
// exact conflict of candidate s with star triangle (p,a,b) (a or b may be the INF slot)
// bit0: conflict; bit1: an orientation test came out exactly zero (collinear with p and a hull neighbour)
// ---------------------------------------------------------------------------------------------
__device__ __noinline__ int exact_conflict(const float *x, const float *y, const uint16_t *orig, int p, int a, int b, int s) {
    int n_exact = 0;
    const double ppx = x[p], ppy = y[p];
    const double sx = (double)x[s] - ppx, sy = (double)y[s] - ppy;
    int code;
    if (b == INF16) {                    // ghost (p, a, inf): outside lies LEFT of p->a
        double ax = (double)x[a] - ppx, ay = (double)y[a] - ppy;
        int o = cross_sign(ax, ay, sx, sy, n_exact);
        code = (o > 0 || (o == 0 && strictly_between(ax, ay, sx, sy))) | ((o == 0) << 1);
    } else if (a == INF16) {             // ghost (p, inf, b): outside lies RIGHT of p->b
        double bx = (double)x[b] - ppx, by = (double)y[b] - ppy;
        int o = cross_sign(bx, by, sx, sy, n_exact);
        code = (o < 0 || (o == 0 && strictly_between(bx, by, sx, sy))) | ((o == 0) << 1);
    } else {
        double ax = (double)x[a] - ppx, ay = (double)y[a] - ppy, bx = (double)x[b] - ppx, by = (double)y[b] - ppy;
        code = incircle_sos(ax, ay, ax * ax + ay * ay, bx, by, bx * bx + by * by, sx, sy, sx * sx + sy * sy,
                            orig[p], orig[a], orig[b], orig[s], n_exact);
    }
    return code | (n_exact << 2);          // bits 2..: predicate evaluations that needed exact arithmetic
}

// ---------------------------------------------------------------------------------------------
// fast path
// ---------------------------------------------------------------------------------------------
enum { DK_NONE = 0, DK_DISK = 1, DK_ALL = 2, DK_HALF = 3 };

struct GState {
    // group-uniform
    int p, d;
    float ppx, ppy;
    int pcx, pcy;
    // lane: slot gl, next slot, triangle gl = (p, slot gl, slot gl+1)
    int sid, nid;
    float qx, qy;
    float m0, m1, m2, e0, e1, e2;
    // lane: search region of triangle gl (lazy; valid when !dirty)
    int kind; float vx, vy, rs; int rowlo, rowhi;
    bool dirty;
};

// after any change of the star: fetch the next slot, recompute the cofactors of triangle gl
__device__ __forceinline__ void g_coeffs(GState &g, unsigned mask, int gl) {
    const int nx = gl + 1 < g.d ? gl + 1 : 0;
    g.nid = __shfl_sync(mask, g.sid, nx, GL);
    const float bx = __shfl_sync(mask, g.qx, nx, GL), by = __shfl_sync(mask, g.qy, nx, GL);
    const float ax = g.qx, ay = g.qy;
    if (gl >= g.d) { g.m0 = 1.f; g.m1 = g.m2 = 0.f; g.e0 = g.e1 = g.e2 = 0.f; }       // det = |s|^2 > 0: never in conflict
    else if (g.nid == INF16) {           // ghost (p,a,inf): det3 = ay*sx - ax*sy  (< 0 <=> s left of p->a)
        g.m0 = 0.f; g.m1 = ay; g.m2 = -ax; g.e0 = 0.f; g.e1 = KERR * fabsf(ay); g.e2 = KERR * fabsf(ax);
    } else if (g.sid == INF16) {         // ghost (p,inf,b): det3 = bx*sy - by*sx  (< 0 <=> s right of p->b)
        g.m0 = 0.f; g.m1 = -by; g.m2 = bx; g.e0 = 0.f; g.e1 = KERR * fabsf(by); g.e2 = KERR * fabsf(bx);
    } else {
        const float al = fmaf(ax, ax, ay * ay), bl = fmaf(bx, bx, by * by);
        const float t0 = ax * by, t1 = ay * bx, t2 = ay * bl, t3 = al * by, t4 = al * bx, t5 = ax * bl;
        g.m0 = t0 - t1; g.m1 = t2 - t3; g.m2 = t4 - t5;
        g.e0 = KERR * (fabsf(t0) + fabsf(t1)); g.e1 = KERR * (fabsf(t2) + fabsf(t3)); g.e2 = KERR * (fabsf(t4) + fabsf(t5));
    }
}

// lane: conservative search region of triangle gl, and the strips it touches
__device__ __forceinline__ void g_regions(GState &g, int gl, const SortedSet &ps) {
    int kind = DK_NONE; float vx = 0.f, vy = 0.f, rs = 0.f;
    int rowlo = 0x7FFFFFFF, rowhi = -1;
    if (gl < g.d) {
        if (g.nid == INF16) { kind = DK_HALF; vx = g.qx; vy = g.qy; }                  // region: vx*y - vy*x > 0
        else if (g.sid == INF16) { kind = DK_HALF; vx = -g.m2; vy = g.m1; }              // (-bx,-by): m1 = -by, m2 = bx
        else if (g.m0 > 64.f * g.e0) {
            // circumcentre (-m1,-m2)/(2 m0) with a bound on its error; the disk passes through p (the origin)
            const float inv = 0.5f / g.m0, rho = 2.f * g.e0 * inv;
            vx = -g.m1 * inv; vy = -g.m2 * inv;
            const float dv = ((g.e1 + g.e2) + (fabsf(g.m1) + fabsf(g.m2)) * rho) * inv * 1.5f;
            rs = sqrtf(fmaf(vx, vx, vy * vy)) * 1.0001f + 2.f * dv + 1.0e-3f;
            kind = DK_DISK;
            if (!(rs < 1.0e6f)) kind = DK_ALL;
        } else kind = DK_ALL;                                                           // too flat to bound
        if (kind == DK_DISK) {
            const float pad = strip_slack(ps);
            rowlo = row_of(ps, g.ppy + vy - rs - pad); rowhi = row_of(ps, g.ppy + vy + rs + pad);
        } else { rowlo = 0; rowhi = ps.R - 1; }
    }
    g.kind = kind; g.vx = vx; g.vy = vy; g.rs = rs; g.rowlo = rowlo; g.rowhi = rowhi;
    g.dirty = false;
}

// lane: x-interval (relative to p) of strip `row` that triangle gl's region can touch ([lo,hi], empty: lo > hi)
__device__ __forceinline__ void g_row_interval(const GState &g, const SortedSet &ps, int row, float &lo, float &hi) {
    const float INF = CUDART_INF_F;
    lo = INF; hi = -INF;
    const float pad = strip_slack(ps);
    const float Y0 = row_ylo(ps, row) - g.ppy - pad, Y1 = row_yhi(ps, row) - g.ppy + pad;
    if (g.kind == DK_DISK) {
        const float dy = fmaxf(fmaxf(Y0 - g.vy, g.vy - Y1), 0.f), rem = g.rs * g.rs - dy * dy;
        if (rem > 0.f) { const float hw = sqrtf(rem) * 1.0001f + 1.0e-3f; lo = g.vx - hw; hi = g.vx + hw; }
    } else if (g.kind == DK_ALL) { lo = -INF; hi = INF; }
    else if (g.kind == DK_HALF) {
        if (g.vy > 0.f) { const float t = fmaxf(g.vx * Y0, g.vx * Y1) / g.vy; hi = t + 1.0e-5f * fabsf(t) + 1.0e-3f; lo = -INF; }
        else if (g.vy < 0.f) { const float t = fminf(g.vx * Y0 / g.vy, g.vx * Y1 / g.vy); lo = t - 1.0e-5f * fabsf(t) - 1.0e-3f; hi = INF; }
        else if (g.vx > 0.f ? Y1 > 0.f : (g.vx < 0.f ? Y0 < 0.f : true)) { lo = -INF; hi = INF; }
    }
}

// min / max over the GL lanes of a half-warp
__device__ __forceinline__ float gminf(float v, unsigned gmask) {
#pragma unroll
    for (int o = GL / 2; o; o >>= 1) v = fminf(v, __shfl_xor_sync(gmask, v, o, GL));
    return v;
}
__device__ __forceinline__ float gmaxf(float v, unsigned gmask) {
#pragma unroll
    for (int o = GL / 2; o; o >>= 1) v = fmaxf(v, __shfl_xor_sync(gmask, v, o, GL));
    return v;
}

// The stars of list[0..n_list) (sorted positions), one per half-warp; stars it gives up on are appended to defer[].
template <bool EMIT>
__device__ __noinline__ void stars_fast(const SortedSet &ps, const FrameView &fv, StarCtl *sc, const uint16_t *list, int n_list,
                                           uint16_t *defer, int &n_exact) {
    const int lane = threadIdx.x & 31, gl = lane & (GL - 1), gshift = lane & GL;
    const unsigned gmask = 0xFFFFu << gshift;
    GState g;
    g.p = -1; g.d = 0; g.ppx = g.ppy = 0.f; g.pcx = g.pcy = 0; g.sid = g.nid = INF16; g.qx = g.qy = 0.f;
    g.m0 = 1.f; g.m1 = g.m2 = g.e0 = g.e1 = g.e2 = 0.f; g.kind = DK_NONE; g.vx = g.vy = g.rs = 0.f; g.rowlo = 0; g.rowhi = -1; g.dirty = true;
    bool active = true, need_point = true, bail = false;
    int phase = 0, k = 0, t = 0, ri = 0, re = 0, rdir = 1, pa2 = 0, pb2 = 0, currow = 0;      // [ri,re): current run; [pa2,pb2): pending right-hand run
    float exlo = 0.f, exhi = 0.f;                                  // x-window examined first in the point's own strip and its two neighbours
    unsigned F = 0; int bbase = 0, bdir = 1; float cx = 0.f, cy = 0.f;
#ifdef MVOSR_GROUP_COUNTERS
    unsigned c_test = 0, c_splice = 0, c_batch = 0, c_row = 0, c_run = 0, c_exact = 0, c_iter = 0, c_refill = 0, c_b1 = 0, c_b2 = 0, c_b3 = 0, c_star = 0; long long c_t0 = 0;
#define CNT(x) ++x
#else
#define CNT(x)
#endif

    while (active) {
        CNT(c_iter);
        if (F == 0) {
            // ================= refill: next batch / run / row / point =================
            for (;;) {
                CNT(c_refill);
                if (bail) {
                    if (gl == 0) { int slot = atomicAdd(&sc->n_defer2, 1); defer[slot] = (uint16_t)g.p; }
                    bail = false; need_point = true;
                }
                if (need_point) {
#ifdef MVOSR_GROUP_COUNTERS
                    if (gl == 0 && g.p >= 0) {
                        if (c_star > 64) atomicAdd(&sc->cnt[2], 1ull);
                        if (c_star > 256) atomicAdd(&sc->cnt[3], 1ull);
                        if (c_star > 1024) atomicAdd(&sc->cnt[4], 1ull);
                        atomicMax(&sc->cnt[7], (unsigned long long)c_star);
                        atomicMax(&sc->cnt[6], (unsigned long long)(clock64() - c_t0));
                    }
                    c_star = 0; c_t0 = clock64();
#endif
                    int idx = 0;
                    if (gl == 0) idx = atomicAdd(&sc->next_pos, 1);
                    idx = __shfl_sync(gmask, idx, 0, GL);
                    if (idx >= n_list) { active = false; break; }
                    const int pos = list[idx];
                    g.p = pos; g.d = 0; g.ppx = ps.x[pos]; g.ppy = ps.y[pos];
                    { const Block bk = block_of(ps, pos, g.ppx, g.ppy); g.pcx = 0; g.pcy = bk.row; exlo = bk.xlo; exhi = bk.xhi; }
                    g.sid = g.nid = INF16; g.qx = g.qy = 0.f; g.m0 = 1.f; g.m1 = g.m2 = g.e0 = g.e1 = g.e2 = 0.f; g.dirty = true;
                    need_point = false; phase = 0; k = 0; ri = re = 0; pa2 = pb2 = 0;
                }
                if (ri < re) {
                    if (phase == 1 && g.dirty && g.d > 0) {
                        // the star changed since this row's interval was taken: clip the rest of the run to the new one
                        g_regions(g, gl, ps);
                        float la, lb;
                        g_row_interval(g, ps, currow, la, lb);
                        const float na = gminf(la, gmask), nb = gmaxf(lb, gmask);
                        if (na > nb) { ri = re; pa2 = pb2 = 0; continue; }
                        const float xa = na + g.ppx - 1.0e-3f, xb = nb + g.ppx + 1.0e-3f;
                        if (rdir > 0) re = upper_x(ps.x, ri, re, xb);
                        else ri = lower_x(ps.x, ri, re, xa);
                        if (pb2 > pa2) { pa2 = lower_x(ps.x, pa2, pb2, xa); pb2 = upper_x(ps.x, pa2, pb2, xb); }      // the pending run is the right-hand one
                        if (ri >= re) continue;
                    }
                    // next batch of the run, nearest cell first: bit j of F <-> position bbase + j * bdir
                    const int pos = rdir > 0 ? ri + gl : re - 1 - gl;
                    const bool v = pos >= ri && pos < re && pos != g.p && ps.orig[pos] != INF16;
                    cx = v ? ps.x[pos] - g.ppx : 0.f; cy = v ? ps.y[pos] - g.ppy : 0.f;
                    bdir = rdir;
                    if (rdir > 0) { bbase = ri; ri += GL; } else { bbase = re - 1; re -= GL; }
                    CNT(c_batch);
                    F = (__ballot_sync(gmask, v) >> gshift) & 0xFFFFu;
                    if (F && phase == 1 && g.d > 0) {
                        // sweep: screen the batch candidate-parallel (lane = candidate) against every triangle of the star as it is
                        // now -- the regions only shrink while the star grows, so what is certainly outside all of them stays
                        // outside.  Far from p nearly every candidate goes here, for d x 14 instructions per 16 instead of ~40 each.
                        const float sl = fmaf(cx, cx, cy * cy);
                        bool maybe = false;
                        for (int kk = 0; kk < g.d; ++kk) {
                            const float m0 = __shfl_sync(gmask, g.m0, kk, GL), m1 = __shfl_sync(gmask, g.m1, kk, GL), m2 = __shfl_sync(gmask, g.m2, kk, GL);
                            const float e0 = __shfl_sync(gmask, g.e0, kk, GL), e1 = __shfl_sync(gmask, g.e1, kk, GL), e2 = __shfl_sync(gmask, g.e2, kk, GL);
                            const float det = fmaf(m0, sl, fmaf(m1, cx, m2 * cy));
                            const float err = fmaf(e0, sl, fmaf(e1, fabsf(cx), e2 * fabsf(cy))) + 1.0e-30f;
                            maybe |= !(det > err);
                        }
                        F &= (__ballot_sync(gmask, v && maybe) >> gshift) & 0xFFFFu;
                    }
                    if (F) break;
                    continue;
                }
                int row, dir;
                if (pb2 > pa2) { row = currow; ri = pa2; re = pb2; dir = 1; pa2 = pb2 = 0; }
                else if (phase == 0) {
                    if (k >= 3) { phase = 1; t = 0; continue; }
                    row = k == 0 ? g.pcy : (k == 1 ? g.pcy - 1 : g.pcy + 1); ++k;        // own strip first
                    if (row < 0 || row >= ps.R) continue;
                    ri = row_lower(ps, row, exlo); re = row_upper(ps, row, exhi); dir = 1;
                } else {
                    if (g.dirty) g_regions(g, gl, ps);
                    int r0 = __reduce_min_sync(gmask, g.rowlo), r1 = __reduce_max_sync(gmask, g.rowhi);
                    if (g.d == 0) { r0 = 0; r1 = ps.R - 1; }
                    bool found = false;
                    row = 0;
                    const int far = max(g.pcy - r0, r1 - g.pcy);
                    while (r0 <= r1 && t <= 2 * far) {
                        const int off = (t & 1) ? -((t + 1) >> 1) : (t >> 1);
                        ++t;
                        row = g.pcy + off;
                        if (row >= r0 && row <= r1) { found = true; break; }
                    }
                    if (!found) {
                        // ---- the sweep has left the union of the regions: the star is final
                        if (EMIT) consume_emit<GL>(gmask, gl, g.d, g.p, g.sid, g.nid, ps, fv);
                        else consume_vote<GL>(gmask, gl, g.d, g.p, g.sid, g.nid, ps, fv);
                        need_point = true;
                        continue;
                    }
                    CNT(c_row);
                    float la, lb;
                    g_row_interval(g, ps, row, la, lb);
                    float xa = gminf(la, gmask), xb = gmaxf(lb, gmask);
                    if (g.d == 0) { xa = -CUDART_INF_F; xb = CUDART_INF_F; }
                    if (xa > xb) continue;
                    const int ia = row_lower(ps, row, xa + g.ppx - 1.0e-3f), ib = row_upper(ps, row, xb + g.ppx + 1.0e-3f);
                    if (ia >= ib) continue;
                    // split at p (in the three strips examined first: around the examined window): the left part is walked right to
                    // left, the right part left to right -- nearest points first
                    const bool nearrow = row >= g.pcy - 1 && row <= g.pcy + 1;
                    const int sl = lower_x(ps.x, ia, ib, nearrow ? exlo : g.ppx);
                    const int sr = nearrow ? upper_x(ps.x, sl, ib, exhi) : sl;
                    if (sl > ia) { pa2 = sr; pb2 = ib; ri = ia; re = sl; dir = -1; }
                    else { ri = sr; re = ib; dir = 1; }
                    if (ri >= re) continue;
                }
                currow = row; CNT(c_run);
                rdir = dir;
            }
        }
        if (!active) break;
        // ================= test one candidate against all triangles of the star =================
        const int j = __ffs(F) - 1;
        F &= F - 1; CNT(c_test); CNT(c_star);
        const float sx = __shfl_sync(gmask, cx, j, GL), sy = __shfl_sync(gmask, cy, j, GL);
        const int spos = bbase + j * bdir;
        const float sl = fmaf(sx, sx, sy * sy);
        const float det = fmaf(g.m0, sl, fmaf(g.m1, sx, g.m2 * sy));
        const float err = fmaf(g.e0, sl, fmaf(g.e1, fabsf(sx), g.e2 * fabsf(sy))) + 1.0e-30f;
        const bool lane_on = gl < g.d;
        unsigned cf = (__ballot_sync(gmask, lane_on && det < -err) >> gshift) & 0xFFFFu;
        if (__ballot_sync(gmask, lane_on && !(fabsf(det) > err)) & gmask) {
            // ---- the float32 filter could not decide for some triangle: exact evaluation of this candidate
            CNT(c_exact);
            int code = 0;
            if (lane_on) { code = exact_conflict(ps.x, ps.y, ps.orig, g.p, g.sid, g.nid, spos); n_exact += code >> 2; }
            cf = (__ballot_sync(gmask, code & 1) >> gshift) & 0xFFFFu;
            const unsigned zm = __ballot_sync(gmask, code & 2) & gmask;
            if (g.d == 2 && zm) { cf = 0; bail = true; F = 0; CNT(c_b1); }      // collinear bootstrap: fallback path
        }
        const bool first = g.d == 0;
        if (cf != 0 || first) {
            // ---- splice: remove the conflicting arc, insert s after its first slot
            bool ins = cf != 0;
            int src = gl, nd = g.d;
            if (ins) {
                const int d = g.d;
                const unsigned full = (1u << d) - 1u;
                const unsigned prevm = ((cf << 1) | (cf >> (d - 1))) & full;
                const unsigned starts = cf & ~prevm;
                const int i0 = __ffs(starts) - 1, len = __popc(cf);
                const unsigned rot = i0 > 0 ? (((cf >> i0) | (cf << (d - i0))) & full) : cf;
                nd = d - len + 2;
                if (__popc(starts) != 1 || rot != ((1u << len) - 1u) || len >= d || nd > GL) { if (nd > GL) { CNT(c_b3); } else { CNT(c_b2); } ins = false; bail = true; F = 0; nd = d; }
                else if (gl > 0 && gl < nd) { src = i0 + len + gl - 1; if (src >= d) src -= d; }
            }
            if (ins) {
                int sid2 = __shfl_sync(gmask, g.sid, src, GL);
                float qx2 = __shfl_sync(gmask, g.qx, src, GL), qy2 = __shfl_sync(gmask, g.qy, src, GL);
                if (gl == 0) { sid2 = spos; qx2 = sx; qy2 = sy; }
                g.sid = sid2; g.qx = qx2; g.qy = qy2; g.d = nd;
            } else if (first) {
                g.sid = gl == 0 ? spos : (int)INF16; g.qx = gl == 0 ? sx : 0.f; g.qy = gl == 0 ? sy : 0.f; g.d = 2;
            }
            if (ins || first) { g.dirty = true; CNT(c_splice); g_coeffs(g, gmask, gl); }
        }
    }
#ifdef MVOSR_GROUP_COUNTERS
    if (gl == 0) {
        atomicAdd(&sc->cnt[0], (unsigned long long)c_test); atomicAdd(&sc->cnt[1], (unsigned long long)c_splice);
 atomicAdd(&sc->cnt[5], (unsigned long long)c_b1 + ((unsigned long long)c_b2 << 16) + ((unsigned long long)c_b3 << 32));
    }
#endif
#undef CNT
}

// ---------------------------------------------------------------------------------------------
// fallback: one warp, 32 slots, exact predicates, all points
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ double fb_reach2(int lane, int d, int sid, int nid, double ax, double ay, double al, double bx, double by, double bl) {
    // upper bound of (2 * largest circumradius)^2; +inf when the star is open or a triangle is too flat to bound
    double v = 0;
    if (d < 3) v = 1.0e300;
    else if (lane < d) {
        if (sid == INF16 || nid == INF16) v = 1.0e300;
        else {
            double l = ax * by, r = ay * bx, w = (l - r) - 4.0e-16 * (fabs(l) + fabs(r));
            double ex = ax - bx, ey = ay - by;
            v = w > 0 ? al * bl * (ex * ex + ey * ey) / (w * w) * (1.0 + 1.0e-9) : 1.0e300;
        }
    }
#pragma unroll
    for (int o = 16; o; o >>= 1) v = fmax(v, __shfl_xor_sync(0xFFFFFFFFu, v, o));
    return v;
}

// Builds the star of p with the whole warp; on STAR_OK lane i holds slot i (sid) and slot i+1 (nid), d slots.
struct FbResult { int rc, d, sid, nid, n_exact; };
__device__ __noinline__ FbResult fb_build(const SortedSet &ps, int p) {
    const unsigned FULL = 0xFFFFFFFFu;
    const int lane = threadIdx.x & 31;
    const double ppx = ps.x[p], ppy = ps.y[p];
    int n_exact = 0;
    int d = 0, sid = INF16, nid = INF16;
    double qx = 0, qy = 0, ql = 0, bx = 0, by = 0, bl = 0;
    int qpos = -1, qneg = -1;            // collinear bootstrap: nearest point on either side of p on the common line
    double reach2 = 1.0e300;
    int rc = STAR_OK;
    // candidates in growing square windows around p (nearest first keeps the intermediate stars small): round k examines, in
    // every strip the window of half-width w_k = 2^k x (height of p's strip) touches, the points with |x - x_p| <= w_k that
    // the previous round has not seen
    auto scan = [&](int b, int e) {
            for (int base = b; base < e && rc == STAR_OK; base += 32) {
                const int pos = base + lane;
                const bool v = pos < e && pos != p && ps.orig[pos] != INF16;
                double sxl = 0, syl = 0, sll = 0;
                if (v) { sxl = (double)ps.x[pos] - ppx; syl = (double)ps.y[pos] - ppy; sll = sxl * sxl + syl * syl; }
                unsigned F = __ballot_sync(FULL, v && sll <= reach2);
                if (F && d >= 2) {
                    // screen the batch candidate-parallel (lane = candidate) with the float64 filters of the predicates: whatever is
                    // certainly in conflict with no triangle of the star as it is now stays so while the star grows (the regions
                    // only shrink).  An open star sweeps the whole set: nearly every far candidate goes here.
                    bool maybe = false;
                    for (int kk = 0; kk < d; ++kk) {
                        const int ks = __shfl_sync(FULL, sid, kk), kn = __shfl_sync(FULL, nid, kk);
                        const double ax = __shfl_sync(FULL, qx, kk), ay = __shfl_sync(FULL, qy, kk), al = __shfl_sync(FULL, ql, kk);
                        const double cx = __shfl_sync(FULL, bx, kk), cy = __shfl_sync(FULL, by, kk), cl = __shfl_sync(FULL, bl, kk);
                        bool out;
                        if (kn == INF16) { const double l = ax * syl, r = ay * sxl; out = l - r < -3.3306690738754731e-16 * (fabs(l) + fabs(r)); }       // strictly right of p->a
                        else if (ks == INF16) { const double l = cx * syl, r = cy * sxl; out = l - r > 3.3306690738754731e-16 * (fabs(l) + fabs(r)); }   // strictly left of p->b
                        else out = det3_lift_sign_filter(ax, ay, al, cx, cy, cl, sxl, syl, sll) == 1;                                          // strictly outside the circle
                        maybe |= !out;
                    }
                    F &= __ballot_sync(FULL, maybe);
                }
                while (F) {
                    const int j = __ffs(F) - 1; F &= F - 1;
                    const int s = base + j;
                    const double sx = __shfl_sync(FULL, sxl, j), sy = __shfl_sync(FULL, syl, j), sl = __shfl_sync(FULL, sll, j);
                    if (sl > reach2) continue;
                    bool changed = false;
                    if (d == 0) {
                        // ---- bootstrap (warp-uniform): wait for the first point off the line through p and the first candidate
                        if (qpos < 0) { qpos = s; continue; }
                        const double ux = (double)ps.x[qpos] - ppx, uy = (double)ps.y[qpos] - ppy;
                        const int o = cross_sign(ux, uy, sx, sy, n_exact);
                        if (o == 0) {
                            const bool same = (fabs(ux) >= fabs(uy)) ? ((sx > 0) == (ux > 0)) : ((sy > 0) == (uy > 0));
                            if (same) { if (fabs(sx) + fabs(sy) < fabs(ux) + fabs(uy)) qpos = s; }
                            else if (qneg < 0) qneg = s;
                            else {
                                const double nx = (double)ps.x[qneg] - ppx, ny = (double)ps.y[qneg] - ppy;
                                if (fabs(sx) + fabs(sy) < fabs(nx) + fabs(ny)) qneg = s;
                            }
                            continue;
                        }
                        int ids[4], dd = 0;
                        if (o > 0) { ids[dd++] = qpos; ids[dd++] = s; if (qneg >= 0) ids[dd++] = qneg; ids[dd++] = INF16; }
                        else { if (qneg >= 0) ids[dd++] = qneg; ids[dd++] = s; ids[dd++] = qpos; ids[dd++] = INF16; }
                        sid = INF16;
                        for (int i = 0; i < dd; ++i) if (lane == i) sid = ids[i];
                        qx = qy = ql = 0;
                        if (lane < dd && sid != INF16) { qx = (double)ps.x[sid] - ppx; qy = (double)ps.y[sid] - ppy; ql = qx * qx + qy * qy; }
                        d = dd; changed = true;
                    } else {
                        bool c = false;
                        if (lane < d) {
                            if (nid == INF16) {
                                int o = cross_sign(qx, qy, sx, sy, n_exact);
                                c = o > 0 || (o == 0 && strictly_between(qx, qy, sx, sy));
                            } else if (sid == INF16) {
                                int o = cross_sign(bx, by, sx, sy, n_exact);
                                c = o < 0 || (o == 0 && strictly_between(bx, by, sx, sy));
                            } else {
                                c = incircle_sos(qx, qy, ql, bx, by, bl, sx, sy, sl, ps.orig[p], ps.orig[sid], ps.orig[nid], ps.orig[s], n_exact);
                            }
                        }
                        const unsigned cf = __ballot_sync(FULL, c);
                        if (!cf) continue;
                        const unsigned full = d >= 32 ? 0xFFFFFFFFu : ((1u << d) - 1u);
                        const unsigned prevm = ((cf << 1) | (cf >> (d - 1))) & full;
                        const unsigned starts = cf & ~prevm;
                        const int i0 = __ffs(starts) - 1, len = __popc(cf);
                        const unsigned rot = i0 > 0 ? (((cf >> i0) | (cf << (d - i0))) & full) : cf;
                        if (__popc(starts) != 1 || rot != (len >= 32 ? 0xFFFFFFFFu : ((1u << len) - 1u)) || len >= d) { rc = STAR_INCONSISTENT; break; }
                        const int nd = d - len + 2;
                        if (nd > 32) { rc = STAR_OVERFLOW; break; }
                        int src = lane;
                        if (lane > 0 && lane < nd) { src = i0 + len + lane - 1; if (src >= d) src -= d; }
                        int sid2 = __shfl_sync(FULL, sid, src);
                        double qx2 = __shfl_sync(FULL, qx, src), qy2 = __shfl_sync(FULL, qy, src), ql2 = __shfl_sync(FULL, ql, src);
                        if (lane == 0) { sid2 = s; qx2 = sx; qy2 = sy; ql2 = sl; }
                        sid = sid2; qx = qx2; qy = qy2; ql = ql2; d = nd; changed = true;
                    }
                    if (changed) {
                        const int nx = lane + 1 < d ? lane + 1 : 0;
                        nid = __shfl_sync(FULL, sid, nx); bx = __shfl_sync(FULL, qx, nx); by = __shfl_sync(FULL, qy, nx); bl = __shfl_sync(FULL, ql, nx);
                        reach2 = fb_reach2(lane, d, sid, nid, qx, qy, ql, bx, by, bl);
                    }
                }
            }
    };
    const int prow = row_of(ps, (float)ppy);
    const double slack = 2.0e-3 + 1.0e-4 * (double)ps.bh;
    double w = fmax((double)(row_yhi(ps, prow) - row_ylo(ps, prow)), 1.0e-2);
    int pr0 = 1, pr1 = 0; float pxlo = 0.f, pxhi = 0.f;           // what the previous round examined: strips pr0..pr1, x in [pxlo, pxhi]
    for (int round = 0; round < 64 && rc == STAR_OK; ++round, w *= 2.0) {
        const int r0 = row_of(ps, (float)(ppy - w)), r1 = row_of(ps, (float)(ppy + w));
        const float xlo = (float)(ppx - w) - 1.0e-3f, xhi = (float)(ppx + w) + 1.0e-3f;
        for (int row = r0; row <= r1 && rc == STAR_OK; ++row) {
            const int ia = row_lower(ps, row, xlo), ib = row_upper(ps, row, xhi);
            int ea = ia, eb = ia;                                  // [ea, eb): seen in the previous round
            if (row >= pr0 && row <= pr1) { ea = lower_x(ps.x, ia, ib, pxlo); eb = upper_x(ps.x, ea, ib, pxhi); }
            scan(ia, ea);
            if (rc == STAR_OK) scan(eb, ib);
        }
        pr0 = r0; pr1 = r1; pxlo = xlo; pxhi = xhi;
        const bool all_x = xlo <= ps.xmin && xhi >= ps.xmax, all_y = r0 == 0 && r1 == ps.R - 1;
        if (all_x && all_y) break;                                 // every point examined
        // final once every unexamined point is farther than twice the largest circumradius
        if (reach2 < 1.0e299) {
            const double INFD = 1.0e300;
            const double mx = all_x ? INFD : w, mb = r0 == 0 ? INFD : ppy - (double)row_ylo(ps, r0), mt = r1 == ps.R - 1 ? INFD : (double)row_yhi(ps, r1) - ppy;
            const double m = fmin(mx, fmin(mb, mt)) - slack;       // slack: rounding in the strip assignment and of the window
            if (m > 0 && m * m >= reach2) break;
        }
    }
    FbResult res; res.d = d; res.sid = sid; res.nid = nid; res.n_exact = n_exact;
    res.rc = rc != STAR_OK ? rc : (d > 0 ? STAR_OK : STAR_NONE);     // d == 0: p alone, or all other points on one line through p
    return res;
}

// ---------------------------------------------------------------------------------------------
// wrap path: one warp per star, lanes = candidates (the production path)
// ---------------------------------------------------------------------------------------------
// Gift-wrapping around p with the candidates held in registers, two per lane: the points of the 5x5 block of grid
// cells around p.  The nearest point q0 is a Delaunay neighbour; from edge (p,cur) the next neighbour counter-clockwise
// is the candidate w on the left of p->cur with the smallest circumcentre parameter
//     t(s) = (|s|^2 - s.cur) / cross(cur, s)              (coordinates relative to p)
// -- the circle (p,cur,w) then holds no other point on the left.  Every lane evaluates t for its two candidates with a
// forward error bound (float32), one REDUX picks the winner, a second one proves that no other candidate's interval
// overlaps the winner's.  The winner is global once the left cap of its circle lies inside the block; otherwise (and on
// hull edges, where no candidate lies on the left) the grid rows the cap covers are streamed 32 candidates at a time.
// An edge with no left point anywhere is a hull edge: the walk restarts clockwise from q0 (mirrored orientation).
// All decisions are certified by the error bounds; whenever one is not (ties, collinearities, crowded cells) the star
// is handed to the exact path above, which decides with exact predicates.
constexpr int WRAP_BLOCK = 2;                     // half-width of the candidate block in cells

struct WEval { float t, eps; bool cand, susp; };


// candidate s (relative to p) against edge (p,cur), orientation sigma (+1 counter-clockwise walk, -1 clockwise)
__device__ __forceinline__ WEval w_eval(bool valid, float sx, float sy, float sl, float cx, float cy, float sigma) {
    WEval e;
    const float p1 = cx * sy, p2 = cy * sx;
    const float cr = sigma * (p1 - p2);
    const float ecr = 8.f * WU * (fabsf(p1) + fabsf(p2));       // |cr - exact| <= 4u (|p1|+|p2|)
    e.cand = valid && cr > 2.f * ecr;                             // certainly on the walk's left
    e.susp = valid && !(fabsf(cr) > 2.f * ecr);                   // side not certain
    const float q1 = sx * cx, q2 = sy * cy;
    const float num = (sl - q1) - q2;
    const float en = 8.f * WU * (sl + fabsf(q1) + fabsf(q2));     // |num - exact| <= 6u (...)
    const float r = rcp_approx(cr);
    e.t = num * r;
    e.eps = (en + fabsf(e.t) * ecr) * r * 1.01f + 16.f * WU * fabsf(e.t) + 1.0e-30f;
    return e;
}

__device__ __forceinline__ unsigned w_key(float t) {             // order-preserving float -> uint
    unsigned k = __float_as_uint(t);
    return (k & 0x80000000u) ? ~k : (k | 0x80000000u);
}

struct WBest {                          // warp-uniform: current best of a step
    bool have; float t, eps, x, y; int pos;
    float vx, vy, rs;                   // its circle (p,cur,best): centre relative to p, padded radius
};

__device__ __forceinline__ void w_circle(WBest &b, float cx, float cy, float sigma) {
    // centre = cur/2 + (sigma t / 2) * (-cy, cx)
    b.vx = 0.5f * (cx - sigma * b.t * cy); b.vy = 0.5f * (cy + sigma * b.t * cx);
    const float r = sqrt_approx(fmaf(b.vx, b.vx, b.vy * b.vy));
    const float pad = b.eps * (fabsf(cx) + fabsf(cy)) * 0.51f + 1.0e-3f + 1.0e-4f * r;
    b.rs = r + 2.f * pad;
}

// One batch of candidates (one per lane) against the current best of the step.  Returns false if a decision could not
// be certified (the star goes to the exact path).
__device__ __forceinline__ bool w_batch(WBest &b, bool valid, float sx, float sy, int pos, float cx, float cy, float sigma) {
    const unsigned FULL = 0xFFFFFFFFu;
    const float sl = fmaf(sx, sx, sy * sy);
    const WEval e = w_eval(valid, sx, sy, sl, cx, cy, sigma);
    if (__any_sync(FULL, e.susp)) return false;
    const bool flag = e.cand && (!b.have || e.t - e.eps < b.t + b.eps);
    const unsigned fm = __ballot_sync(FULL, flag);
    if (!fm) return true;
    const unsigned k = flag ? w_key(e.t) : 0xFFFFFFFFu;
    const unsigned kmin = __reduce_min_sync(FULL, k);
    const int wl = __ffs(__ballot_sync(FULL, k == kmin)) - 1;
    const float wt = __shfl_sync(FULL, e.t, wl), we = __shfl_sync(FULL, e.eps, wl);
    // the winner must beat every other flagged candidate and the previous best with disjoint intervals
    const bool clash = flag && threadIdx.x % 32 != wl && !(e.t - e.eps > wt + we);
    if (__any_sync(FULL, clash)) return false;
    if (b.have && !(wt + we < b.t - b.eps)) return false;
    b.have = true; b.t = wt; b.eps = we;
    b.x = __shfl_sync(FULL, sx, wl); b.y = __shfl_sync(FULL, sy, wl); b.pos = __shfl_sync(FULL, pos, wl);
    w_circle(b, cx, cy, sigma);
    return true;
}

// Stream the grid cells that the left cap of the best's circle (the whole left half-plane while there is no best)
// covers outside the block [bx0,bx1]x[by0,by1].  false: not certified.
// Dense gather, nearest rows first: lane l takes grid row pcy + (0, -1, +1, -2, +2, ...)[32 g + l] and computes the cell runs
// of that row inside the region as it is now (32 rows at once); a warp scan turns the run lengths into one candidate
// sequence, and the candidates are evaluated 32 at a time whatever row they come from.  A pass over a region defined by ANY
// earlier best certifies the final best (the regions are nested and every comparison is certified with disjoint intervals,
// hence transitive); when the best improves while much of the pass is still ahead -- always, when there was none -- the
// gather is redone with the smaller region.
__device__ __noinline__ bool w_stream(WBest &b, const SortedSet &ps, int p, float ppx, float ppy, int pcy,
                                      int by0, int by1, int blk_rb, int blk_rn, float cx, float cy, float sigma, int cpos) {
    // blk_rb / blk_rn: lane r holds the block's run in strip by0 + r (excluded here: the caller has evaluated it)
    const unsigned FULL = 0xFFFFFFFFu;
    const int lane = threadIdx.x & 31;
    const float INF = CUDART_INF_F;
    const float hx = sigma * cx, hy = sigma * cy;                 // half-plane hx*y - hy*x > 0
    const float pad = strip_slack(ps);
    const int tmax = 2 * max(pcy, ps.R - 1 - pcy);                // largest strip offset index
    for (int g0 = 0; g0 <= tmax; ) {
        const int t = g0 + lane;
        const int row = pcy + ((t & 1) ? -((t + 1) >> 1) : (t >> 1));
        const bool disk = b.have && b.rs < 1.0e6f;                // a larger (or non-finite) circle bounds nothing useful: half-plane only
        bool on = row >= 0 && row < ps.R;
        const int rowc = on ? row : 0;
        const float Y0 = row_ylo(ps, rowc) - ppy - pad, Y1 = row_yhi(ps, rowc) - ppy + pad;
        float lo = -INF, hi = INF;
        if (disk) {
            const float dy = fmaxf(fmaxf(Y0 - b.vy, b.vy - Y1), 0.f), rem = b.rs * b.rs - dy * dy;
            if (rem > 0.f) { const float hw = sqrtf(rem) * 1.0001f + 1.0e-3f; lo = b.vx - hw; hi = b.vx + hw; }
            else on = false;
        }
        if (hy > 0.f) { const float u = fmaxf(hx * Y0, hx * Y1) / hy; hi = fminf(hi, u + 1.0e-5f * fabsf(u) + 1.0e-3f); }
        else if (hy < 0.f) { const float u = fminf(hx * Y0 / hy, hx * Y1 / hy); lo = fmaxf(lo, u - 1.0e-5f * fabsf(u) - 1.0e-3f); }
        else if (!(hx > 0.f ? Y1 > 0.f : (hx < 0.f ? Y0 < 0.f : true))) on = false;
        int beg1 = 0, n1 = 0, beg2 = 0, n2 = 0;
        // the block's run on the block's strips (every lane takes part in the shuffles)
        const int bsrc = min(max(row - by0, 0), 31);
        const int xb0 = __shfl_sync(FULL, blk_rb, bsrc), xbn = __shfl_sync(FULL, blk_rn, bsrc);
        if (on && lo <= hi) {
            const int rs0 = ps.row_start[row], re0 = ps.row_start[row + 1];
            // the strip's points with x in [lo, hi] (relative to p; padded for the rounding of the sum)
            int ia = lo == -INF ? rs0 : row_lower(ps, row, lo + ppx - 1.0e-3f);
            int ib = hi == INF ? re0 : row_upper(ps, row, hi + ppx + 1.0e-3f);
            if (row >= by0 && row <= by1 && xbn > 0) {
                // up to two runs: left and right of the block's own run
                const int e1 = min(ib, xb0), a2 = max(ia, xb0 + xbn);
                if (e1 > ia) { beg1 = ia; n1 = e1 - ia; }
                if (ib > a2) { beg2 = a2; n2 = ib - a2; }
            } else if (ib > ia) { beg1 = ia; n1 = ib - ia; }
        }
        const int c = n1 + n2;
        int S = c;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { const int u = __shfl_up_sync(FULL, S, o); if (lane >= o) S += u; }
        const int total = __shfl_sync(FULL, S, 31);
        bool redo = false;
        for (int base = 0; base < total; base += 32) {
            const int e = base + lane;
            int l0 = 0, l1 = 31;                                  // smallest lane r with S[r] > e
#pragma unroll
            for (int it = 0; it < 5; ++it) {
                const int mid = (l0 + l1) >> 1;
                if (__shfl_sync(FULL, S, mid) > e) l1 = mid; else l0 = mid + 1;
            }
            const int r = l0 & 31;
            const int off = e - (__shfl_sync(FULL, S, r) - __shfl_sync(FULL, c, r));
            const int rb1 = __shfl_sync(FULL, beg1, r), rn1 = __shfl_sync(FULL, n1, r), rb2 = __shfl_sync(FULL, beg2, r);
            const int pos = off < rn1 ? rb1 + off : rb2 + (off - rn1);
            const bool v = e < total && pos != p && pos != cpos && pos != b.pos && ps.orig[pos] != INF16;
            const float sx = v ? ps.x[pos] - ppx : 0.f, sy = v ? ps.y[pos] - ppy : 0.f;
            const int prev = b.pos; const bool had = b.have;
            if (!w_batch(b, v, sx, sy, pos, cx, cy, sigma)) return false;
            if (b.pos != prev && total - base - 32 > (had ? 64 : 0)) { redo = true; break; }
        }
        if (!redo) g0 += 32;
    }
    return true;
}

// The stars of list[0..n_list) (sorted positions), one per warp; stars that need the exact path are appended to
// defer[] (sc->n_defer2).
template <bool EMIT>
__device__ __noinline__ void stars_wrap(const SortedSet &ps, const FrameView &fv, StarCtl *sc, const uint16_t *list, int n_list, uint16_t *defer,
                                        const uint16_t *list_hi = nullptr, int n_hi = 0) {
    // list_hi[0], list_hi[-1], ... list_hi[-(n_hi-1)]: the stars expected to take longest (something lay beyond the pair path's whole block:
    // a hull edge or a far neighbour) are started first so that they do not form the tail of the pass
    const unsigned FULL = 0xFFFFFFFFu;
    const int lane = threadIdx.x & 31;
#ifdef MVOSR_WRAP_COUNTERS
    unsigned w_steps = 0, w_out = 0, w_sstars = 0, w_hull = 0, w_big = 0, w_stars = 0, w_nocand = 0;
    long long w_cyc = 0, w_cyc_open = 0, w_cyc_stream = 0, w_t0 = 0, w_ts = 0; unsigned w_open = 0;     // cycles: all stars / open stars / inside w_stream
#define WCNT(x) ++x
#else
#define WCNT(x)
#endif
    for (;;) {
        int p = 0;
        bool done = false;
        {
            int i = 0;
            if (lane == 0) i = atomicAdd(&sc->next_pos, 1);
            i = __shfl_sync(FULL, i, 0);
            if (i >= n_list + n_hi) done = true; else p = i < n_hi ? list_hi[-i] : list[i - n_hi];
        }
        if (done) break;
        bool ok = true;
#ifdef MVOSR_WRAP_COUNTERS
        w_t0 = clock64();
#endif
        const float ppx = ps.x[p], ppy = ps.y[p];
        // The stars that come here have a circle that left the pair path's block (or a hull edge): a LARGER block -- one more strip
        // on either side, the window widened -- settles most of them without streaming; two candidates per lane hold up to 64.
        Block bk = block_of(ps, p, ppx, ppy);
        const int pcy = bk.row;
        bk.r0 = max(pcy - WRAP_ROWS / 2, 0); bk.r1 = min(pcy + WRAP_ROWS / 2, ps.R - 1);
        bk.xlo = ppx - WRAP_WIDEN * (ppx - bk.xlo); bk.xhi = ppx + WRAP_WIDEN * (bk.xhi - ppx);
        bk.open = (bk.xlo <= ps.xmin ? 1 : 0) | (bk.xhi >= ps.xmax ? 2 : 0) | (bk.r0 == 0 ? 4 : 0) | (bk.r1 == ps.R - 1 ? 8 : 0);
        int by0 = bk.r0, by1 = bk.r1;
        // ---- the block's candidates, two per lane: element e of the concatenated strip runs goes to lane e & 31.  A block that
        // holds more than 64 points (denser strips next to a sparse one) is narrowed: whatever lies beyond it is streamed anyway.
        int rb = 0, rn = 0, posA = -1, posB = -1, M = 0;
        for (int attempt = 0; attempt < 6; ++attempt) {
            rb = 0; rn = 0;
            if (lane <= by1 - by0) { rb = row_lower(ps, by0 + lane, bk.xlo); rn = row_upper(ps, by0 + lane, bk.xhi) - rb; }
            posA = -1; posB = -1; M = 0;
            int eA = lane, eB = lane + 32;
#pragma unroll
            for (int r = 0; r < WRAP_ROWS; ++r) {
                const int bb = __shfl_sync(FULL, rb, r), nn = __shfl_sync(FULL, rn, r);
                if (posA < 0) { if (eA < nn) posA = bb + eA; else eA -= nn; }
                if (posB < 0) { if (eB < nn) posB = bb + eB; else eB -= nn; }
                M += nn;
            }
            if (M <= 64) break;
            // halve the window, drop the outermost strips
            bk.xlo = ppx - 0.5f * (ppx - bk.xlo); bk.xhi = ppx + 0.5f * (bk.xhi - ppx); bk.open &= ~3;
            if (by0 < pcy) { ++by0; bk.open &= ~4; }
            if (by1 > pcy) { --by1; bk.open &= ~8; }
        }
        if (M > 64) ok = false;                                   // crowded cells
        WCNT(w_stars); if (M > 32) { WCNT(w_big); }
        bool streamed = false;
        const bool vA = posA >= 0 && posA != p && ps.orig[posA] != INF16, vB = posB >= 0 && posB != p && ps.orig[posB] != INF16;
        const float ax = vA ? ps.x[posA] - ppx : 0.f, ay = vA ? ps.y[posA] - ppy : 0.f, al = fmaf(ax, ax, ay * ay);
        const float bx = vB ? ps.x[posB] - ppx : 0.f, by = vB ? ps.y[posB] - ppy : 0.f, bl = fmaf(bx, bx, by * by);
        // block bounds relative to p (sides beyond which the set has no point are open), shrunk by the strip-assignment slack
        const float slack = strip_slack(ps);
        const float BX0 = (bk.open & 1) ? -CUDART_INF_F : bk.xlo - ppx + slack, BX1 = (bk.open & 2) ? CUDART_INF_F : bk.xhi - ppx - slack;
        const float BY0 = (bk.open & 4) ? -CUDART_INF_F : row_ylo(ps, by0) - ppy + slack, BY1 = (bk.open & 8) ? CUDART_INF_F : row_yhi(ps, by1) - ppy - slack;
        const WBox GB = set_box(ps, ppx, ppy);
        // ---- one step of the walk: the neighbour that follows cur (relative (cx,cy), position cpos) in direction sigma.
        // Block candidates first, then streaming if the winner's cap leaves the block (or nothing lies on the left).
        // false: a decision could not be certified.  b.have == false on return: hull edge.
        auto step = [&](float cx, float cy, float sigma, int cpos, WBest &b) -> bool {
            const WEval ea = w_eval(vA && posA != cpos, ax, ay, al, cx, cy, sigma);
            WEval eb; eb.t = 0.f; eb.eps = 0.f; eb.cand = false; eb.susp = false;
            if (M > 32) eb = w_eval(vB && posB != cpos, bx, by, bl, cx, cy, sigma);      // warp-uniform
            if (__any_sync(FULL, ea.susp || eb.susp)) return false;
            const unsigned kA = ea.cand ? w_key(ea.t) : 0xFFFFFFFFu, kB = eb.cand ? w_key(eb.t) : 0xFFFFFFFFu;
            const unsigned kmin = __reduce_min_sync(FULL, min(kA, kB));
            b.have = kmin != 0xFFFFFFFFu; b.t = b.eps = b.x = b.y = 0.f; b.pos = -1; b.vx = b.vy = b.rs = 0.f;
            bool inside = false;
            if (b.have) {
                const int wl = __ffs(__ballot_sync(FULL, min(kA, kB) == kmin)) - 1;
                const bool selB = kA != kmin;                     // meaningful on lane wl
                const bool wB = __shfl_sync(FULL, (int)selB, wl) != 0;
                b.t = __shfl_sync(FULL, selB ? eb.t : ea.t, wl); b.eps = __shfl_sync(FULL, selB ? eb.eps : ea.eps, wl);
                b.x = __shfl_sync(FULL, selB ? bx : ax, wl); b.y = __shfl_sync(FULL, selB ? by : ay, wl); b.pos = __shfl_sync(FULL, selB ? posB : posA, wl);
                const float ub = b.t + b.eps;
                const bool clash = (ea.cand && !(lane == wl && !wB) && !(ea.t - ea.eps > ub)) || (eb.cand && !(lane == wl && wB) && !(eb.t - eb.eps > ub));
                if (__any_sync(FULL, clash)) return false;
                w_circle(b, cx, cy, sigma);
                if (!(fabsf(b.t) < 1.0e18f)) return false;
                inside = w_cap_inside<CAP_CLIP>(cx, cy, sigma, b.vx, b.vy, b.rs, BX0, BX1, BY0, BY1, GB);
            }
            WCNT(w_steps);
            if (!inside) {
                WCNT(w_out); if (!b.have) { WCNT(w_nocand); } if (!streamed) { streamed = true; WCNT(w_sstars); }
                WBest bs = b;                                     // (a copy: keeps b itself in registers)
#ifdef MVOSR_WRAP_COUNTERS
                w_ts = clock64();
#endif
                const bool sok = w_stream(bs, ps, p, ppx, ppy, pcy, by0, by1, rb, rn, cx, cy, sigma, cpos);
#ifdef MVOSR_WRAP_COUNTERS
                w_cyc_stream += clock64() - w_ts;
#endif
                if (!sok) return false;
                b = bs;
            }
            return true;
        };
        // ---- the star: lane i keeps the i-th counter-clockwise neighbour (sidC) and the (i+1)-th clockwise one (sidW)
        int sidC = INF16, sidW = INF16, nC = 1, nW = 0;
        bool closed = false;
        // ---- seeded rebuild (Delaunay #2): the surviving neighbours of the star in Delaunay #1 are still neighbours; only
        // the gaps left by dropped neighbours are walked (see stars_pair).  A gap that turns out to hold a hull edge -- the
        // star opened up -- or a winner already in the ring falls back to the plain walk below.
        if (EMIT && fv.oldof && ok) {
            const uint32_t info = fv.rinfo[fv.oldof[ps.orig[p]]];
            const int d0 = (int)(info & 0xFFu);
            const int e = lane < d0 ? (int)fv.rpool[(info >> 8) + lane] : (int)INF16;
            const unsigned am = __ballot_sync(FULL, lane < d0 && e < RING_DROPPED), im = __ballot_sync(FULL, lane < d0 && e == INF16);
            const int m = __popc(am);
            if (m >= 1 && im == 0u && d0 < 31) {
                const int src = lane < m ? (int)__fns(am, 0, lane + 1) : 0;
                const int se = __shfl_sync(FULL, e, src);
                int nxt = src + 1; if (nxt >= d0) nxt = 0;
                unsigned gapm = __ballot_sync(FULL, lane < m && !((am >> nxt) & 1u));
                int sid = lane < m ? se : (int)INF16, n = m;
                int j = max(__ffs(gapm) - 1, 0);
                int cpos = __shfl_sync(FULL, sid, j), tpos = __shfl_sync(FULL, sid, j + 1 < n ? j + 1 : 0);
                float cx = ps.x[cpos] - ppx, cy = ps.y[cpos] - ppy;
                bool fallback = false;
                while (gapm) {
                    WBest b;
                    if (!step(cx, cy, 1.f, cpos, b)) { ok = false; break; }
                    if (!b.have) { fallback = true; break; }
                    if (b.pos == tpos) {
                        gapm &= gapm - 1u;
                        if (gapm) { j = __ffs(gapm) - 1; cpos = __shfl_sync(FULL, sid, j); cx = ps.x[cpos] - ppx; cy = ps.y[cpos] - ppy; }
                    } else {
                        if (__any_sync(FULL, lane < n && sid == b.pos) || n >= 30) { fallback = true; break; }
                        const int up = __shfl_up_sync(FULL, sid, 1);
                        if (lane == j + 1) sid = b.pos; else if (lane > j + 1) sid = up;
                        ++n; gapm = ((gapm >> (j + 1)) << (j + 2)) | (1u << (j + 1)); ++j;
                        cpos = b.pos; cx = b.x; cy = b.y;
                    }
                    tpos = __shfl_sync(FULL, sid, j + 1 < n ? j + 1 : 0);
                }
                if (ok && !fallback) { closed = true; sidC = sid; nC = n; }
            }
        }
        // ---- a partial ring from stars_pair (vote pass): q0 and the counter-clockwise neighbours up to cur are certified already
        int q0 = -1; float q0x = 0.f, q0y = 0.f;
        int resume = 0, rcur = -1;
        if (!EMIT && fv.rpool && ok) {
            const int op = ps.orig[p];
            uint32_t info = 0u;
            if (lane == 0) info = fv.rinfo[op];                       // (one lane reads, clears and -- in the consumer -- rewrites this word)
            info = __shfl_sync(FULL, info, 0);
            if (info & RING_PARTIAL) {
                resume = (int)(info & 0x7Fu);
                const int e = lane < resume ? (int)fv.rpool[(info >> 8) + lane] : (int)INF16;
                if (lane == 0) fv.rinfo[op] = 0u;
                sidC = e; nC = resume;
                q0 = __shfl_sync(FULL, e, 0); rcur = __shfl_sync(FULL, e, resume - 1);
                q0x = ps.x[q0] - ppx; q0y = ps.y[q0] - ppy;
            }
        }
        // ---- q0 = the nearest point: certified by the distance to the block's boundary, unique up to rounding
        if (ok && !closed && !resume) {
            const unsigned kA = vA ? __float_as_uint(al) : 0xFFFFFFFFu, kB = vB ? __float_as_uint(bl) : 0xFFFFFFFFu;
            const unsigned kmin = __reduce_min_sync(FULL, min(kA, kB));
            const float lmin = __uint_as_float(kmin);
            const float mg = fminf(fminf(-BX0, BX1), fminf(-BY0, BY1));
            if (kmin == 0xFFFFFFFFu || !(lmin * 1.000001f < mg * mg)) ok = false;
            else {
                const float thr = lmin * 1.000002f;
                const int cnt = __popc(__ballot_sync(FULL, vA && al <= thr)) + __popc(__ballot_sync(FULL, vB && bl <= thr));
                if (cnt != 1) ok = false;                         // two points at (nearly) the same distance
                const int wl = __ffs(__ballot_sync(FULL, min(kA, kB) == kmin)) - 1;
                const bool selB = kA != kmin;
                q0 = __shfl_sync(FULL, selB ? posB : posA, wl); q0x = __shfl_sync(FULL, selB ? bx : ax, wl); q0y = __shfl_sync(FULL, selB ? by : ay, wl);
            }
        }
        // ---- the plain walk: counter-clockwise from q0 until it closes or meets a hull edge, then clockwise from q0
        if (!closed && !resume) { sidC = lane == 0 ? q0 : (int)INF16; nC = 1; }
        float sigma = 1.f, cx = q0x, cy = q0y; int cpos = q0;
        if (resume) { cpos = rcur; cx = ps.x[rcur] - ppx; cy = ps.y[rcur] - ppy; }
        while (ok && !closed) {
            WBest b;
            if (!step(cx, cy, sigma, cpos, b)) { ok = false; break; }
            if (!b.have) {
                // no point on the walk's left of p->cur anywhere: hull edge
                WCNT(w_hull);
                if (sigma > 0.f) { sigma = -1.f; cx = q0x; cy = q0y; cpos = q0; continue; }
                break;
            }
            if (sigma > 0.f && b.pos == q0) { closed = true; break; }
            if (nC + nW >= 31) { ok = false; break; }
            if (sigma > 0.f) { if (lane == nC) sidC = b.pos; ++nC; } else { if (lane == nW) sidW = b.pos; ++nW; }
            cx = b.x; cy = b.y; cpos = b.pos;
        }
#ifdef MVOSR_WRAP_COUNTERS
        { const long long dt = clock64() - w_t0; w_cyc += dt; if (ok && !closed) { w_cyc_open += dt; ++w_open; } }
#endif
        if (!ok) {
            if (lane == 0) { const int slot = atomicAdd(&sc->n_defer2, 1); defer[slot] = (uint16_t)p; }
            continue;
        }
        // ---- counter-clockwise slot order: clockwise part reversed, q0, counter-clockwise part, INF when open
        int d, sid;
        if (closed) { d = nC; sid = sidC; }
        else {
            d = nW + nC + 1;
            const int fromW = __shfl_sync(FULL, sidW, max(nW - 1 - lane, 0)), fromC = __shfl_sync(FULL, sidC, min(max(lane - nW, 0), 31));
            sid = lane < nW ? fromW : (lane < nW + nC ? fromC : (int)INF16);
        }
        const int nid = __shfl_sync(FULL, sid, lane + 1 < d ? lane + 1 : 0);
        if (EMIT) consume_emit<32>(FULL, lane, d, p, sid, nid, ps, fv);
        else consume_vote<32>(FULL, lane, d, p, sid, nid, ps, fv);
    }
#ifdef MVOSR_WRAP_COUNTERS
    if (lane == 0) {
        atomicAdd(&sc->cnt[0], (unsigned long long)w_steps); atomicAdd(&sc->cnt[1], (unsigned long long)w_out);
        atomicAdd(&sc->cnt[2], (unsigned long long)w_sstars); atomicAdd(&sc->cnt[3], (unsigned long long)w_open);
        atomicAdd(&sc->cnt[4], (unsigned long long)w_cyc); atomicAdd(&sc->cnt[5], (unsigned long long)w_stars);
        atomicAdd(&sc->cnt[6], (unsigned long long)w_cyc_open); atomicAdd(&sc->cnt[7], (unsigned long long)w_cyc_stream);
    }
#endif
#undef WCNT
}

// ---------------------------------------------------------------------------------------------
// pair path: two stars per warp in lock step (the production path for closed stars)
// ---------------------------------------------------------------------------------------------
// Same gift-wrapping as stars_wrap, one star per HALF-warp with up to four block candidates per lane; the two
// half-warps execute one instruction stream (every branch is warp-uniform, the per-star differences are predicated), so
// a star costs half the issue slots.  Only what is cheap in this shape is kept: closed stars whose every step is
// certified inside the 5x5 block.  Hull edges, circles leaving the block, crowded blocks and every uncertified decision
// send the star to stars_wrap (one warp per star, streaming) and from there, if need be, to the exact paths.
__device__ __forceinline__ unsigned gmin_u32(unsigned v, int g) {
    const unsigned a = __reduce_min_sync(0xFFFFFFFFu, g == 0 ? v : 0xFFFFFFFFu);
    const unsigned b = __reduce_min_sync(0xFFFFFFFFu, g == 1 ? v : 0xFFFFFFFFu);
    return g ? b : a;
}

template <bool EMIT>
__device__ __noinline__ void stars_pair(const SortedSet &ps, const FrameView &fv, StarCtl *sc, uint16_t *defer, uint16_t *defer_hi,
                                        const uint16_t *todo, int n_todo) {
    const unsigned FULL = 0xFFFFFFFFu;
    const int lane = threadIdx.x & 31, g = lane >> 4, gl = lane & (GL - 1), gshift = lane & GL;
    const int rot = ps.row_start[ps.R - 1];                       // hull strips first (see stars_wrap's note on the tail)
    const float slack = strip_slack(ps);
#ifdef MVOSR_STAR_COUNTERS
#define PR(k) do { if (gl == 0) atomicAdd(&sc->cnt[k], 1ull); } while (0)
#else
#define PR(k)
#endif
#define GBALLOT(pred) ((__ballot_sync(FULL, (pred)) >> gshift) & 0xFFFFu)
#define GSHFL(v, src) __shfl_sync(FULL, (v), (src), GL)
    for (;;) {
        int i = 0;
        if (lane == 0) i = atomicAdd(&sc->next_pos, 2);
        i = __shfl_sync(FULL, i, 0);
        int p; bool have;
        if (todo) {                                                // only the listed stars (sorted positions, holes excluded)
            if (i >= n_todo) break;
            have = i + g < n_todo;
            p = todo[have ? i + g : i];
        } else {
            if (i >= ps.n) break;
            p = i + g + rot; if (p >= ps.n) p -= ps.n; if (p >= ps.n) p -= ps.n;
            have = i + g < ps.n && ps.orig[p] != INF16;
        }
        bool ok = have, far = false;
        const float ppx = ps.x[p], ppy = ps.y[p];
        const Block bk = block_of(ps, p, ppx, ppy);
        // ---- block candidates: the runs of the block's strips inside its x-window (lane 2r / 2r+1 of the half-warp searches the
        // lower / upper end of strip r0 + r); element e of the concatenated runs goes to lane e & 15, slot e >> 4
        static_assert(2 * BLOCK_ROWS <= GL, "one lane of the half-warp per end of a strip's run");
        int rb[BLOCK_ROWS], cum[BLOCK_ROWS], M = 0;               // cum[r]: candidates in strips r0 .. r0 + r
        {
            const int row = bk.r0 + (gl >> 1);
            int res = 0;
            if (gl < 2 * BLOCK_ROWS && row <= bk.r1) res = (gl & 1) ? row_upper(ps, row, bk.xhi) : row_lower(ps, row, bk.xlo);
#pragma unroll
            for (int r = 0; r < BLOCK_ROWS; ++r) {
                const int lo = GSHFL(res, 2 * r), hi = GSHFL(res, 2 * r + 1);
                M += bk.r0 + r <= bk.r1 ? max(hi - lo, 0) : 0;
                rb[r] = lo - (r ? cum[r - 1] : 0);                  // element e of strip r sits at position e + rb[r]
                cum[r] = M;
            }
        }
        if (M > 64) { PR(0); ok = false; }
        // slots 2 and 3 are only populated (and later evaluated) when one of the two blocks holds more than 32 / 48 points
        const bool m32 = __any_sync(FULL, ok && M > 32), m48 = __any_sync(FULL, ok && M > 48);
        float sx[4], sy[4]; int sp[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            sp[k] = INF16; sx[k] = 0.f; sy[k] = 0.f;
            if (k < 2 || (k == 2 ? m32 : m48)) {                   // warp-uniform
                const int e = gl + GL * k;
                int off = rb[BLOCK_ROWS - 1];
#pragma unroll
                for (int r = BLOCK_ROWS - 2; r >= 0; --r) off = e < cum[r] ? rb[r] : off;
                const int pos = e + off;
                const bool v = e < M && pos != p && ps.orig[pos] != INF16;
                if (v) { sp[k] = pos; sx[k] = ps.x[pos] - ppx; sy[k] = ps.y[pos] - ppy; }
            }
        }
        const float BX0 = (bk.open & 1) ? -CUDART_INF_F : bk.xlo - ppx + slack, BX1 = (bk.open & 2) ? CUDART_INF_F : bk.xhi - ppx - slack;
        const float BY0 = (bk.open & 4) ? -CUDART_INF_F : row_ylo(ps, bk.r0) - ppy + slack, BY1 = (bk.open & 8) ? CUDART_INF_F : row_yhi(ps, bk.r1) - ppy - slack;
        // ---- seed (Delaunay #2 only): the neighbours of this point in Delaunay #1 that survived the graph check are still
        // its neighbours, and two consecutive survivors with nothing dropped between them still span a triangle of the star
        // (their circle was empty before points were removed).  Only the gaps left by dropped neighbours are walked.
        bool seeded = false; int sid = INF16, nC = 1; unsigned gapm = 1u;
        if (EMIT && fv.oldof) {                                    // warp-uniform
            int e = INF16, d0 = 0;
            if (have) {
                const uint32_t info = fv.rinfo[fv.oldof[ps.orig[p]]];
                d0 = (int)(info & 0xFFu);
                if (d0 > GL) d0 = 0;
                if (gl < d0) e = fv.rpool[(info >> 8) + gl];
            }
            const unsigned am = GBALLOT(gl < d0 && e < RING_DROPPED), im = GBALLOT(gl < d0 && e == INF16);
            const int m = __popc(am);
            seeded = m >= 1 && im == 0;
            const int src = gl < m ? (int)__fns(am, 0, gl + 1) : 0;  // old slot of the (gl+1)-th survivor
            const int se = GSHFL(e, src);
            int nxt = src + 1; if (nxt >= d0) nxt = 0;
            const unsigned gm = GBALLOT(seeded && gl < m && !((am >> nxt) & 1u));
            if (seeded) { sid = gl < m ? se : (int)INF16; nC = m; gapm = gm; }
        }
        // ---- q0 (stars without a seed): the nearest point, certified and unique up to rounding
        int q0; float cx, cy;
        {
            unsigned kq = 0xFFFFFFFFu; float mx = 0.f, my = 0.f; int mp = INF16;
            float l[4];
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                l[k] = fmaf(sx[k], sx[k], sy[k] * sy[k]);
                const unsigned key = sp[k] != INF16 ? __float_as_uint(l[k]) : 0xFFFFFFFFu;
                if (key < kq) { kq = key; mx = sx[k]; my = sy[k]; mp = sp[k]; }
            }
            const unsigned kmin = gmin_u32(kq, g);
            const float lmin = __uint_as_float(kmin);
            const float mg = fminf(fminf(-BX0, BX1), fminf(-BY0, BY1));
            if (ok && !seeded && (kmin == 0xFFFFFFFFu || !(lmin * 1.000001f < mg * mg))) { PR(1); ok = false; }
            const float thr = lmin * 1.000002f;
            int cnt = 0;
#pragma unroll
            for (int k = 0; k < 4; ++k) cnt += sp[k] != INF16 && l[k] <= thr;
            const unsigned b1 = GBALLOT(cnt >= 1), b2 = GBALLOT(cnt >= 2);
            if (ok && !seeded && (__popc(b1) != 1 || b2)) { PR(2); ok = false; }
            const int wl = max(__ffs(GBALLOT(kq == kmin)) - 1, 0);
            q0 = GSHFL(mp, wl); cx = GSHFL(mx, wl); cy = GSHFL(my, wl);
        }
        if (!seeded) sid = gl == 0 ? q0 : (int)INF16;
        // ---- the walk.  Lane i of the half-warp keeps the i-th neighbour (counter-clockwise); bit i of gapm: the star is
        // still open between slots i and i+1.  Each step looks for the neighbour that follows slot j (the lowest open
        // slot): either the occupant of slot j+1 -- the gap closes -- or a new neighbour, inserted there.
        int j = max(__ffs(gapm) - 1, 0);
        int cpos = GSHFL(sid, j), tpos = GSHFL(sid, j + 1 < nC ? j + 1 : 0);
        if (seeded) { cx = ps.x[cpos] - ppx; cy = ps.y[cpos] - ppy; }
        bool closed = ok && gapm == 0u, walking = ok && gapm != 0u;
        while (__any_sync(FULL, walking)) {
            float tb = CUDART_INF_F, eb = 0.f, xb = 0.f, yb = 0.f; int pb = INF16;       // tb = +inf: no candidate yet
            float lb_best = CUDART_INF_F, lb_rest = CUDART_INF_F; bool susp = false;
#define PAIR_SLOT(k) { \
                const WEval e = w_eval(sp[k] != INF16 && sp[k] != cpos, sx[k], sy[k], fmaf(sx[k], sx[k], sy[k] * sy[k]), cx, cy, 1.f); \
                susp |= e.susp; \
                const float tk = e.cand ? e.t : CUDART_INF_F, lb = e.cand ? e.t - e.eps : CUDART_INF_F; \
                if (tk < tb) { lb_rest = fminf(lb_rest, lb_best); tb = tk; eb = e.eps; xb = sx[k]; yb = sy[k]; pb = sp[k]; lb_best = lb; } \
                else lb_rest = fminf(lb_rest, lb); }
            PAIR_SLOT(0) PAIR_SLOT(1)
            if (m32) PAIR_SLOT(2)
            if (m48) PAIR_SLOT(3)
#undef PAIR_SLOT
            const unsigned kbest = tb < CUDART_INF_F ? w_key(tb) : 0xFFFFFFFFu;           // order-preserving key of the lane's best
            const unsigned gsusp = GBALLOT(susp && walking);               // (ballots are executed by all 32 lanes)
            if (walking && gsusp) { PR(3); ok = false; walking = false; }
            const unsigned kmin = gmin_u32(walking ? kbest : 0xFFFFFFFFu, g);
            if (walking && kmin == 0xFFFFFFFFu) { PR(4); ok = false; walking = false; far = true; }      // nothing on the left inside the block: hull edge or far neighbour
            const int wl = max(__ffs(GBALLOT(walking && kbest == kmin)) - 1, 0);
            const float wt = GSHFL(tb, wl), we = GSHFL(eb, wl), wx = GSHFL(xb, wl), wy = GSHFL(yb, wl);
            const int wpos = GSHFL(pb, wl);
            const float ub = wt + we;
            const float lbo = gl == wl ? lb_rest : fminf(lb_best, lb_rest);
            const unsigned gclash = GBALLOT(walking && !(lbo > ub));
            if (walking && gclash) { PR(5); ok = false; walking = false; }   // another candidate's interval overlaps the winner's
            if (walking && !(fabsf(wt) < 1.0e18f)) { PR(0); ok = false; walking = false; }
            // the winner's circle must lie inside the block
            const float vx = 0.5f * (cx - wt * cy), vy = 0.5f * (cy + wt * cx);
            const float r = sqrt_approx(fmaf(vx, vx, vy * vy));
            const float rs = r + 2.f * (we * (fabsf(cx) + fabsf(cy)) * 0.51f + 1.0e-3f + 1.0e-4f * r);
            // (quick accept: the whole padded disk inside the block; the cap test proper only when some star needs it)
            const bool disk_in = vx - rs >= BX0 && vx + rs <= BX1 && vy - rs >= BY0 && vy + rs <= BY1;
            if (__any_sync(FULL, walking && !disk_in)) {
                if (walking && !disk_in && !w_cap_inside<CAP_CLIP>(cx, cy, 1.f, vx, vy, rs, BX0, BX1, BY0, BY1, set_box(ps, ppx, ppy))) { PR(6); ok = false; walking = false; }
            }
            if (!EMIT) {
                // vote pass: never seeded -- the plain walk from q0 back to q0
                if (walking) {
                    if (wpos == q0) { closed = true; walking = false; }
                    else if (nC >= GL) { PR(7); ok = false; walking = false; }
                    else { if (gl == nC) sid = wpos; ++nC; cx = wx; cy = wy; cpos = wpos; }
                }
            } else {
                // a winner that already sits in another slot than the expected one contradicts the seed: rebuild elsewhere
                const unsigned dupm = GBALLOT(gl < nC && sid == wpos);
                if (walking && wpos != tpos && (dupm || nC >= GL)) { PR(7); ok = false; walking = false; }
                const bool ins = walking && wpos != tpos, cls = walking && wpos == tpos;
                const int up = GSHFL(sid, max(gl - 1, 0));
                if (ins) {
                    if (gl == j + 1) sid = wpos; else if (gl > j + 1) sid = up;
                    ++nC; gapm = ((gapm >> (j + 1)) << (j + 2)) | (1u << (j + 1)); ++j;
                    cpos = wpos; cx = wx; cy = wy;
                }
                if (cls) {
                    gapm &= gapm - 1u;
                    if (!gapm) { closed = true; walking = false; } else j = __ffs(gapm) - 1;
                }
                const int cp2 = GSHFL(sid, j);
                tpos = GSHFL(sid, j + 1 < nC ? j + 1 : 0);
                if (cls && walking) { cpos = cp2; cx = ps.x[cpos] - ppx; cy = ps.y[cpos] - ppy; }
            }
        }
        const bool fin = ok && closed;
        if (have && ok && !closed) PR(2);
#ifndef MVOSR_NO_WRAP_RESUME
        if (!EMIT && fv.rpool) {
            // what was certified of a star that is given up is handed over: its neighbours from the nearest one to cur (the step from
            // cur failed), as sorted positions; stars_wrap continues the walk from cur instead of starting over
            const bool part = have && !fin && nC >= 2;
            if (__any_sync(FULL, part)) {
                int rb = 0;
                if (gl == 0 && part) rb = atomicAdd(fv.rcount, nC);
                rb = GSHFL(rb, 0);
                if (part && rb + nC <= fv.rpool_cap) {
                    if (gl < nC) fv.rpool[rb + gl] = (uint16_t)sid;
                    if (gl == 0) fv.rinfo[ps.orig[p]] = ((uint32_t)rb << 8) | RING_PARTIAL | (uint32_t)nC;
                }
            }
        }
#endif
        if (have && !fin && gl == 0) {                              // (defer_hi: the last entry of defer[]; the two ends cannot meet, there are at most n stars)
            if (far) { const int slot = atomicAdd(&sc->n_defer_hi, 1); defer_hi[-slot] = (uint16_t)p; }
            else { const int slot = atomicAdd(&sc->n_defer, 1); defer[slot] = (uint16_t)p; }
        }
        // ---- consumers, both half-warps in lock step (d = 0: nothing)
        const int d = fin ? nC : 0;
        const int nid = GSHFL(sid, gl + 1 < d ? gl + 1 : 0);
        const bool tri = gl < d;
        const int op = ps.orig[p];
        if (EMIT) {
            unsigned key = 0xFFFFFFFFu;
            if (tri) {
                const int oa = ps.orig[sid], ob = ps.orig[nid];
                if (op < oa && op < ob) key = ((unsigned)min(oa, ob) << 16) | (unsigned)max(oa, ob);
            }
            const bool own = key != 0xFFFFFFFFu;
            const int k = __popc(GBALLOT(own));
            int base = 0;
            if (gl == 0 && k) base = atomicAdd(fv.T, k);
            base = GSHFL(base, 0);
            const bool over = base + k > fv.tri_cap;
            if (over && k && gl == 0) atomicOr(fv.status, MVOSR_ST_OVERFLOW);
            int rk = 0;
#pragma unroll
            for (int j = 0; j < GL; ++j) { const unsigned kj = GSHFL(key, j); rk += kj < key; }
            if (own && !over) { uint16_t *t = fv.tri + 3 * (base + rk); t[0] = (uint16_t)op; t[1] = (uint16_t)(key >> 16); t[2] = (uint16_t)(key & 0xFFFFu); }
            if (gl == 0 && k && !over) { fv.tbase[op] = (uint16_t)base; fv.tcnt[op] = (uint8_t)k; }
        } else {
            // ring store first (see consume_vote): the votes of a stored ring are taken by votes_from_rings
            bool stored = false;
            if (fv.rpool) {
                int rb = 0;
                if (gl == 0 && d > 0) rb = atomicAdd(fv.rcount, d);
                rb = GSHFL(rb, 0);
                stored = rb + d <= fv.rpool_cap;
                if (stored) {
                    if (tri) fv.rpool[rb + gl] = ps.orig[sid];
                    if (gl == 0 && d > 0) fv.rinfo[op] = ((uint32_t)rb << 8) | (uint32_t)d;
                }
            }
            if (__any_sync(FULL, d > 0 && !stored)) {                           // pool full (or no store): vote here
                bool vote = false;
                if (tri && !stored) vote = graph_vote_at(ps, fv, p, sid, nid);
                const unsigned bt = GBALLOT(tri), bv = GBALLOT(vote);
                if (gl == 0 && d > 0 && !stored && 2 * __popc(bv) > __popc(bt)) fv.pflag[op] |= 2;
            }
        }
    }
#undef GBALLOT
#undef GSHFL
#undef PR
}

// ---------------------------------------------------------------------------------------------
// thread path: one star per LANE (gthread.cuh), vote pass
// ---------------------------------------------------------------------------------------------
// A warp takes 32 consecutive stars of the pass order (the same rotated order as stars_pair: index i -> position i + rot), every lane
// walks its own (thread_star), and the warp then consumes the certified ones together: graph vote per triangle, keep flag, the ring
// store with ONE allocation per warp.  The stars it returns go where stars_pair's go: defer[] (from the end when nothing lay on the
// left inside the block).  It covers the indices below `limit` (a multiple of 32); stars_pair takes the rest with its finer grain, so
// that the warps that run out of 32-star tasks fill the tail of the pass.
constexpr int THREAD_FILL = 192;         // stars left to the pair path to balance the last wave of 32-star tasks
__device__ __noinline__ void stars_thread(const SortedSet &ps, const FrameView &fv, StarCtl *sc, uint16_t *defer, uint16_t *defer_hi, int limit) {
    const unsigned FULL = 0xFFFFFFFFu;
    const int lane = threadIdx.x & 31;
    const int rot = ps.row_start[ps.R - 1];
    for (;;) {
        int i = 0;
        if (lane == 0) i = atomicAdd(&sc->next_thread, 32);
        i = __shfl_sync(FULL, i, 0);
        if (i >= limit) break;
        int p = i + lane + rot; if (p >= ps.n) p -= ps.n; if (p >= ps.n) p -= ps.n;
        const int op = ps.orig[p];
        const bool have = op != INF16;
        TRing ring; int deg = 0, st = TS_DEFER;
        if (have) st = thread_star(ps, p, ring, deg);
        __syncwarp();
        const bool fin = have && st == TS_OK;
        if (have && !fin) {
            if (st == TS_NOCAND) { const int slot = atomicAdd(&sc->n_defer_hi, 1); defer_hi[-slot] = (uint16_t)p; }
            else { const int slot = atomicAdd(&sc->n_defer, 1); defer[slot] = (uint16_t)p; }
        }
        // ---- consumers: the vote of every triangle (p, ring[k], ring[k+1]) for p, the keep flag, the ring store
        const int d = fin ? deg : 0;
        int nv = 0, prev = fin ? tring_get(ring, 0) : 0;
        const int first = prev;
        for (int k = 0; k < d; ++k) {
            const int nxt = k + 1 < d ? tring_get(ring, k + 1) : first;
            nv += graph_vote_at(ps, fv, p, prev, nxt) ? 1 : 0;
            prev = nxt;
        }
        if (fin && 2 * nv > d) fv.pflag[op] |= 2;
        if (fv.rpool) {
            int inc = d;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) { const int t = __shfl_up_sync(FULL, inc, o); if (lane >= o) inc += t; }
            int base = 0;
            if (lane == 31 && inc) base = atomicAdd(fv.rcount, inc);
            base = __shfl_sync(FULL, base, 31);
            const int rb = base + inc - d;
            if (fin && rb + d <= fv.rpool_cap) {
                for (int k = 0; k < d; ++k) fv.rpool[rb + k] = ps.orig[tring_get(ring, k)];
                fv.rinfo[op] = ((uint32_t)rb << 8) | (uint32_t)d;
            }
        }
    }
}

// ---------------------------------------------------------------------------------------------
// level 5: stars of any degree (hubs)
// ---------------------------------------------------------------------------------------------
// fb_build keeps a star on the 32 lanes of a warp; a point with more neighbours -- the centre of a ring of features -- used to fail
// its frame (MVOSR_ST_OVERFLOW), which Qhull never does.  Such stars are rebuilt here by one warp with the star in MEMORY (scratch:
// two rings of dmax sorted positions + dmax conflict flags), the same exact incremental insertion: every point of the set is examined
// (growing windows, nearest first), screened candidate-parallel with the float64 filters against all triangles of the star, and the
// survivors inserted one by one with the exact predicates -- the conflict set of a candidate is a contiguous run of the ring, replaced
// by the candidate.  Cost O(n x degree) filter evaluations: a cold path for a rare case.  Returns STAR_OK / STAR_NONE, or
// STAR_OVERFLOW beyond dmax neighbours (dmax <= 254: 8-bit triangle counts per feature).
template <bool EMIT>
__device__ __noinline__ int hub_star(const SortedSet &ps, const FrameView &fv, int p, uint16_t *scratch, int dmax, int &n_exact) {
    const unsigned FULL = 0xFFFFFFFFu;
    const int lane = threadIdx.x & 31;
    uint16_t *A = scratch, *B = scratch + dmax, *flag = scratch + 2 * dmax;
    const double ppx = ps.x[p], ppy = ps.y[p];
    const int op = ps.orig[p];
    int d = 0, rc = STAR_OK;
    int qpos = -1, qneg = -1;                                      // collinear bootstrap, as in fb_build
    __syncwarp();
    auto rel = [&](int pos, double &x, double &y, double &l) { x = (double)ps.x[pos] - ppx; y = (double)ps.y[pos] - ppy; l = x * x + y * y; };
    auto scan = [&](int b, int e) {
        for (int base = b; base < e && rc == STAR_OK; base += 32) {
            const int pos = base + lane;
            const bool v = pos < e && pos != p && ps.orig[pos] != INF16;
            double sxl = 0, syl = 0, sll = 0;
            if (v) rel(pos, sxl, syl, sll);
            unsigned F = __ballot_sync(FULL, v);
            if (F && d >= 2) {
                bool maybe = false;
                for (int kk = 0; kk < d; ++kk) {                  // (every lane reads the same ring entries: broadcasts)
                    const int ks = A[kk], kn = A[kk + 1 < d ? kk + 1 : 0];
                    double ax = 0, ay = 0, al = 0, cx = 0, cy = 0, cl = 0;
                    if (ks != INF16) rel(ks, ax, ay, al);
                    if (kn != INF16) rel(kn, cx, cy, cl);
                    bool out;
                    if (kn == INF16) { const double l = ax * syl, r = ay * sxl; out = l - r < -3.3306690738754731e-16 * (fabs(l) + fabs(r)); }
                    else if (ks == INF16) { const double l = cx * syl, r = cy * sxl; out = l - r > 3.3306690738754731e-16 * (fabs(l) + fabs(r)); }
                    else out = det3_lift_sign_filter(ax, ay, al, cx, cy, cl, sxl, syl, sll) == 1;
                    maybe |= !out;
                }
                F &= __ballot_sync(FULL, maybe);
            }
            while (F && rc == STAR_OK) {
                const int j = __ffs(F) - 1; F &= F - 1;
                const int s = base + j;
                double sx, sy, sl; rel(s, sx, sy, sl);
                if (d == 0) {
                    // ---- bootstrap (warp-uniform): the first point off the line through p and the first candidate
                    if (qpos < 0) { qpos = s; continue; }
                    double ux, uy, ul; rel(qpos, ux, uy, ul);
                    const int o = cross_sign(ux, uy, sx, sy, n_exact);
                    if (o == 0) {
                        const bool same = (fabs(ux) >= fabs(uy)) ? ((sx > 0) == (ux > 0)) : ((sy > 0) == (uy > 0));
                        if (same) { if (fabs(sx) + fabs(sy) < fabs(ux) + fabs(uy)) qpos = s; }
                        else if (qneg < 0) qneg = s;
                        else { double nx, ny, nl; rel(qneg, nx, ny, nl); if (fabs(sx) + fabs(sy) < fabs(nx) + fabs(ny)) qneg = s; }
                        continue;
                    }
                    int ids[4], dd = 0;
                    if (o > 0) { ids[dd++] = qpos; ids[dd++] = s; if (qneg >= 0) ids[dd++] = qneg; ids[dd++] = INF16; }
                    else { if (qneg >= 0) ids[dd++] = qneg; ids[dd++] = s; ids[dd++] = qpos; ids[dd++] = INF16; }
                    if (lane == 0) for (int i = 0; i < dd; ++i) A[i] = (uint16_t)ids[i];
                    d = dd;
                    __syncwarp();
                    continue;
                }
                // ---- conflicts of s with every triangle (lanes stride over the ring), exact
                int cnt = 0, starts = 0, i0 = 0x7FFFFFFF;
                for (int k0 = 0; k0 < d; k0 += 32) {
                    const int k = k0 + lane;
                    bool c = false;
                    if (k < d) {
                        const int code = exact_conflict(ps.x, ps.y, ps.orig, p, A[k], A[k + 1 < d ? k + 1 : 0], s);
                        c = code & 1; n_exact += code >> 2;
                        flag[k] = (uint16_t)c;
                    }
                    cnt += __popc(__ballot_sync(FULL, c));
                }
                __syncwarp();
                if (!cnt) continue;
                for (int k0 = 0; k0 < d; k0 += 32) {
                    const int k = k0 + lane;
                    const bool st = k < d && flag[k] && !flag[k > 0 ? k - 1 : d - 1];
                    const unsigned m = __ballot_sync(FULL, st);
                    starts += __popc(m);
                    if (m) i0 = min(i0, k0 + __ffs(m) - 1);
                }
                if (starts != 1 || cnt >= d) { rc = STAR_INCONSISTENT; break; }
                const int nd = d - cnt + 2;
                if (nd > dmax) { rc = STAR_OVERFLOW; break; }
                for (int t = lane; t < nd; t += 32) {
                    int src = i0 + cnt + t - 1; if (src >= d) src -= d;
                    B[t] = t == 0 ? (uint16_t)s : A[src];
                }
                __syncwarp();
                uint16_t *tswap = A; A = B; B = tswap;
                d = nd;
            }
        }
    };
    // every point, in growing square windows around p (nearest first keeps the intermediate stars small)
    const int prow = row_of(ps, (float)ppy);
    double w = fmax((double)(row_yhi(ps, prow) - row_ylo(ps, prow)), 1.0e-2);
    int pr0 = 1, pr1 = 0; float pxlo = 0.f, pxhi = 0.f;
    for (int round = 0; round < 64 && rc == STAR_OK; ++round, w *= 2.0) {
        const int r0 = row_of(ps, (float)(ppy - w)), r1 = row_of(ps, (float)(ppy + w));
        const float xlo = (float)(ppx - w) - 1.0e-3f, xhi = (float)(ppx + w) + 1.0e-3f;
        for (int row = r0; row <= r1 && rc == STAR_OK; ++row) {
            const int ia = row_lower(ps, row, xlo), ib = row_upper(ps, row, xhi);
            int ea = ia, eb = ia;
            if (row >= pr0 && row <= pr1) { ea = lower_x(ps.x, ia, ib, pxlo); eb = upper_x(ps.x, ea, ib, pxhi); }
            scan(ia, ea);
            if (rc == STAR_OK) scan(eb, ib);
        }
        pr0 = r0; pr1 = r1; pxlo = xlo; pxhi = xhi;
        if (xlo <= ps.xmin && xhi >= ps.xmax && r0 == 0 && r1 == ps.R - 1) break;
    }
    if (rc != STAR_OK) return rc;
    if (d == 0) return STAR_NONE;
    // ---- consumers, from the ring in memory
    auto nxt = [&](int k) { return (int)A[k + 1 < d ? k + 1 : 0]; };
    if (!EMIT) {
        bool stored = false;
        if (fv.rpool && d < (int)RING_PARTIAL) {
            int rb = 0;
            if (lane == 0) rb = atomicAdd(fv.rcount, d);
            rb = __shfl_sync(FULL, rb, 0);
            if (rb + d <= fv.rpool_cap) {
                for (int k = lane; k < d; k += 32) fv.rpool[rb + k] = A[k] != INF16 ? ps.orig[A[k]] : INF16;
                __syncwarp();
                if (lane == 0) fv.rinfo[op] = ((uint32_t)rb << 8) | (uint32_t)d;
                stored = true;                                     // votes_from_rings takes the votes
            }
        }
        if (!stored) {
            int nt = 0, nv = 0;
            for (int k0 = 0; k0 < d; k0 += 32) {
                const int k = k0 + lane;
                const bool tri = k < d && A[k] != INF16 && nxt(k) != INF16;
                const bool vote = tri && graph_vote_at(ps, fv, p, A[k], nxt(k));
                nt += __popc(__ballot_sync(FULL, tri)); nv += __popc(__ballot_sync(FULL, vote));
            }
            if (lane == 0 && 2 * nv > nt) fv.pflag[op] |= 2;
        }
    } else {
        // triangles (op < oa, ob) as one block sorted by (min, max) of the other two vertices; keys go through the flag array
        int kown = 0;
        for (int k0 = 0; k0 < d; k0 += 32) {
            const int k = k0 + lane;
            bool own = false;
            if (k < d && A[k] != INF16 && nxt(k) != INF16) { const int oa = ps.orig[A[k]], ob = ps.orig[nxt(k)]; own = op < oa && op < ob; }
            if (k < d) flag[k] = (uint16_t)own;
            kown += __popc(__ballot_sync(FULL, own));
        }
        __syncwarp();
        if (kown) {
            int base = 0;
            if (lane == 0) base = atomicAdd(fv.T, kown);
            base = __shfl_sync(FULL, base, 0);
            if (base + kown > fv.tri_cap) { if (lane == 0) atomicOr(fv.status, MVOSR_ST_OVERFLOW); }
            else {
                auto keyof = [&](int k) { const unsigned oa = ps.orig[A[k]], ob = ps.orig[nxt(k)]; return (min(oa, ob) << 16) | max(oa, ob); };
                for (int k = lane; k < d; k += 32) {
                    if (!flag[k]) continue;
                    const unsigned key = keyof(k);
                    int r = 0;
                    for (int j = 0; j < d; ++j) if (flag[j] && keyof(j) < key) ++r;
                    uint16_t *t = fv.tri + 3 * (base + r); t[0] = (uint16_t)op; t[1] = (uint16_t)(key >> 16); t[2] = (uint16_t)(key & 0xFFFFu);
                }
                if (lane == 0) { fv.tbase[op] = (uint16_t)base; fv.tcnt[op] = (uint8_t)kown; }
            }
        }
    }
    __syncwarp();                                                  // the next hub of the frame reuses the scratch
    return STAR_OK;
}

// All stars of the staged point set.  EMIT: triangles into fv.tri; otherwise the graph vote into fv.pflag.
// Block-wide; sc, defer[] and defer2[] are shared scratch.  Four levels: pair path (all stars) -> wrap path (hull stars,
// circles leaving the block) -> exact half-warp path (what float32 could not certify) -> exact full-warp path (more than
// 16 slots, collinear bootstrap).  Returns (#stars of level 3) + (#stars of level 4 << 16); *n_wrap receives level 2's count.
template <bool EMIT>
__device__ __noinline__ int run_stars(const SortedSet &ps, const FrameView &fv, StarCtl *sc, uint16_t *defer, uint16_t *defer2,
                                      int &n_exact, long long *t_fast, const uint16_t *todo = nullptr, int n_todo = 0) {
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    // vote pass over all stars: 32-star tasks of the thread path first, the rest (and small frames) on the pair path
    int limit = 0;
#ifdef MVOSR_THREAD_PATH
    if (!EMIT && !todo) limit = max(ps.n - THREAD_FILL, 0) & ~31;
#endif
    if (tid == 0) { sc->next_pos = limit; sc->next_thread = 0; sc->n_defer = 0; sc->n_defer_hi = 0; sc->n_defer2 = 0; sc->n_hub = 0; }
    __syncthreads();
    long long tc0 = clock64();
    uint16_t *const defer_hi = defer2 - 1;                       // defer[] is filled from both ends (defer2 = defer + cap)
    if (limit) stars_thread(ps, fv, sc, defer, defer_hi, limit);
    stars_pair<EMIT>(ps, fv, sc, defer, defer_hi, todo, n_todo);
    __syncthreads();
    if (tid == 0 && t_fast) *t_fast += clock64() - tc0;
    const int n1l = sc->n_defer, n1h = sc->n_defer_hi, n1 = n1l + n1h;
    __syncthreads();
    if (tid == 0) sc->next_pos = 0;
    __syncthreads();
    if (n1) stars_wrap<EMIT>(ps, fv, sc, defer, n1l, defer2, defer_hi, n1h);
    __syncthreads();
    const int n2 = sc->n_defer2;
    __syncthreads();
    if (tid == 0) { sc->next_pos = 0; sc->n_defer2 = 0; }
    __syncthreads();
    if (n2) stars_fast<EMIT>(ps, fv, sc, defer2, n2, defer, n_exact);       // overflow list: defer[] again (its first use is over)
    __syncthreads();
    const int n3 = sc->n_defer2;
    for (int k = warp; k < n3; k += NWARP) {
        const int p = defer[k];
        const FbResult r = fb_build(ps, p);
        n_exact += r.n_exact;
        if (r.rc == STAR_OK) {
            if (EMIT) consume_emit<32>(0xFFFFFFFFu, lane, r.d, p, r.sid, r.nid, ps, fv);
            else consume_vote<32>(0xFFFFFFFFu, lane, r.d, p, r.sid, r.nid, ps, fv);
        } else if (r.rc == STAR_OVERFLOW) {                          // more than 32 neighbours: level 5, below
            if (lane == 0) { const int slot = atomicAdd(&sc->n_hub, 1); defer[n3 + slot] = (uint16_t)p; }
        } else if (r.rc != STAR_NONE && lane == 0) {
            atomicOr(fv.status, MVOSR_ST_OVERFLOW);
#ifdef MVOSR_DEBUG_PRINT
            printf("fb_build failed: rc=%d p=%d orig=%d d=%d n=%d\n", r.rc, p, (int)ps.orig[p], r.d, ps.n);
#endif
        }
    }
    __syncthreads();
    const int n4 = sc->n_hub;
    if (n4 && warp == 0) {
        // hubs, one after the other on warp 0: the star lives in defer2[] (dead by now), two rings and the conflict flags
        const int dmax = min(254, (int)(defer2 - defer) / 3);
        for (int k = 0; k < n4; ++k) {
            const int p = defer[n3 + k];
            const int rc = hub_star<EMIT>(ps, fv, p, defer2, dmax, n_exact);
            if (rc != STAR_OK && rc != STAR_NONE && lane == 0) atomicOr(fv.status, MVOSR_ST_OVERFLOW);
        }
    }
    __syncthreads();
    if (tid == 0) sc->n_wrap += n1;
    return n2 + (n3 << 16);
}

}  // namespace mvosr
