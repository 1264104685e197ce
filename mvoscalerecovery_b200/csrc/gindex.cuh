// The strip-sorted point set of one frame and its neighbourhood queries: scalar code shared by every star path (gstrip.cuh, gthread.cuh).
// __host__ __device__ throughout and free of warp intrinsics, so that tests/host_sim can run the per-lane star builder on the CPU.
#pragma once
#include <stdint.h>
#include <math.h>
#include <string.h>

#ifndef MVOSR_HD
#ifdef __CUDACC__
#define MVOSR_HD __host__ MVOSR_HD
#else
#define MVOSR_HD inline
#endif
#endif

#ifndef __CUDACC__
#include <algorithm>
struct float2 { float x, y; };
#endif

namespace mvosr {
#ifndef __CUDACC__
using std::min; using std::max;
#endif

#define MVOSR_INFF (__builtin_huge_valf())
MVOSR_HD unsigned f2u(float t) {
#ifdef __CUDA_ARCH__
    return __float_as_uint(t);
#else
    unsigned k; memcpy(&k, &t, 4); return k;
#endif
}
MVOSR_HD float u2f(unsigned k) {
#ifdef __CUDA_ARCH__
    return __uint_as_float(k);
#else
    float t; memcpy(&t, &k, 4); return t;
#endif
}

constexpr uint16_t INF16 = 0xFFFF;
constexpr float GRID_DENSITY = 1.5f;     // k: a strip is closed when count x height >= k x extent (k points per square cell)
constexpr int WIN_M = 4;                 // the local spacing in a point's own strip is measured over WIN_M neighbours on either side
constexpr float WIN_FACTOR = 2.5f;       // half-width of the candidate window in local cell sides (the block is +-2 strips tall)
// Forward error bound of the float32 in-circle evaluation: inputs are float32 roundings of exact
// differences (relative error u = 2^-24); every term of the expanded determinant accumulates at most
// 11u (4 inputs, <= 7 roundings), so |det_fl - det| <= 11u * perm.  KERR = 16u leaves room for the
// rounding of the bound itself.
constexpr float KERR = 9.5367431640625e-07f;   // 2^-20
constexpr float WU = 5.9604644775390625e-08f;  // 2^-24, unit roundoff of float32

// The staged point set: entries sorted by strip (bottom-up in y), inside a strip by (x, original index); an exact duplicate
// of an earlier point stays in place as a hole (orig == INF16).  Strip r is the union of the histogram bins row_bin[r] ..
// row_bin[r+1]-1 of width bh starting at ymin; a point's strip is bin_row[bin_of(y)].
struct SortedSet {
    const float *x, *y;                  // [n] pixel coordinates (float32; |x| < 4096, multiples of 2^-40)
    const uint16_t *orig;                // [n] index in the frame's feature order, INF16 for a hole
    const uint16_t *row_start;           // [R+1] first entry of every strip
    const uint16_t *row_bin;             // [R+1] first histogram bin of every strip
    const uint16_t *bin_row;             // [NB] strip of every bin
    // every strip is cut into uniform sub-cells of about two points (the unit of the counting sort that built it): the first
    // position with x >= v is found in the sub-cell of v, cell_start[row_cell[r] + (int)((v - x0) * inv)], x0 / inv = row_xi[r]
    const uint16_t *row_cell;            // [R+1] first sub-cell of every strip
    const float2 *row_xi;                // [R] (x origin, 1 / sub-cell width) of every strip
    const uint16_t *cell_start;          // [row_cell[R]+1] first entry of every sub-cell
    int n, R, NB, win_m;
    float xmin, xmax, ymin, ymax, bh, inv_bh, kdens, wfac;
};
// order-preserving float <-> uint (atomicMin / atomicMax on float coordinates, REDUX on float keys)
MVOSR_HD unsigned fkey(float t) { const unsigned k = f2u(t); return (k & 0x80000000u) ? ~k : (k | 0x80000000u); }
MVOSR_HD float funkey(unsigned k) { return u2f((k & 0x80000000u) ? (k & 0x7FFFFFFFu) : ~k); }

MVOSR_HD int bin_of(const SortedSet &ps, float y) {
    int b = (int)((y - ps.ymin) * ps.inv_bh);
    return b < 0 ? 0 : (b >= ps.NB ? ps.NB - 1 : b);
}
MVOSR_HD int row_of(const SortedSet &ps, float y) { return ps.bin_row[bin_of(ps, y)]; }
// strip r holds exactly the points with ylo(r) <= y < yhi(r) up to the rounding of bin_of (callers shrink by STRIP_SLACK)
MVOSR_HD float row_ylo(const SortedSet &ps, int r) { return ps.ymin + (float)ps.row_bin[r] * ps.bh; }
MVOSR_HD float row_yhi(const SortedSet &ps, int r) { return ps.ymin + (float)ps.row_bin[r + 1] * ps.bh; }
MVOSR_HD float strip_slack(const SortedSet &ps) { return 1.0e-3f + 1.0e-4f * ps.bh; }
// first position in [b, e) whose x is >= v (lower) / > v (upper); the strip is sorted by x
MVOSR_HD int lower_x(const float *x, int b, int e, float v) {
    while (b < e) { const int m = (b + e) >> 1; if (x[m] < v) b = m + 1; else e = m; }
    return b;
}
MVOSR_HD int upper_x(const float *x, int b, int e, float v) {
    while (b < e) { const int m = (b + e) >> 1; if (x[m] <= v) b = m + 1; else e = m; }
    return b;
}

// single-instruction approximations (MUFU, 2 ulp); every use below is covered by explicit padding
MVOSR_HD float rcp_approx(float x) {
#ifdef __CUDA_ARCH__
    float r; asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x)); return r;
#else
    return 1.0f / x;
#endif
}
MVOSR_HD float sqrt_approx(float x) {
#ifdef __CUDA_ARCH__
    float r; asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x)); return r;
#else
    return sqrtf(x);
#endif
}

// sub-cell of x in a strip of nc sub-cells (monotone in x: the float product and the truncation are)
MVOSR_HD int strip_cell(float2 xi, int nc, float x) {
    const int c = (int)((x - xi.x) * xi.y);                      // (the conversion saturates; NaN -> 0)
    return c < 0 ? 0 : (c >= nc ? nc - 1 : c);
}
// first position of strip `row` whose x is >= v (row_lower) / > v (row_upper)
MVOSR_HD int row_lower(const SortedSet &ps, int row, float v) {
    const int o = ps.row_cell[row], c = o + strip_cell(ps.row_xi[row], ps.row_cell[row + 1] - o, v);
    return lower_x(ps.x, ps.cell_start[c], ps.cell_start[c + 1], v);
}
MVOSR_HD int row_upper(const SortedSet &ps, int row, float v) {
    const int o = ps.row_cell[row], c = o + strip_cell(ps.row_xi[row], ps.row_cell[row + 1] - o, v);
    return upper_x(ps.x, ps.cell_start[c], ps.cell_start[c + 1], v);
}

// The candidate block of the point at sorted position p: its own strip and the two above and below (strips are about one local
// cell side tall by construction), cut to the x-window [xlo, xhi] (absolute coordinates) of half-width WIN_FACTOR local cell
// sides -- the side c of the square that holds k points at the local density: c^2 = k x (strip height) x (spacing along x in
// the strip, measured over the WIN_M neighbours on either side).  `open` bits (1 left, 2 right, 4 below, 8 above): no point of
// the whole set lies beyond that side.
// (Measured and dropped: choosing the NUMBER of strips from c as well -- fewer strips inside a dense cluster of a tall strip,
// up to eight in the sparse part of a thin one.  The noise of the local spacing then enters the candidate count squared; on all
// three bench densities more stars left the pair path than with the fixed five strips.)
constexpr int BLOCK_ROWS = 5;
constexpr int WRAP_ROWS = 7;             // strips of the (larger) block of the one-warp-per-star path
constexpr float WRAP_WIDEN = 1.35f;      // ... and its window relative to the pair path's
struct Block { int row, r0, r1; float xlo, xhi; int open; };
MVOSR_HD Block block_of(const SortedSet &ps, int p, float ppx, float ppy) {
    Block k;
    k.row = row_of(ps, ppy);
    k.r0 = max(k.row - BLOCK_ROWS / 2, 0); k.r1 = min(k.row + BLOCK_ROWS / 2, ps.R - 1);
    const int rb = ps.row_start[k.row], re = ps.row_start[k.row + 1];
    const int ia = max(p - ps.win_m, rb), ib = min(p + ps.win_m, re - 1);
    const float span = ps.x[ib] - ps.x[ia], H = row_yhi(ps, k.row) - row_ylo(ps, k.row);
    // (the window is a heuristic: single-instruction reciprocal and square root; the box and its candidates both follow from it)
    const float side = (ib > ia && span > 0.f) ? sqrt_approx(ps.kdens * H * span * rcp_approx((float)(ib - ia))) : H;
    const float w = ps.wfac * fmaxf(side, 1.0e-3f);
    k.xlo = ppx - w; k.xhi = ppx + w;
    k.open = (k.xlo <= ps.xmin ? 1 : 0) | (k.xhi >= ps.xmax ? 2 : 0) | (k.r0 == 0 ? 4 : 0) | (k.r1 == ps.R - 1 ? 8 : 0);
    return k;
}

// Does the walk's-left cap of the padded circle (centre v, radius rs, through p = origin and cur) lie inside the box?
// The cap's bounding box is spanned by p, cur and those axis-extreme points of the circle that lie on the left of p->cur.
// CLIP: only the part of the cap inside the bounding box G of the point set matters (there is nothing to find outside it),
// so the box is intersected with the box of (disk n G) -- this lets the flat triangles along the boundary of the point set
// pass, whose circles are huge but only a thin sliver of them lies inside G.  Measured on the bench workload it keeps
// ~40 stars per frame out of streaming but the extra instructions per step cost more than that saves (pair path +13 %,
// wrap path +3 %), so both paths run with CLIP = false; the variant is kept for sparser inputs.
struct WBox { float x0, x1, y0, y1; };
#ifndef MVOSR_CAP_CLIP
#define MVOSR_CAP_CLIP 0                 // 1: clip the cap test to the bounding box of the point set.  Measured again in round 2 behind the quick accept: uniform -0.5 %, perspective -2.5 %, clustered -5 % (fewer streaming stars, dearer slow path)
#endif
constexpr bool CAP_CLIP = MVOSR_CAP_CLIP != 0;
// bounding box of the point set relative to p (a superset is safe)
MVOSR_HD WBox set_box(const SortedSet &ps, float ppx, float ppy) {
    WBox G; G.x0 = ps.xmin - ppx - 1.0e-3f; G.x1 = ps.xmax - ppx + 1.0e-3f; G.y0 = ps.ymin - ppy - 1.0e-3f; G.y1 = ps.ymax - ppy + 1.0e-3f;
    return G;
}
template <bool CLIP>
MVOSR_HD bool w_cap_inside(float cx, float cy, float sigma, float vx, float vy, float rs,
                                             float BX0, float BX1, float BY0, float BY1, const WBox &G) {
    const float tol = 1.0e-4f * (fabsf(cx) + fabsf(cy)) * (rs + fabsf(vx) + fabsf(vy));       // include when in doubt
    float lox = fminf(0.f, cx), hix = fmaxf(0.f, cx), loy = fminf(0.f, cy), hiy = fmaxf(0.f, cy);
    if (sigma * (cx * vy - cy * (vx - rs)) > -tol) lox = fminf(lox, vx - rs);
    if (sigma * (cx * vy - cy * (vx + rs)) > -tol) hix = fmaxf(hix, vx + rs);
    if (sigma * (cx * (vy - rs) - cy * vx) > -tol) loy = fminf(loy, vy - rs);
    if (sigma * (cx * (vy + rs) - cy * vx) > -tol) hiy = fmaxf(hiy, vy + rs);
    if (CLIP) {
        if (rs < 1.0e6f) {
            // half-widths of the disk inside the strips G.y0..G.y1 and G.x0..G.x1 ((rs-d)(rs+d): no cancellation)
            const float dy = fmaxf(fmaxf(G.y0 - vy, vy - G.y1), 0.f), dx = fmaxf(fmaxf(G.x0 - vx, vx - G.x1), 0.f);
            const float hwx = sqrt_approx(fmaxf((rs - dy) * (rs + dy), 0.f)) * 1.001f + 1.0e-3f;
            const float hwy = sqrt_approx(fmaxf((rs - dx) * (rs + dx), 0.f)) * 1.001f + 1.0e-3f;
            lox = fmaxf(lox, vx - hwx); hix = fminf(hix, vx + hwx); loy = fmaxf(loy, vy - hwy); hiy = fminf(hiy, vy + hwy);
        }
        lox = fmaxf(lox, G.x0); hix = fminf(hix, G.x1); loy = fmaxf(loy, G.y0); hiy = fminf(hiy, G.y1);
    }
    return lox >= BX0 && hix <= BX1 && loy >= BY0 && hiy <= BY1;
}


}  // namespace mvosr
