// Exact geometric predicates for the per-frame Delaunay stage (device code, sm_100a).
//
// Inputs are float32 pixel coordinates with |x| < 4096 that are multiples of 2^-40 (every
// float32 of magnitude >= 2^-17, and 0; smaller magnitudes are flushed to 0 at load time).
// Differences of two such values are exact in float64 (<= 54 significant bits incl. sign
// are not needed: the difference is a multiple of 2^-40 below 2^13, i.e. <= 53 bits).
// Every predicate takes such exact differences ("relative coordinates"), evaluates the
// determinant in float64 with a forward error bound (Shewchuk's stage-A bounds) and falls
// through to exact integer arithmetic (int128 / 256-bit) only when the bound cannot decide.
//
// Replaces: the non-exact FP64 predicates of Qhull 8.0.2 behind scipy.spatial.Delaunay at
// reference src/rescale.py:124,136.  For points in general position both give THE Delaunay
// triangulation; co-circular ties are broken symbolically (see incircle_sos below).
#pragma once
#include <stdint.h>
#include <math.h>

// The predicates and the thread-path star code are plain sequential C++ and also compile for the
// host, so that tests can run the very same source on the CPU (tests/host_sim) where no GPU exists.
#if defined(__CUDACC__)
#define MVOSR_HD __host__ __device__ __forceinline__
#define MVOSR_HD_NOINLINE __host__ __device__ __noinline__
#else
#define MVOSR_HD inline
#define MVOSR_HD_NOINLINE inline
#endif
#if !defined(__CUDA_ARCH__)
static inline uint64_t mvosr_umul64hi(uint64_t a, uint64_t b) { return (uint64_t)(((unsigned __int128)a * b) >> 64); }
static inline int mvosr_popc(uint32_t x) { return __builtin_popcount(x); }
static inline int mvosr_ffs(uint32_t x) { return __builtin_ffs((int)x); }
#else
#define mvosr_umul64hi __umul64hi
#define mvosr_popc __popc
#define mvosr_ffs __ffs
#endif

namespace mvosr {

struct S256 { uint64_t w[4]; };   // two's complement, little endian

MVOSR_HD int64_t fix40(double d) { return (int64_t)(d * 1099511627776.0); }  // exact by construction

MVOSR_HD void mul128(uint64_t a0, uint64_t a1, uint64_t b0, uint64_t b1, uint64_t r[4]) {
    // (a1:a0) * (b1:b0) -> 256 bit, schoolbook on 64-bit limbs
    uint64_t p00l = a0 * b0, p00h = mvosr_umul64hi(a0, b0);
    uint64_t p01l = a0 * b1, p01h = mvosr_umul64hi(a0, b1);
    uint64_t p10l = a1 * b0, p10h = mvosr_umul64hi(a1, b0);
    uint64_t p11l = a1 * b1, p11h = mvosr_umul64hi(a1, b1);
    r[0] = p00l;
    unsigned __int128 mid = (unsigned __int128)p00h + p01l + p10l;
    r[1] = (uint64_t)mid;
    unsigned __int128 hi = (mid >> 64) + p01h + p10h + p11l;
    r[2] = (uint64_t)hi;
    r[3] = (uint64_t)(hi >> 64) + p11h;
}

MVOSR_HD void neg256(uint64_t r[4]) {
    uint64_t c = 1;
#pragma unroll
    for (int i = 0; i < 4; ++i) { uint64_t v = ~r[i]; uint64_t s = v + c; c = (c && s == 0) ? 1 : 0; r[i] = s; }
}

MVOSR_HD void add256(uint64_t a[4], const uint64_t b[4]) {
    unsigned __int128 c = 0;
#pragma unroll
    for (int i = 0; i < 4; ++i) { unsigned __int128 s = (unsigned __int128)a[i] + b[i] + c; a[i] = (uint64_t)s; c = s >> 64; }
}

MVOSR_HD_NOINLINE void mul_i128(__int128 a, __int128 b, uint64_t r[4]) {
    bool neg = (a < 0) != (b < 0);
    unsigned __int128 ua = a < 0 ? (unsigned __int128)(-a) : (unsigned __int128)a;
    unsigned __int128 ub = b < 0 ? (unsigned __int128)(-b) : (unsigned __int128)b;
    mul128((uint64_t)ua, (uint64_t)(ua >> 64), (uint64_t)ub, (uint64_t)(ub >> 64), r);
    if (neg) neg256(r);
}

// ---- sign of ux*vy - uy*vx ------------------------------------------------------------------
MVOSR_HD_NOINLINE int cross_sign_exact(double ux, double uy, double vx, double vy) {
    __int128 l = (__int128)fix40(ux) * (__int128)fix40(vy);
    __int128 r = (__int128)fix40(uy) * (__int128)fix40(vx);
    return (l > r) - (l < r);
}

MVOSR_HD int cross_sign(double ux, double uy, double vx, double vy, int &n_exact) {
    double l = ux * vy, r = uy * vx;
    double det = l - r;
    double err = 3.3306690738754731e-16 * (fabs(l) + fabs(r));
    if (det > err) return 1;
    if (det < -err) return -1;
    ++n_exact;
    return cross_sign_exact(ux, uy, vx, vy);
}

// ---- sign of the 3x3 determinant | a al ; b bl ; c cl | (rows (x, y, x^2+y^2)) -----------------
// With the fourth point at the origin the standard in-circle determinant of (o, a, b, c) is the
// NEGATIVE of this one (expand the 4x4 along the zero row).
MVOSR_HD_NOINLINE int det3_lift_sign_exact(double ax, double ay, double bx, double by, double cx, double cy) {
    __int128 AX = fix40(ax), AY = fix40(ay), BX = fix40(bx), BY = fix40(by), CX = fix40(cx), CY = fix40(cy);
    __int128 al = AX * AX + AY * AY, bl = BX * BX + BY * BY, cl = CX * CX + CY * CY;
    __int128 ma = BX * CY - BY * CX;      // cofactor of al
    __int128 mb = CX * AY - CY * AX;      // cofactor of bl  ( -(AX*CY - AY*CX) )
    __int128 mc = AX * BY - AY * BX;      // cofactor of cl
    uint64_t acc[4], t[4];
    mul_i128(al, ma, acc);
    mul_i128(bl, mb, t); add256(acc, t);
    mul_i128(cl, mc, t); add256(acc, t);
    if (acc[3] >> 63) return -1;
    return (acc[0] | acc[1] | acc[2] | acc[3]) ? 1 : 0;
}

// float64 filter; returns +-1 when certain, 2 when undecided
MVOSR_HD int det3_lift_sign_filter(double ax, double ay, double al, double bx, double by, double bl,
                                                     double cx, double cy, double cl) {
    double bxcy = bx * cy, bycx = by * cx;
    double cxay = cx * ay, cyax = cy * ax;
    double axby = ax * by, aybx = ay * bx;
    double det = al * (bxcy - bycx) + bl * (cxay - cyax) + cl * (axby - aybx);
    double perm = al * (fabs(bxcy) + fabs(bycx)) + bl * (fabs(cxay) + fabs(cyax)) + cl * (fabs(axby) + fabs(aybx));
    double err = 2.0e-15 * perm;            // Shewchuk's iccerrboundA is 1.11e-15; slack for contraction order
    if (det > err) return 1;
    if (det < -err) return -1;
    return 2;
}

// ---- in-circle with symbolic tie-break ---------------------------------------------------------
// Star triangle (p, q0, q1), counter-clockwise, candidate s; all coordinates relative to p (exact).
// Returns true iff s is strictly inside the circumcircle of (p,q0,q1) under the perturbation
//   lift_i' = lift_i + eps_i,  eps_0 >> eps_1 >> ... > 0   (smaller index = larger perturbation):
// if the exact determinant vanishes its sign is that of the cofactor of the smallest-index point:
//   a=p: +orient(q0,q1,s)   b=q0: -orient(p,q1,s)   c=q1: +orient(p,q0,s)   d=s: -orient(p,q0,q1) (<0).
MVOSR_HD bool incircle_sos(double q0x, double q0y, double q0l, double q1x, double q1y, double q1l,
                                             double sx, double sy, double sl,
                                             int ip, int iq0, int iq1, int is, int &n_exact) {
    int f = det3_lift_sign_filter(q0x, q0y, q0l, q1x, q1y, q1l, sx, sy, sl);
    if (f != 2) return f < 0;                // in-circle = -det3
    ++n_exact;
    int e = det3_lift_sign_exact(q0x, q0y, q1x, q1y, sx, sy);
    if (e != 0) return e < 0;
    int k = ip, which = 0;
    if (iq0 < k) { k = iq0; which = 1; }
    if (iq1 < k) { k = iq1; which = 2; }
    if (is < k) { which = 3; }
    switch (which) {
        case 0: return cross_sign_exact(q1x - q0x, q1y - q0y, sx - q0x, sy - q0y) > 0;   // orient(q0,q1,s); differences exact
        case 1: return cross_sign_exact(q1x, q1y, sx, sy) < 0;                            // -orient(p,q1,s)
        case 2: return cross_sign_exact(q0x, q0y, sx, sy) > 0;                            // +orient(p,q0,s)
        default: return false;                                                            // -orient(p,q0,q1) < 0
    }
}

// s strictly between p (origin) and q on their common line (collinearity already established)
MVOSR_HD bool strictly_between(double qx, double qy, double sx, double sy) {
    if (fabs(qx) >= fabs(qy)) return (sx > 0) == (qx > 0) && sx != 0 && fabs(sx) < fabs(qx);
    return (sy > 0) == (qy > 0) && sy != 0 && fabs(sy) < fabs(qy);
}

}  // namespace mvosr
