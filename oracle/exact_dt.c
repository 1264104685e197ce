/* Exact 2-D Delaunay triangulation, brute force -- TEST INFRASTRUCTURE (oracle side).
 *
 * Part of oracle/: compiled by oracle/Makefile into oracle/_build/liboracle.so and loaded
 * only by tests/, __graft_entry__.smoke() and bench.py's CPU-baseline leg.  The product
 * has its own, independently written device implementation (csrc/gstar.cuh, csrc/predicates.cuh).
 *
 * What it restates: the triangle set scipy.spatial.Delaunay (Qhull 8.0.2, options
 * "Qbb Qc Qz Q12" + Qt) hands to the reference at src/rescale.py:124-125,136-137.
 * Qhull is third-party, absent from /root/reference, non-exact FP64 with facet merging;
 * for inputs in general position it returns THE Delaunay triangulation, which is unique,
 * so any exact algorithm reproduces its (canonicalised) simplex set.  For exactly
 * co-circular inputs the triangulation is not unique and Qhull's choice is arbitrary
 * (SURVEY.md H1); there this oracle defines the tie-break the CUDA path must follow:
 *
 *   symbolic perturbation of the lifted coordinate, h_i' = x_i^2 + y_i^2 + eps_i with
 *   eps_0 >> eps_1 >> ... > 0 (smaller index = larger perturbation).  When the exact
 *   in-circle determinant of (a,b,c,d) is zero its sign is the sign of the cofactor of
 *   the smallest-index point among the four: +orient(b,c,d), -orient(a,c,d),
 *   +orient(a,b,d), -orient(a,b,c) for a,b,c,d respectively.
 *
 * With that rule the (perturbed) triangulation is unique, independent of the order of
 * construction, so oracle and GPU agree as sets on ANY input.  Duplicate points: the
 * lowest index is kept, later copies appear in no triangle (Qhull also drops them, into
 * .coplanar).
 *
 * Arithmetic: inputs are float32 pixel coordinates with |x| < 4096 that are multiples of
 * 2^-40 (every float32 with |x| >= 2^-17, and 0); scaled by 2^40 they are integers below
 * 2^52, differences fit in 54 bits, orient2d in __int128 and in-circle in 256 bits.
 * Everything here is exact integer arithmetic -- no floating-point filter.
 *
 * Algorithm: for every point p build its star (cyclic CCW neighbour list, with one
 * INF marker on hull points standing for the two ghost triangles) by inserting every
 * other point in index order and removing the arc of star triangles in conflict
 * (Bowyer-Watson restricted to one star).  O(n^2 * degree): an oracle, not a product.
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <math.h>

typedef __int128 i128;
typedef unsigned __int128 u128;

#define INF_ID (-1)
#define MAX_DEG 256

typedef struct { uint64_t w[4]; } s256;      /* two's complement, little-endian limbs */

static s256 s256_from_mul(i128 a, i128 b) {
    int neg = (a < 0) != (b < 0);
    u128 ua = a < 0 ? (u128)(-a) : (u128)a;
    u128 ub = b < 0 ? (u128)(-b) : (u128)b;
    uint64_t a0 = (uint64_t)ua, a1 = (uint64_t)(ua >> 64);
    uint64_t b0 = (uint64_t)ub, b1 = (uint64_t)(ub >> 64);
    u128 p00 = (u128)a0 * b0, p01 = (u128)a0 * b1, p10 = (u128)a1 * b0, p11 = (u128)a1 * b1;
    s256 r;
    r.w[0] = (uint64_t)p00;
    u128 mid = (p00 >> 64) + (uint64_t)p01 + (uint64_t)p10;
    r.w[1] = (uint64_t)mid;
    u128 hi = (mid >> 64) + (p01 >> 64) + (p10 >> 64) + (uint64_t)p11;
    r.w[2] = (uint64_t)hi;
    r.w[3] = (uint64_t)((hi >> 64) + (p11 >> 64));
    if (neg) {                                /* two's complement negate */
        unsigned carry = 1;
        for (int i = 0; i < 4; ++i) {
            uint64_t v = ~r.w[i];
            uint64_t s = v + carry;
            carry = (carry && s == 0) ? 1 : 0;
            r.w[i] = s;
        }
    }
    return r;
}

static s256 s256_add(s256 a, s256 b) {
    s256 r;
    unsigned carry = 0;
    for (int i = 0; i < 4; ++i) {
        u128 s = (u128)a.w[i] + b.w[i] + carry;
        r.w[i] = (uint64_t)s;
        carry = (unsigned)(s >> 64);
    }
    return r;
}

static s256 s256_neg(s256 a) {
    unsigned carry = 1;
    for (int i = 0; i < 4; ++i) {
        uint64_t v = ~a.w[i];
        uint64_t s = v + carry;
        carry = (carry && s == 0) ? 1 : 0;
        a.w[i] = s;
    }
    return a;
}

static int s256_sign(s256 a) {
    if (a.w[3] >> 63) return -1;
    return (a.w[0] | a.w[1] | a.w[2] | a.w[3]) ? 1 : 0;
}

/* sign of orient2d(a,b,c): >0 when a,b,c are counter-clockwise (y up; only consistency matters) */
static int orient_sign(const int64_t *X, const int64_t *Y, int a, int b, int c) {
    i128 l = (i128)(X[a] - X[c]) * (i128)(Y[b] - Y[c]);
    i128 r = (i128)(Y[a] - Y[c]) * (i128)(X[b] - X[c]);
    return (l > r) - (l < r);
}

/* sign of the in-circle determinant of (a,b,c,d), a,b,c CCW => >0 iff d strictly inside.
 * Exact; ties broken by the symbolic perturbation described in the header. */
static int incircle_sign(const int64_t *X, const int64_t *Y, int a, int b, int c, int d) {
    i128 adx = X[a] - X[d], ady = Y[a] - Y[d];
    i128 bdx = X[b] - X[d], bdy = Y[b] - Y[d];
    i128 cdx = X[c] - X[d], cdy = Y[c] - Y[d];
    i128 al = adx * adx + ady * ady, bl = bdx * bdx + bdy * bdy, cl = cdx * cdx + cdy * cdy;
    i128 ma = bdx * cdy - cdx * bdy;
    i128 mb = cdx * ady - adx * cdy;
    i128 mc = adx * bdy - bdx * ady;
    s256 det = s256_add(s256_add(s256_from_mul(al, ma), s256_from_mul(bl, mb)), s256_from_mul(cl, mc));
    int s = s256_sign(det);
    if (s) return s;
    /* co-circular: cofactor of the smallest index */
    int k = a; int which = 0;
    if (b < k) { k = b; which = 1; }
    if (c < k) { k = c; which = 2; }
    if (d < k) { k = d; which = 3; }
    switch (which) {
        case 0: return orient_sign(X, Y, b, c, d);
        case 1: return -orient_sign(X, Y, a, c, d);
        case 2: return orient_sign(X, Y, a, b, d);
        default: return -orient_sign(X, Y, a, b, c);
    }
}

/* s strictly between p and q on their common line (p,q,s known collinear) */
static int strictly_between(const int64_t *X, const int64_t *Y, int p, int q, int s) {
    i128 dx = X[q] - X[p], dy = Y[q] - Y[p];
    i128 sx = X[s] - X[p], sy = Y[s] - Y[p];
    i128 dot = dx * sx + dy * sy;           /* fits: 2 * 2^106 */
    i128 len = dx * dx + dy * dy;
    return dot > 0 && dot < len;
}

/* conflict of star triangle (p, nb[i], nb[i+1]) with s */
static int tri_conflict(const int64_t *X, const int64_t *Y, int p, int q0, int q1, int s) {
    if (q1 == INF_ID) {                      /* ghost (p, q0, inf): outside is LEFT of p->q0 */
        int o = orient_sign(X, Y, p, q0, s);
        if (o > 0) return 1;
        if (o < 0) return 0;
        return strictly_between(X, Y, p, q0, s);
    }
    if (q0 == INF_ID) {                      /* ghost (p, inf, q1): outside is RIGHT of p->q1 */
        int o = orient_sign(X, Y, p, q1, s);
        if (o < 0) return 1;
        if (o > 0) return 0;
        return strictly_between(X, Y, p, q1, s);
    }
    return incircle_sign(X, Y, p, q0, q1, s) > 0;
}

/* Build the star of p over all non-duplicate points. nb: out cyclic CCW list. returns degree
 * (entries incl. INF marker), or -1 on overflow, 0 if no triangle exists (all collinear). */
static int build_star(const int64_t *X, const int64_t *Y, const uint8_t *dup, int n, int p, int *nb) {
    int d = 0;
    int qpos = -1, qneg = -1;                /* collinear bootstrap: nearest on each side of p */
    for (int s = 0; s < n; ++s) {
        if (s == p || dup[s]) continue;
        if (d == 0) {
            if (qpos < 0) { qpos = s; continue; }
            int o = orient_sign(X, Y, p, qpos, s);
            if (o == 0) {
                i128 dx = X[qpos] - X[p], dy = Y[qpos] - Y[p];
                i128 sx = X[s] - X[p], sy = Y[s] - Y[p];
                i128 dot = dx * sx + dy * sy;
                if (dot > 0) {                /* same side as qpos: keep the nearer */
                    if (sx * sx + sy * sy < dx * dx + dy * dy) qpos = s;
                } else {
                    if (qneg < 0) qneg = s;
                    else {
                        i128 nx = X[qneg] - X[p], ny = Y[qneg] - Y[p];
                        if (sx * sx + sy * sy < nx * nx + ny * ny) qneg = s;
                    }
                }
                continue;
            }
            if (o > 0) {                      /* s left of p->qpos */
                nb[d++] = qpos; nb[d++] = s; if (qneg >= 0) nb[d++] = qneg; nb[d++] = INF_ID;
            } else {
                if (qneg >= 0) nb[d++] = qneg;
                nb[d++] = s; nb[d++] = qpos; nb[d++] = INF_ID;
            }
            continue;
        }
        /* conflict flags */
        int cf[MAX_DEG]; int any = 0, all = 1;
        for (int i = 0; i < d; ++i) {
            cf[i] = tri_conflict(X, Y, p, nb[i], nb[(i + 1) % d], s);
            any |= cf[i]; all &= cf[i];
        }
        if (!any) continue;
        if (all) return -2;                   /* cannot happen once a real triangle exists */
        /* arc start: first conflicting triangle whose predecessor is not conflicting */
        int i0 = 0;
        while (!(cf[i0] && !cf[(i0 + d - 1) % d])) ++i0;
        int len = 0;
        while (cf[(i0 + len) % d]) ++len;
        /* triangles i0..i0+len-1 removed => neighbours i0+1..i0+len-1 removed, s inserted after nb[i0] */
        int tmp[MAX_DEG]; int m = 0;
        for (int k = 0; k < d; ++k) {
            int idx = (i0 + 1 + k) % d;       /* walk starting just after nb[i0] */
            if (k < len - 1) continue;        /* removed interior neighbours */
            tmp[m++] = nb[idx];
        }
        /* tmp now = nb[i0+len], ..., nb[i0] (cyclic); new star = s followed by tmp */
        if (m + 1 > MAX_DEG) return -1;
        nb[0] = s;
        for (int k = 0; k < m; ++k) nb[k + 1] = tmp[k];
        d = m + 1;
    }
    return d;
}

static void snap(const float *x, int64_t *out, int n) {
    for (int i = 0; i < n; ++i) out[i] = (int64_t)ldexp((double)x[i], 40);
}

/* Exact Delaunay of n float32 points.  tri_out: capacity 2n triples, CANONICAL order
 * (each row ascending, rows lexsorted).  dup_out[i]=1 for dropped duplicates (may be NULL).
 * Returns the triangle count, or <0 on error (-1 degree overflow, -3 input out of range). */
int oracle_delaunay(const float *x, const float *y, int n, int32_t *tri_out, uint8_t *dup_out) {
    int64_t *X = (int64_t *)malloc(sizeof(int64_t) * (size_t)(n > 0 ? n : 1));
    int64_t *Y = (int64_t *)malloc(sizeof(int64_t) * (size_t)(n > 0 ? n : 1));
    uint8_t *dup = (uint8_t *)calloc((size_t)(n > 0 ? n : 1), 1);
    int T = 0, rc = 0;
    for (int i = 0; i < n; ++i) {
        if (!(fabsf(x[i]) < 4096.0f) || !(fabsf(y[i]) < 4096.0f)) { rc = -3; goto done; }
        if ((x[i] != 0 && fabsf(x[i]) < 7.62939453125e-06f) || (y[i] != 0 && fabsf(y[i]) < 7.62939453125e-06f)) { rc = -3; goto done; }
    }
    snap(x, X, n); snap(y, Y, n);
    for (int i = 0; i < n; ++i)
        for (int j = 0; j < i; ++j)
            if (X[i] == X[j] && Y[i] == Y[j]) { dup[i] = 1; break; }
    for (int p = 0; p < n; ++p) {
        if (dup[p]) continue;
        int nb[MAX_DEG];
        int d = build_star(X, Y, dup, n, p, nb);
        if (d < 0) { rc = -1; goto done; }
        /* emit finite triangles whose smallest vertex is p, sorted by (v1,v2) */
        int32_t loc[MAX_DEG][2]; int m = 0;
        for (int i = 0; i < d; ++i) {
            int q0 = nb[i], q1 = nb[(i + 1) % d];
            if (q0 == INF_ID || q1 == INF_ID) continue;
            if (q0 < p || q1 < p) continue;
            int a = q0 < q1 ? q0 : q1, b = q0 < q1 ? q1 : q0;
            int k = m++;
            while (k > 0 && (loc[k - 1][0] > a || (loc[k - 1][0] == a && loc[k - 1][1] > b))) {
                loc[k][0] = loc[k - 1][0]; loc[k][1] = loc[k - 1][1]; --k;
            }
            loc[k][0] = a; loc[k][1] = b;
        }
        for (int k = 0; k < m; ++k) {
            tri_out[3 * T + 0] = p; tri_out[3 * T + 1] = loc[k][0]; tri_out[3 * T + 2] = loc[k][1];
            ++T;
        }
    }
    if (dup_out) memcpy(dup_out, dup, (size_t)n);
    rc = T;
done:
    free(X); free(Y); free(dup);
    return rc;
}

/* Exact predicates exported for the Python validator (tests/): signs only. */
int oracle_orient(const float *x, const float *y, int a, int b, int c) {
    int64_t X[3], Y[3];
    float xs[3] = { x[a], x[b], x[c] }, ys[3] = { y[a], y[b], y[c] };
    snap(xs, X, 3); snap(ys, Y, 3);
    return orient_sign(X, Y, 0, 1, 2);
}

/* exact in-circle sign WITHOUT the tie-break (0 = co-circular) */
int oracle_incircle_raw(const float *x, const float *y, int a, int b, int c, int d) {
    int64_t X[4], Y[4];
    float xs[4] = { x[a], x[b], x[c], x[d] }, ys[4] = { y[a], y[b], y[c], y[d] };
    snap(xs, X, 4); snap(ys, Y, 4);
    i128 adx = X[0] - X[3], ady = Y[0] - Y[3];
    i128 bdx = X[1] - X[3], bdy = Y[1] - Y[3];
    i128 cdx = X[2] - X[3], cdy = Y[2] - Y[3];
    i128 al = adx * adx + ady * ady, bl = bdx * bdx + bdy * bdy, cl = cdx * cdx + cdy * cdy;
    s256 det = s256_add(s256_add(s256_from_mul(al, bdx * cdy - cdx * bdy), s256_from_mul(bl, cdx * ady - adx * cdy)),
                        s256_from_mul(cl, adx * bdy - bdx * ady));
    (void)s256_neg;
    return s256_sign(det);
}

/* in-circle sign WITH the tie-break; indices order the perturbation */
int oracle_incircle_sos(const float *x, const float *y, int n, int a, int b, int c, int d) {
    int64_t *X = (int64_t *)malloc(sizeof(int64_t) * (size_t)n), *Y = (int64_t *)malloc(sizeof(int64_t) * (size_t)n);
    snap(x, X, n); snap(y, Y, n);
    int s = incircle_sign(X, Y, a, b, c, d);
    free(X); free(Y);
    return s;
}
