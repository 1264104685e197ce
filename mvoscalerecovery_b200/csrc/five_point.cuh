// Five-point essential matrix + RANSAC (first half): the algorithm behind
//   cv2.findEssentialMat(px_cur, px_ref, cameraMatrix=K, method=cv2.RANSAC, prob=0.999, threshold=0.5)
// as the reference calls it (src/thirdparty/MonocularVO/visual_odometry.py:100-102,129-130).  The algorithm lives in OpenCV
// (calib3d five-point.cpp + ptsetreg.cpp), not in the reference tree; OpenCV draws its samples from its own RNG, so parity is
// defined on a Philox stream of our own (below) and pinned on the pose / inlier set.
//
// The minimal solver is Stewenius' action-matrix form of Nister's problem, with numerics one thread can run in its own
// registers / local memory (no LAPACK): Gauss-Jordan null space with full pivoting, the ten cubic constraints by
// interpolation at 20 fixed nodes, Gauss-Jordan on the 10x20 coefficient matrix, the real eigenvalues of the 10x10 action
// matrix by balancing + Hessenberg reduction + double-shift QR (accurate as they come: a Rayleigh-quotient polish changed
// nothing measurable and was dropped), a 6x5 least-squares back-substitution per root, and a final test that the result IS an essential matrix.
//
// Everything numerical is __host__ __device__ so that tests/host_sim can compile the very same functions for the host and
// check them against the oracle without a GPU; the library itself only ever calls them from find_essential_kernel.
#pragma once
#include <stdint.h>
#include <math.h>

#ifdef __CUDACC__
#define MVOSR_FP5_HD __host__ __device__ inline
#else
#define MVOSR_FP5_HD inline
#endif

namespace mvosr {
namespace fp5 {

struct Tables { double pts[20][3]; double vinv[20][20]; };

// ---------------------------------------------------------------------------------------------------------------------------
// The sample stream: five distinct positions of range(n) for hypothesis `hyp` of frame `frame` of sequence `seq`.
//   key = (seed lo, seed hi); (r0, r1, r2, r3) = Philox4x32-10(counter = (hyp, frame, seq, 1)); r4 = word 0 of counter (.., 2)
//   p_k = (r_k * (n - k)) >> 32, then + 1 for every earlier position (taken in ascending order) it is >= to.
// (Counter word 3 = 0 is the plane RANSAC's stream, philox.cuh.)
MVOSR_FP5_HD void philox(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t k0, uint32_t k1, uint32_t out[4]) {
    for (int r = 0; r < 10; ++r) {
        const uint64_t p0 = (uint64_t)0xD2511F53u * c0, p1 = (uint64_t)0xCD9E8D57u * c2;
        const uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ k0, n2 = (uint32_t)(p0 >> 32) ^ c3 ^ k1;
        c0 = n0; c1 = (uint32_t)p1; c2 = n2; c3 = (uint32_t)p0;
        k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
    }
    out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}

MVOSR_FP5_HD void sample5(uint64_t seed, uint32_t hyp, uint32_t frame, uint32_t seq, uint32_t n, int idx[5]) {
    uint32_t r[8];
    philox(hyp, frame, seq, 1u, (uint32_t)seed, (uint32_t)(seed >> 32), r);
    philox(hyp, frame, seq, 2u, (uint32_t)seed, (uint32_t)(seed >> 32), r + 4);
    int sorted[5];
    for (int k = 0; k < 5; ++k) {
        int p = (int)(((uint64_t)r[k] * (uint64_t)(n - (uint32_t)k)) >> 32);
        for (int j = 0; j < k; ++j) if (p >= sorted[j]) ++p;
        idx[k] = p;
        int j = k;                                                  // insert into the ascending list
        while (j > 0 && sorted[j - 1] > p) { sorted[j] = sorted[j - 1]; --j; }
        sorted[j] = p;
    }
}

// ---------------------------------------------------------------------------------------------------------------------------
// det E and the nine entries of 2 E E^T E - tr(E E^T) E
MVOSR_FP5_HD void constraints_at(const double E[9], double out[10]) {
    out[0] = E[0] * (E[4] * E[8] - E[5] * E[7]) - E[1] * (E[3] * E[8] - E[5] * E[6]) + E[2] * (E[3] * E[7] - E[4] * E[6]);
    double G[9];                                                    // E E^T
    for (int r = 0; r < 3; ++r)
        for (int c = 0; c < 3; ++c) G[3 * r + c] = E[3 * r] * E[3 * c] + E[3 * r + 1] * E[3 * c + 1] + E[3 * r + 2] * E[3 * c + 2];
    const double tr = G[0] + G[4] + G[8];
    for (int r = 0; r < 3; ++r)
        for (int c = 0; c < 3; ++c)
            out[1 + 3 * r + c] = 2.0 * (G[3 * r] * E[c] + G[3 * r + 1] * E[3 + c] + G[3 * r + 2] * E[6 + c]) - tr * E[3 * r + c];
}

// Four orthonormal null vectors of the 5x9 epipolar system (rows kron(x2h, x1h)); false when the system is rank deficient.
MVOSR_FP5_HD bool null4(const double x1[10], const double x2[10], double basis[4][9]) {
    double M[5][9];
    int cols[9];
    for (int i = 0; i < 5; ++i) {
        const double a[3] = { x2[2 * i], x2[2 * i + 1], 1.0 }, b[3] = { x1[2 * i], x1[2 * i + 1], 1.0 };
        for (int p = 0; p < 3; ++p) for (int q = 0; q < 3; ++q) M[i][3 * p + q] = a[p] * b[q];
    }
    for (int c = 0; c < 9; ++c) cols[c] = c;
    #pragma unroll 1
    for (int r = 0; r < 5; ++r) {
        int pi = r, pj = r;
        double best = -1.0;
        for (int i = r; i < 5; ++i)
            for (int j = r; j < 9; ++j) { const double v = fabs(M[i][j]); if (v > best) { best = v; pi = i; pj = j; } }
        if (!(best > 1e-300)) return false;
        if (pi != r) for (int j = 0; j < 9; ++j) { const double t = M[r][j]; M[r][j] = M[pi][j]; M[pi][j] = t; }
        if (pj != r) {
            for (int i = 0; i < 5; ++i) { const double t = M[i][r]; M[i][r] = M[i][pj]; M[i][pj] = t; }
            const int t = cols[r]; cols[r] = cols[pj]; cols[pj] = t;
        }
        const double piv = M[r][r];
        for (int j = 0; j < 9; ++j) M[r][j] /= piv;
        for (int k = 0; k < 5; ++k) {
            if (k == r) continue;
            const double f = M[k][r];
            for (int j = 0; j < 9; ++j) M[k][j] -= f * M[r][j];
        }
    }
    for (int f = 0; f < 4; ++f) {
        for (int c = 0; c < 9; ++c) basis[f][c] = 0.0;
        basis[f][cols[5 + f]] = 1.0;
        for (int c = 0; c < 5; ++c) basis[f][cols[c]] = -M[c][5 + f];
    }
    for (int f = 0; f < 4; ++f) {                                   // modified Gram-Schmidt
        for (int g = 0; g < f; ++g) {
            double d = 0.0;
            for (int c = 0; c < 9; ++c) d += basis[f][c] * basis[g][c];
            for (int c = 0; c < 9; ++c) basis[f][c] -= d * basis[g][c];
        }
        double n2 = 0.0;
        for (int c = 0; c < 9; ++c) n2 += basis[f][c] * basis[f][c];
        const double nrm = sqrt(n2);
        if (!(nrm > 0.0)) return false;
        for (int c = 0; c < 9; ++c) basis[f][c] /= nrm;
    }
    return true;
}

// The 6 non-trivial rows of the action matrix of "multiply by x" on [x^2, xy, xz, y^2, yz, z^2, x, y, z, 1]
// (rows 6..9 are the unit vectors e0, e1, e2, e6).  false when the elimination meets a vanishing pivot.
MVOSR_FP5_HD bool action_rows(const double basis[4][9], const Tables &T, double A6[6][10]) {
    double M[10][20];
    {
        double vals[20][10];
        #pragma unroll 1
        for (int s = 0; s < 20; ++s) {
            double E[9];
            for (int c = 0; c < 9; ++c) E[c] = T.pts[s][0] * basis[0][c] + T.pts[s][1] * basis[1][c] + T.pts[s][2] * basis[2][c] + basis[3][c];
            constraints_at(E, vals[s]);
        }
        #pragma unroll 1
        for (int c = 0; c < 10; ++c)
#pragma unroll 1
            for (int m = 0; m < 20; ++m) {
                double a = 0.0;
                for (int s = 0; s < 20; ++s) a += T.vinv[m][s] * vals[s][c];
                M[c][m] = a;
            }
    }
    #pragma unroll 1
    for (int c = 0; c < 10; ++c) {
        double mx = 0.0;
        for (int m = 0; m < 20; ++m) { const double v = fabs(M[c][m]); if (v > mx) mx = v; }
        if (!(mx > 0.0) || !(mx < 1e300)) return false;
        for (int m = 0; m < 20; ++m) M[c][m] /= mx;
    }
    #pragma unroll 1
    for (int col = 0; col < 10; ++col) {
        int piv = col;
        double best = fabs(M[col][col]);
        for (int r = col + 1; r < 10; ++r) { const double v = fabs(M[r][col]); if (v > best) { best = v; piv = r; } }
        if (!(best >= 1e-13)) return false;
        if (piv != col) for (int m = 0; m < 20; ++m) { const double t = M[col][m]; M[col][m] = M[piv][m]; M[piv][m] = t; }
        const double p = M[col][col];
        for (int m = 0; m < 20; ++m) M[col][m] /= p;
        #pragma unroll 1
        for (int r = 0; r < 10; ++r) {
            if (r == col) continue;
            const double f = M[r][col];
            for (int m = 0; m < 20; ++m) M[r][m] -= f * M[col][m];
        }
    }
    for (int r = 0; r < 6; ++r) for (int c = 0; c < 10; ++c) A6[r][c] = -M[r][10 + c];
    return true;
}

// Real eigenvalues of the 10x10 action matrix: balancing, reduction to Hessenberg form by stabilised elimination, and the
// Francis double-shift QR iteration (the textbook EISPACK balanc / elmhes / hqr sequence, which is also what LAPACK's
// general eigenvalue driver does) -- the well-conditioned route; the coefficients of the characteristic polynomial lose
// the small roots when the eigenvalues spread over several orders of magnitude.  Every loop is bounded.
// Returns the number of real eigenvalues (|imaginary part| <= 1e-9 max(1, |real part|), one per close pair), ascending;
// -1 when the iteration does not converge (the hypothesis is then dropped).
MVOSR_FP5_HD int real_eigenvalues(const double A6[6][10], double roots[10]) {
    const int n = 10;
    double a[10][10];
    for (int i = 0; i < n; ++i) for (int j = 0; j < n; ++j) a[i][j] = i < 6 ? A6[i][j] : 0.0;
    a[6][0] = 1.0; a[7][1] = 1.0; a[8][2] = 1.0; a[9][6] = 1.0;
    // -- balancing: similarity by powers of two until row and column norms are within a factor of two
    for (int pass = 0, again = 1; again && pass < 24; ++pass) {
        again = 0;
        #pragma unroll 1
        for (int i = 0; i < n; ++i) {
            double r = 0.0, c = 0.0;
            for (int j = 0; j < n; ++j) if (j != i) { c += fabs(a[j][i]); r += fabs(a[i][j]); }
            if (!(c > 0.0) || !(r > 0.0) || !(c < 1e300) || !(r < 1e300)) continue;
            double g = r / 2.0, f = 1.0;
            const double s = c + r;
            for (int it = 0; c < g && it < 600; ++it) { f *= 2.0; c *= 4.0; }
            g = r * 2.0;
            for (int it = 0; c > g && it < 600; ++it) { f /= 2.0; c /= 4.0; }
            if ((c + r) / f < 0.95 * s) {
                again = 1;
                g = 1.0 / f;
                for (int j = 0; j < n; ++j) a[i][j] *= g;
                for (int j = 0; j < n; ++j) a[j][i] *= f;
            }
        }
    }
    // -- Hessenberg form by elimination with pivoting
    #pragma unroll 1
    for (int m = 1; m < n - 1; ++m) {
        double x = 0.0;
        int piv = m;
        for (int j = m; j < n; ++j) if (fabs(a[j][m - 1]) > fabs(x)) { x = a[j][m - 1]; piv = j; }
        if (piv != m) {
            for (int j = m - 1; j < n; ++j) { const double t = a[piv][j]; a[piv][j] = a[m][j]; a[m][j] = t; }
            for (int j = 0; j < n; ++j) { const double t = a[j][piv]; a[j][piv] = a[j][m]; a[j][m] = t; }
        }
        if (x != 0.0)
            #pragma unroll 1
            for (int i = m + 1; i < n; ++i) {
                double y = a[i][m - 1];
                if (y == 0.0) continue;
                y /= x;
                a[i][m - 1] = 0.0;
                for (int j = m; j < n; ++j) a[i][j] -= y * a[m][j];
                for (int j = 0; j < n; ++j) a[j][m] += y * a[j][i];
            }
    }
    // -- double-shift QR on the Hessenberg matrix
    double wr[10], wi[10];
    double anorm = 0.0;
    for (int i = 0; i < n; ++i) for (int j = (i > 0 ? i - 1 : 0); j < n; ++j) anorm += fabs(a[i][j]);
    if (!(anorm < 1e300)) return -1;
    int nn = n - 1;
    double t = 0.0;
    while (nn >= 0) {
        int its = 0, l;
        do {
            for (l = nn; l >= 1; --l) {
                double s = fabs(a[l - 1][l - 1]) + fabs(a[l][l]);
                if (s == 0.0) s = anorm;
                if (fabs(a[l][l - 1]) + s == s) { a[l][l - 1] = 0.0; break; }
            }
            double x = a[nn][nn];
            if (l == nn) {                                          // one root
                wr[nn] = x + t; wi[nn] = 0.0; --nn;
            } else {
                double y = a[nn - 1][nn - 1], w = a[nn][nn - 1] * a[nn - 1][nn];
                if (l == nn - 1) {                                  // two roots
                    const double p = 0.5 * (y - x), q = p * p + w;
                    double z = sqrt(fabs(q));
                    x += t;
                    if (q >= 0.0) {
                        z = p + (p >= 0.0 ? fabs(z) : -fabs(z));
                        wr[nn - 1] = wr[nn] = x + z;
                        if (z != 0.0) wr[nn] = x - w / z;
                        wi[nn - 1] = wi[nn] = 0.0;
                    } else {
                        wr[nn - 1] = wr[nn] = x + p;
                        wi[nn] = z; wi[nn - 1] = -z;
                    }
                    nn -= 2;
                } else {                                            // no root yet: one more sweep
                    if (its == 60) return -1;
                    if (its == 10 || its == 20 || its == 40) {      // exceptional shift
                        t += x;
                        for (int i = 0; i <= nn; ++i) a[i][i] -= x;
                        const double s = fabs(a[nn][nn - 1]) + fabs(a[nn - 1][nn - 2]);
                        y = x = 0.75 * s;
                        w = -0.4375 * s * s;
                    }
                    ++its;
                    int m;
                    double p = 0.0, q = 0.0, r = 0.0, z;
                    for (m = nn - 2; m >= l; --m) {
                        z = a[m][m];
                        r = x - z;
                        double s = y - z;
                        p = (r * s - w) / a[m + 1][m] + a[m][m + 1];
                        q = a[m + 1][m + 1] - z - r - s;
                        r = a[m + 2][m + 1];
                        s = fabs(p) + fabs(q) + fabs(r);
                        p /= s; q /= s; r /= s;
                        if (m == l) break;
                        const double u = fabs(a[m][m - 1]) * (fabs(q) + fabs(r));
                        const double v = fabs(p) * (fabs(a[m - 1][m - 1]) + fabs(z) + fabs(a[m + 1][m + 1]));
                        if (u + v == v) break;
                    }
                    if (m < l) m = l;                               // only on non-finite data; keeps the indices in range
                    for (int i = m + 2; i <= nn; ++i) {
                        a[i][i - 2] = 0.0;
                        if (i != m + 2) a[i][i - 3] = 0.0;
                    }
                    for (int k = m; k <= nn - 1; ++k) {
                        if (k != m) {
                            p = a[k][k - 1];
                            q = a[k + 1][k - 1];
                            r = k != nn - 1 ? a[k + 2][k - 1] : 0.0;
                            x = fabs(p) + fabs(q) + fabs(r);
                            if (x != 0.0) { p /= x; q /= x; r /= x; }
                        }
                        const double nrm = sqrt(p * p + q * q + r * r);
                        const double s = p >= 0.0 ? nrm : -nrm;
                        if (s != 0.0) {
                            if (k == m) {
                                if (l != m) a[k][k - 1] = -a[k][k - 1];
                            } else {
                                a[k][k - 1] = -s * x;
                            }
                            p += s;
                            x = p / s; y = q / s; z = r / s;
                            q /= p; r /= p;
                            for (int j = k; j <= nn; ++j) {
                                p = a[k][j] + q * a[k + 1][j];
                                if (k != nn - 1) { p += r * a[k + 2][j]; a[k + 2][j] -= p * z; }
                                a[k + 1][j] -= p * y;
                                a[k][j] -= p * x;
                            }
                            const int mmin = nn < k + 3 ? nn : k + 3;
                            for (int i = l; i <= mmin; ++i) {
                                p = x * a[i][k] + y * a[i][k + 1];
                                if (k != nn - 1) { p += z * a[i][k + 2]; a[i][k + 2] -= p * r; }
                                a[i][k + 1] -= p * q;
                                a[i][k] -= p;
                            }
                        }
                    }
                }
            }
        } while (l < nn - 1);
    }
    int n_roots = 0;
    for (int i = 0; i < n; ++i) {
        if (!(fabs(wi[i]) <= 1e-9 * fmax(1.0, fabs(wr[i]))) || wi[i] < 0.0) continue;    // complex, or the second member of a close pair
        const double v = wr[i];
        int j = n_roots++;
        while (j > 0 && roots[j - 1] > v) { roots[j] = roots[j - 1]; --j; }
        roots[j] = v;
    }
    return n_roots;
}


// Solves S z = b in place (Gaussian elimination, partial pivoting); S is destroyed.  false on an exactly zero pivot.
template <int N>
MVOSR_FP5_HD bool solve_in_place(double S[N][N], double b[N]) {
    #pragma unroll 1
    for (int col = 0; col < N; ++col) {
        int piv = col;
        double best = fabs(S[col][col]);
        for (int r = col + 1; r < N; ++r) { const double v = fabs(S[r][col]); if (v > best) { best = v; piv = r; } }
        if (!(best > 0.0)) return false;
        if (piv != col) {
            for (int j = 0; j < N; ++j) { const double t = S[col][j]; S[col][j] = S[piv][j]; S[piv][j] = t; }
            const double t = b[col]; b[col] = b[piv]; b[piv] = t;
        }
        #pragma unroll 1
        for (int r = col + 1; r < N; ++r) {
            const double f = S[r][col] / S[col][col];
            if (f == 0.0) continue;
            for (int j = col; j < N; ++j) S[r][j] -= f * S[col][j];
            b[r] -= f * b[col];
        }
    }
    #pragma unroll 1
    for (int r = N - 1; r >= 0; --r) {
        double a = b[r];
        for (int j = r + 1; j < N; ++j) a -= S[r][j] * b[j];
        b[r] = a / S[r][r];
    }
    return true;
}

// All real essential matrices through five correspondences (normalised coordinates, x2^T E x1 = 0): up to ten 3x3 matrices
// of unit Frobenius norm, row-major, in ascending order of the eigenvalue they belong to.  Returns how many.
MVOSR_FP5_HD int solve(const double x1[10], const double x2[10], const Tables &T, double Eout[10][9]) {
    double basis[4][9], A6[6][10], roots[10];
    if (!null4(x1, x2, basis)) return 0;
    if (!action_rows(basis, T, A6)) return 0;
    const int nr = real_eigenvalues(A6, roots);
    int n_sol = 0;
    #pragma unroll 1
    for (int k = 0; k < nr; ++k) {
        const double x = roots[k];
        // (A - x I) v = 0, v = [x^2, xy, xz, y^2, yz, z^2, x, y, z, 1]: rows 0..5 are linear in (y^2, yz, z^2, y, z) once
        // xy = x*y and xz = x*z are substituted; 6x5 least squares through the normal equations
        double L[6][5], rhs[6];
        for (int r = 0; r < 6; ++r) {
            double R[10];
            for (int j = 0; j < 10; ++j) R[j] = A6[r][j] - (j == r ? x : 0.0);
            L[r][0] = R[3]; L[r][1] = R[4]; L[r][2] = R[5]; L[r][3] = R[7] + x * R[1]; L[r][4] = R[8] + x * R[2];
            rhs[r] = -(R[0] * x * x + R[6] * x + R[9]);
        }
        double N[5][5], g[5];
        for (int i = 0; i < 5; ++i) {
            for (int j = 0; j < 5; ++j) { double a = 0.0; for (int r = 0; r < 6; ++r) a += L[r][i] * L[r][j]; N[i][j] = a; }
            double a = 0.0;
            for (int r = 0; r < 6; ++r) a += L[r][i] * rhs[r];
            g[i] = a;
        }
        if (!solve_in_place<5>(N, g)) continue;
        const double y = g[3], z = g[4];
        double E[9], n2 = 0.0;
        for (int i = 0; i < 9; ++i) { E[i] = x * basis[0][i] + y * basis[1][i] + z * basis[2][i] + basis[3][i]; n2 += E[i] * E[i]; }
        const double nrm = sqrt(n2);
        for (int i = 0; i < 9; ++i) E[i] /= nrm;
        double cons[10];
        constraints_at(E, cons);
        bool ok = true;
        for (int i = 0; i < 10; ++i) if (!(fabs(cons[i]) < 1e-6)) ok = false;
        if (!ok) continue;
        for (int i = 0; i < 9; ++i) Eout[n_sol][i] = E[i];
        ++n_sol;
    }
    return n_sol;
}

// Is the squared Sampson distance of one correspondence (normalised coordinates) to the epipolar constraint of E below thr2?
// Written without the division: s^2 < thr2 * (|E x1|_xy^2 + |E^T x2|_xy^2); a vanishing denominator is never an inlier.
MVOSR_FP5_HD bool sampson_inlier(const double E[9], double ax, double ay, double bx, double by, double thr2) {
    const double e0 = E[0] * ax + E[1] * ay + E[2], e1 = E[3] * ax + E[4] * ay + E[5], e2 = E[6] * ax + E[7] * ay + E[8];
    const double f0 = bx * E[0] + by * E[3] + E[6], f1 = bx * E[1] + by * E[4] + E[7];
    const double s = bx * e0 + by * e1 + e2;
    return s * s < thr2 * (e0 * e0 + e1 * e1 + f0 * f0 + f1 * f1);
}

// The adaptive stopping rule, checked at the end of every round: tried >= log(1 - confidence) / log(1 - w^5), w = best / n.
// confidence <= 0 disables it; a perfect model (w = 1) or confidence >= 1 ... stops / never stops as the formula's limits say.
MVOSR_FP5_HD bool enough_hypotheses(int tried, int best_count, int n, double confidence) {
    if (!(confidence > 0.0) || best_count <= 0) return false;
    if (confidence >= 1.0) return false;
    const double w = (double)best_count / (double)n, w5 = w * w * w * w * w;
    if (w5 >= 1.0) return true;
    return (double)tried * log(1.0 - w5) <= log(1.0 - confidence);      // both logs negative
}

}  // namespace fp5
}  // namespace mvosr
