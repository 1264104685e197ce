// Stage 1: per-correspondence linear triangulation (DLT), solved in registers.
//
// Replaces the triangulation inside cv2.recoverPose(E, px_cur, px_ref, K, distanceThresh=100)
// and the dehomogenisation at reference src/thirdparty/MonocularVO/visual_odometry.py:129-147,
// plus the reprojection of src/main.py:102-104 (fx on BOTH axes).
//
// Per point: normalise both pixels by K in float64, build the 4x4 system
//   A = [ x0*P0[2]-P0[0] ; y0*P0[2]-P0[1] ; x1*P1[2]-P1[0] ; y1*P1[2]-P1[1] ],  P0=[I|0], P1=[R|t],
// take the right singular vector of the smallest singular value by a one-sided (Hestenes)
// Jacobi SVD -- the same family OpenCV's own cv::SVD uses -- entirely in registers (all pair
// indices are compile-time), dehomogenise, and apply recoverPose's mask:
//   Z*W > 0,  Z/W < dist,  0 < (P1 X)_z < dist.
#pragma once
#include <stdint.h>

namespace mvosr {

struct Pose { double R[9]; double t[3]; };

template <int P, int Q>
__device__ __forceinline__ bool jacobi_pair(double (&a)[4][4], double (&v)[4][4]) {
    // columns P and Q of a (a[row][col]); rotate so that they become orthogonal
    double alpha = 0, beta = 0, gamma = 0;
#pragma unroll
    for (int r = 0; r < 4; ++r) { alpha += a[r][P] * a[r][P]; beta += a[r][Q] * a[r][Q]; gamma += a[r][P] * a[r][Q]; }
    if (fabs(gamma) <= 1.0e-17 * sqrt(alpha * beta) || gamma == 0.0) return false;
    double zeta = (beta - alpha) / (2.0 * gamma);
    double tt = (zeta >= 0 ? 1.0 : -1.0) / (fabs(zeta) + sqrt(1.0 + zeta * zeta));
    double c = rsqrt(1.0 + tt * tt), s = c * tt;
#pragma unroll
    for (int r = 0; r < 4; ++r) {
        double ap = a[r][P], aq = a[r][Q];
        a[r][P] = c * ap - s * aq; a[r][Q] = s * ap + c * aq;
        double vp = v[r][P], vq = v[r][Q];
        v[r][P] = c * vp - s * vq; v[r][Q] = s * vp + c * vq;
    }
    return true;
}

// Returns the mask; X (float64, current-camera frame) and the reprojected pixel are written always.
__device__ __noinline__ bool triangulate_point(float cu, float cv, float ru, float rv, const Pose &pose,
                                                  double fx, double fy, double cx, double cy, double dist,
                                                  double &X, double &Y, double &Z, double &u, double &v) {
    double x0 = ((double)cu - cx) / fx, y0 = ((double)cv - cy) / fy;
    double x1 = ((double)ru - cx) / fx, y1 = ((double)rv - cy) / fy;
    const double *R = pose.R, *t = pose.t;
    double a[4][4], V[4][4];
    a[0][0] = -1.0; a[0][1] = 0.0;  a[0][2] = x0; a[0][3] = 0.0;
    a[1][0] = 0.0;  a[1][1] = -1.0; a[1][2] = y0; a[1][3] = 0.0;
    a[2][0] = x1 * R[6] - R[0]; a[2][1] = x1 * R[7] - R[1]; a[2][2] = x1 * R[8] - R[2]; a[2][3] = x1 * t[2] - t[0];
    a[3][0] = y1 * R[6] - R[3]; a[3][1] = y1 * R[7] - R[4]; a[3][2] = y1 * R[8] - R[5]; a[3][3] = y1 * t[2] - t[1];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) V[i][j] = (i == j) ? 1.0 : 0.0;
    for (int sweep = 0; sweep < 12; ++sweep) {
        bool any = false;
        any |= jacobi_pair<0, 1>(a, V); any |= jacobi_pair<0, 2>(a, V); any |= jacobi_pair<0, 3>(a, V);
        any |= jacobi_pair<1, 2>(a, V); any |= jacobi_pair<1, 3>(a, V); any |= jacobi_pair<2, 3>(a, V);
        if (!any) break;
    }
    double nrm[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) nrm[j] = a[0][j] * a[0][j] + a[1][j] * a[1][j] + a[2][j] * a[2][j] + a[3][j] * a[3][j];
    double q0 = V[0][0], q1 = V[1][0], q2 = V[2][0], q3 = V[3][0], best = nrm[0];
#pragma unroll
    for (int j = 1; j < 4; ++j)
        if (nrm[j] < best) { best = nrm[j]; q0 = V[0][j]; q1 = V[1][j]; q2 = V[2][j]; q3 = V[3][j]; }
    bool m = (q2 * q3) > 0.0;
    X = q0 / q3; Y = q1 / q3; Z = q2 / q3;
    m = m && (Z < dist);
    double z1 = X * R[6] + Y * R[7] + Z * R[8] + t[2];
    m = m && (z1 > 0.0) && (z1 < dist);
    u = X * fx / Z + cx;
    v = Y * fx / Z + cy;          // fx on both axes: src/main.py:103-104
    return m;
}

}  // namespace mvosr
