// Stage 1: per-correspondence linear triangulation (DLT), solved in registers.
//
// Replaces the triangulation inside cv2.recoverPose(E, px_cur, px_ref, K, distanceThresh=100)
// and the dehomogenisation at reference src/thirdparty/MonocularVO/visual_odometry.py:129-147,
// plus the reprojection of src/main.py:102-104 (fx on BOTH axes).
//
// Per point: normalise both pixels by K in float64, build the 4x4 system
//   A = [ x0*P0[2]-P0[0] ; y0*P0[2]-P0[1] ; x1*P1[2]-P1[0] ; y1*P1[2]-P1[1] ],  P0=[I|0], P1=[R|t],
// take the right singular vector of the smallest singular value (inverse iteration on A^T A, all in
// registers), dehomogenise, and apply recoverPose's mask:
//   Z*W > 0,  Z/W < dist,  0 < (P1 X)_z < dist.
#pragma once
#include <stdint.h>

namespace mvosr {

struct Pose { double R[9]; double t[3]; };

// Returns the mask; X (float64, current-camera frame) and the reprojected pixel are written always.
//
// The right singular vector of A for its smallest singular value is the eigenvector of M = A^T A for the smallest
// eigenvalue.  It is found by inverse iteration on M + mu I (Cholesky factor in registers, three solves): the
// iteration converges by (lambda4 + mu) / (lambda3 + mu) per step, and the eigenvector of the COMPUTED M differs from
// the exact one by ~ eps |M| / (lambda3 - lambda4) -- the gap, not lambda4, sets the accuracy -- so forming A^T A
// loses nothing that matters here (checked against cv2.recoverPose's output to the last float32 bit in the tests).
__device__ __forceinline__ bool triangulate_point(float cu, float cv, float ru, float rv, const Pose &pose,
                                                  double fx, double fy, double cx, double cy, double dist,
                                                  double &X, double &Y, double &Z, double &u, double &v) {
    const double x0 = ((double)cu - cx) / fx, y0 = ((double)cv - cy) / fy;
    const double x1 = ((double)ru - cx) / fx, y1 = ((double)rv - cy) / fy;
    const double *R = pose.R, *t = pose.t;
    // rows 2,3 of A: x1*P1[2]-P1[0], y1*P1[2]-P1[1] with P1 = [R|t]; rows 0,1: (-1,0,x0,0), (0,-1,y0,0)
    const double a0 = x1 * R[6] - R[0], a1 = x1 * R[7] - R[1], a2 = x1 * R[8] - R[2], a3 = x1 * t[2] - t[0];
    const double b0 = y1 * R[6] - R[3], b1 = y1 * R[7] - R[4], b2 = y1 * R[8] - R[5], b3 = y1 * t[2] - t[1];
    double m00 = 1.0 + a0 * a0 + b0 * b0, m10 = a1 * a0 + b1 * b0, m11 = 1.0 + a1 * a1 + b1 * b1;
    double m20 = -x0 + a2 * a0 + b2 * b0, m21 = -y0 + a2 * a1 + b2 * b1, m22 = x0 * x0 + y0 * y0 + a2 * a2 + b2 * b2;
    double m30 = a3 * a0 + b3 * b0, m31 = a3 * a1 + b3 * b1, m32 = a3 * a2 + b3 * b2, m33 = a3 * a3 + b3 * b3;
    const double mu = 1.0e-13 * (m00 + m11 + m22 + m33);
    m00 += mu; m11 += mu; m22 += mu; m33 += mu;
    // Cholesky M = L L^T, reciprocals of the diagonal kept
    const double i0 = rsqrt(m00);
    const double l10 = m10 * i0, l20 = m20 * i0, l30 = m30 * i0;
    const double i1 = rsqrt(m11 - l10 * l10);
    const double l21 = (m21 - l20 * l10) * i1, l31 = (m31 - l30 * l10) * i1;
    const double i2 = rsqrt(m22 - l20 * l20 - l21 * l21);
    const double l32 = (m32 - l30 * l20 - l31 * l21) * i2;
    const double i3 = rsqrt(m33 - l30 * l30 - l31 * l31 - l32 * l32);
    double q0 = 0.5, q1 = 0.5, q2 = 0.5, q3 = 0.5;
#pragma unroll
    for (int it = 0; it < 3; ++it) {
        // L y = q
        const double y0_ = q0 * i0;
        const double y1_ = (q1 - l10 * y0_) * i1;
        const double y2_ = (q2 - l20 * y0_ - l21 * y1_) * i2;
        const double y3_ = (q3 - l30 * y0_ - l31 * y1_ - l32 * y2_) * i3;
        // L^T q = y
        q3 = y3_ * i3;
        q2 = (y2_ - l32 * q3) * i2;
        q1 = (y1_ - l21 * q2 - l31 * q3) * i1;
        q0 = (y0_ - l10 * q1 - l20 * q2 - l30 * q3) * i0;
        const double s = rsqrt(q0 * q0 + q1 * q1 + q2 * q2 + q3 * q3);
        q0 *= s; q1 *= s; q2 *= s; q3 *= s;
    }
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
