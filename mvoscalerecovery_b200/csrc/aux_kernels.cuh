// Stand-alone kernels for callers that bring their own triangles / point lists / motions (device code, sm_100a).
//
// Reference calls replaced (file:line into the reference tree):
//   triangle_planes_kernel   the per-triangle body of flat_selection (src/rescale.py:77-84), Reconstruct.triangle_model
//                            (src/reconstruct.py:70-90), feature_selection_by_tri (src/scale_calculator.py:228-238) and of
//                            the batch scripts (src/calculate_height_pitch.py:76-93, src/triangle_batch.py:31-44)
//   ransac_planes_kernel     get_pitch_ransac / run_ransac (src/estimate_road_norm.py:66-70,
//                            src/thirdparty/Ransac/ransac.py:3-23) on explicit point lists
//   integrate_paths_kernel   get_path / motion2pose (src/main_offline.py:95-119)
//   triangle_votes_kernel    find_outliers / check_triangle (src/rescale.py:45-72, src/scale_calculator.py:105-119,151-167)
//   recover_pose_kernel      the pose selection of cv2.recoverPose(E, px_cur, px_ref, K, distanceThresh=100)
//                            (src/thirdparty/MonocularVO/visual_odometry.py:129-133): decomposeEssentialMat + cheirality count
//   pose_mask_kernel         recoverPose's per-correspondence mask under the chosen pose (visual_odometry.py:134-136)
//   raster_mesh_kernel + mesh_depth_kernel   Reconstruct.depth_generate (src/reconstruct.py:91-107): tri.find_simplex of every
//                            pixel + the depth of the triangle's plane along the pixel's ray
// All of them are HBM-bound gathers / streams; the arithmetic is float64 because the callers hand float64 arrays.
#pragma once
#include <stdint.h>
#include <math_constants.h>
#include "../../include/mvosr.h"
#include "philox.cuh"
#include "triangulate.cuh"

namespace mvosr {

// n = P^-1 . 1 for the 3x3 matrix P whose rows are the triangle's vertices, in closed form:
// n = (e1 x e2) / (p0 . (e1 x e2)).  One thread per triangle, grid-stride.
__global__ void __launch_bounds__(256) triangle_planes_kernel(int n_tri, const int32_t *__restrict__ tri, const double *__restrict__ xyz,
                                                              double *normal, double *height, double *mean_y) {
    for (int t = blockIdx.x * blockDim.x + threadIdx.x; t < n_tri; t += gridDim.x * blockDim.x) {
        const int i0 = tri[3 * t], i1 = tri[3 * t + 1], i2 = tri[3 * t + 2];
        const double p0x = xyz[3 * i0], p0y = xyz[3 * i0 + 1], p0z = xyz[3 * i0 + 2];
        const double p1y = xyz[3 * i1 + 1], p2y = xyz[3 * i2 + 1];
        const double e1x = xyz[3 * i1] - p0x, e1y = p1y - p0y, e1z = xyz[3 * i1 + 2] - p0z;
        const double e2x = xyz[3 * i2] - p0x, e2y = p2y - p0y, e2z = xyz[3 * i2 + 2] - p0z;
        const double cx = e1y * e2z - e1z * e2y, cy = e1z * e2x - e1x * e2z, cz = e1x * e2y - e1y * e2x;
        const double det = p0x * cx + p0y * cy + p0z * cz;
        const double inv = det != 0.0 ? 1.0 / det : CUDART_NAN;
        if (normal) { normal[3 * t] = cx * inv; normal[3 * t + 1] = cy * inv; normal[3 * t + 2] = cz * inv; }
        if (height) height[t] = det != 0.0 ? fabs(det) / sqrt(cx * cx + cy * cy + cz * cz) : CUDART_NAN;
        if (mean_y) mean_y[t] = (p0y + p1y + p2y) / 3.0;
    }
}

// Depth-order votes of caller-supplied triangles: per vertex the number of triangles that flag it
// (a = (v0-v1)(d0-d1) > 0, b = (v0-v2)(d0-d2) > 0, c = (v1-v2)(d1-d2) > 0; flags [a|b, a|b|c, c]) and the number of
// triangles it belongs to.  v = pixel row, d = depth.
__global__ void __launch_bounds__(256) triangle_votes_kernel(int n_tri, const int32_t *__restrict__ tri, const double *__restrict__ v,
                                                             const double *__restrict__ d, int32_t *flagged, int32_t *incident) {
    for (int t = blockIdx.x * blockDim.x + threadIdx.x; t < n_tri; t += gridDim.x * blockDim.x) {
        const int i0 = tri[3 * t], i1 = tri[3 * t + 1], i2 = tri[3 * t + 2];
        const double v0 = v[i0], v1 = v[i1], v2 = v[i2], d0 = d[i0], d1 = d[i1], d2 = d[i2];
        const bool a = (v0 - v1) * (d0 - d1) > 0, b = (v0 - v2) * (d0 - d2) > 0, c = (v1 - v2) * (d1 - d2) > 0;
        if (a || b) atomicAdd(flagged + i0, 1);
        if (a || b || c) atomicAdd(flagged + i1, 1);
        if (c) atomicAdd(flagged + i2, 1);
        if (incident) { atomicAdd(incident + i0, 1); atomicAdd(incident + i1, 1); atomicAdd(incident + i2, 1); }
    }
}

// Plane RANSAC over explicit point lists: one CTA per list, one warp per hypothesis, RW hypotheses per round; after each
// round every thread replays run_ransac's sequential bookkeeping (keep the first strictly larger count, stop at the first
// count above the goal), so the result is the one the sequential loop returns.  Model = closed-form null vector of [p 1].
constexpr int RW = 8;                    // warps per CTA
__global__ void __launch_bounds__(32 * RW) ransac_planes_kernel(int n_sets, const int32_t *__restrict__ offsets, const double *__restrict__ xyz,
        int iterations, double threshold, double goal_fraction, int stop_at_goal, uint64_t seed, const int32_t *frame_index, int seq_id,
        double *model, int32_t *ic_out, int32_t *best_hyp_out, int32_t *hyps_used_out) {
    __shared__ int r_ic[RW];
    __shared__ double r_m[RW][5];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (int s = blockIdx.x; s < n_sets; s += gridDim.x) {
        const int base = offsets[s], n = offsets[s + 1] - base;
        const double *P = xyz + 3 * (size_t)base;
        const uint32_t frame = (uint32_t)(frame_index ? frame_index[s] : s);
        const double goal = (double)n * goal_fraction;
        int h_done = 0, best = -1, best_ic = 0, used = 0;
        bool stop = false;
        double b0 = 0, b1 = 0, b2 = 0, b3 = 0, b4 = 1;
        while (n >= 3 && h_done < iterations && !stop) {
            const int h = h_done + warp;
            if (h < iterations) {
                uint32_t i0, i1, i2;
                sample3_positions(seed, (uint32_t)h, frame, (uint32_t)seq_id, (uint32_t)n, i0, i1, i2);
                const double p0x = P[3 * i0], p0y = P[3 * i0 + 1], p0z = P[3 * i0 + 2];
                const double e1x = P[3 * i1] - p0x, e1y = P[3 * i1 + 1] - p0y, e1z = P[3 * i1 + 2] - p0z;
                const double e2x = P[3 * i2] - p0x, e2y = P[3 * i2 + 1] - p0y, e2z = P[3 * i2 + 2] - p0z;
                // products rounded separately (no FMA contraction): the cross product of two equal edges is exactly zero
                const double nx = __dsub_rn(__dmul_rn(e1y, e2z), __dmul_rn(e1z, e2y)), ny = __dsub_rn(__dmul_rn(e1z, e2x), __dmul_rn(e1x, e2z)),
                             nz = __dsub_rn(__dmul_rn(e1x, e2y), __dmul_rn(e1y, e2x));
                const double dd = -(nx * p0x + ny * p0y + nz * p0z);
                const double n4 = sqrt(nx * nx + ny * ny + nz * nz + dd * dd);
                // a list with multiplicity can place the same point at two positions: the 3x4 system then has a 2-D null
                // space and the reference's SVD returns an arbitrary member of it; such hypotheses are skipped
                const bool same = (e1x == 0 && e1y == 0 && e1z == 0) || (e2x == 0 && e2y == 0 && e2z == 0) || (e1x == e2x && e1y == e2y && e1z == e2z);
                const bool degenerate = same || !(nx * nx + ny * ny + nz * nz > 0);
                int ic = 0;
                if (!degenerate) {
                    const double thr = threshold * n4;
                    for (int q = lane; q < n; q += 32) {
                        const double r = nx * P[3 * q] + ny * P[3 * q + 1] + nz * P[3 * q + 2] + dd;
                        ic += fabs(r) < thr;
                    }
#pragma unroll
                    for (int o = 16; o; o >>= 1) ic += __shfl_xor_sync(0xFFFFFFFFu, ic, o);
                }
                if (lane == 0) { r_ic[warp] = degenerate ? -1 : ic; r_m[warp][0] = nx; r_m[warp][1] = ny; r_m[warp][2] = nz; r_m[warp][3] = dd; r_m[warp][4] = n4; }
            }
            __syncthreads();
            const int hi = min(iterations, h_done + RW);
            for (int hh = h_done; hh < hi && !stop; ++hh) {
                const int w = hh - h_done, ic = r_ic[w];
                used = hh + 1;
                if (ic > best_ic) {
                    best_ic = ic; best = hh;
                    b0 = r_m[w][0]; b1 = r_m[w][1]; b2 = r_m[w][2]; b3 = r_m[w][3]; b4 = r_m[w][4];
                    if ((double)ic > goal && stop_at_goal) stop = true;
                }
            }
            h_done = hi;
            __syncthreads();
        }
        if (threadIdx.x == 0) {
            const double sg = (b1 < 0 ? -1.0 : 1.0) / b4;
            for (int k = 0; k < 4; ++k) model[4 * s + k] = CUDART_NAN;
            if (best >= 0) { model[4 * s] = sg * b0; model[4 * s + 1] = sg * b1; model[4 * s + 2] = sg * b2; model[4 * s + 3] = sg * b3; }
            if (ic_out) ic_out[s] = best_ic;
            if (best_hyp_out) best_hyp_out[s] = best;
            if (hyps_used_out) hyps_used_out[s] = used;
        }
        __syncthreads();
    }
}

// ---- SE(3) prefix products ------------------------------------------------------------------------
struct Rt { double m[12]; };             // row-major [R|t]

__device__ __forceinline__ Rt rt_identity() { Rt a; for (int i = 0; i < 12; ++i) a.m[i] = (i % 5 == 0) ? 1.0 : 0.0; return a; }

__device__ __forceinline__ Rt rt_mul(const Rt &a, const Rt &b) {          // a * b as 4x4 with last row (0,0,0,1)
    Rt c;
#pragma unroll
    for (int r = 0; r < 3; ++r) {
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            double v = a.m[4 * r] * b.m[k] + a.m[4 * r + 1] * b.m[4 + k] + a.m[4 * r + 2] * b.m[8 + k];
            if (k == 3) v += a.m[4 * r + 3];
            c.m[4 * r + k] = v;
        }
    }
    return c;
}

__device__ __forceinline__ Rt rt_shfl_up(const Rt &a, int delta) {
    Rt b;
#pragma unroll
    for (int i = 0; i < 12; ++i) b.m[i] = __shfl_up_sync(0xFFFFFFFFu, a.m[i], delta);
    return b;
}

// poses[0] = I, poses[i+1] = poses[i] * [R_i | s_i t_i]: one CTA per sequence.  Every thread multiplies a contiguous run of
// motions, the run products are scanned (warp shuffles, then across warps through shared memory), and each thread replays its
// run from its prefix.  The association order differs from the reference's left-to-right loop: results agree to rounding.
constexpr int PT = 256;
__global__ void __launch_bounds__(PT) integrate_paths_kernel(int n_seq, const int32_t *__restrict__ seq_offsets, const double *__restrict__ motions,
                                                             const double *__restrict__ scales, double *poses) {
    __shared__ Rt wtot[PT / 32];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    for (int s = blockIdx.x; s < n_seq; s += gridDim.x) {
        const int f0 = seq_offsets[s], n = seq_offsets[s + 1] - f0;
        double *out = poses + 12 * (size_t)(f0 + s);
        const int per = (n + PT - 1) / PT, b = min(n, tid * per), e = min(n, b + per);
        auto load = [&](int i) {
            Rt m; const double *src = motions + 12 * (size_t)(f0 + i); const double sc = scales ? scales[f0 + i] : 1.0;
#pragma unroll
            for (int k = 0; k < 12; ++k) m.m[k] = (k & 3) == 3 ? sc * src[k] : src[k];
            return m;
        };
        Rt run = rt_identity();
        for (int i = b; i < e; ++i) run = rt_mul(run, load(i));
        // inclusive scan of the run products in thread order (non-commutative: earlier * later)
        Rt inc = run;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { const Rt up = rt_shfl_up(inc, o); if (lane >= o) inc = rt_mul(up, inc); }
        if (lane == 31) wtot[warp] = inc;
        __syncthreads();
        Rt pre = rt_identity();                                  // product of all earlier warps
        for (int w = 0; w < warp; ++w) pre = rt_mul(pre, wtot[w]);
        Rt excl = rt_shfl_up(inc, 1);
        if (lane == 0) excl = rt_identity();
        Rt cur = rt_mul(pre, excl);                               // pose before this thread's run
        if (tid == 0) { for (int k = 0; k < 12; ++k) out[k] = cur.m[k]; }
        for (int i = b; i < e; ++i) {
            cur = rt_mul(cur, load(i));
#pragma unroll
            for (int k = 0; k < 12; ++k) out[12 * (size_t)(i + 1) + k] = cur.m[k];
        }
        __syncthreads();
    }
}

// ---- dense depth from the mesh ---------------------------------------------------------------------
// Point location of all W x H integer pixels in the triangulation, done as a rasteriser: one warp per triangle walks the
// pixels of the triangle's bounding box and claims those inside (closed edge functions in float64); a pixel on a shared edge
// is claimed by the smaller triangle index (atomicMin), which makes the result deterministic (scipy's find_simplex returns
// "one of" the simplices there).  tri_id must be pre-filled with a value above every triangle index (mvosr_depth_from_mesh: 0x7F7F7F7F).
__global__ void __launch_bounds__(256) raster_mesh_kernel(int n_tri, const int32_t *__restrict__ tri, const double *__restrict__ uv,
                                                          int width, int height, int32_t *tri_id) {
    const int lane = threadIdx.x & 31, wpb = blockDim.x >> 5;
    for (int t = blockIdx.x * wpb + (threadIdx.x >> 5); t < n_tri; t += gridDim.x * wpb) {
        const int i0 = tri[3 * t], i1 = tri[3 * t + 1], i2 = tri[3 * t + 2];
        const double ax = uv[2 * i0], ay = uv[2 * i0 + 1], bx = uv[2 * i1], by = uv[2 * i1 + 1], cx = uv[2 * i2], cy = uv[2 * i2 + 1];
        const double area = (bx - ax) * (cy - ay) - (by - ay) * (cx - ax);
        if (area == 0.0) continue;
        const double sg = area > 0 ? 1.0 : -1.0;
        const int x0 = max((int)ceil(fmin(ax, fmin(bx, cx))), 0), x1 = min((int)floor(fmax(ax, fmax(bx, cx))), width - 1);
        const int y0 = max((int)ceil(fmin(ay, fmin(by, cy))), 0), y1 = min((int)floor(fmax(ay, fmax(by, cy))), height - 1);
        if (x0 > x1 || y0 > y1) continue;
        const int w = x1 - x0 + 1, n = w * (y1 - y0 + 1);
        for (int i = lane; i < n; i += 32) {
            const int px = x0 + i % w, py = y0 + i / w;
            const double X = px, Y = py;
            const double e0 = sg * ((bx - ax) * (Y - ay) - (by - ay) * (X - ax));
            const double e1 = sg * ((cx - bx) * (Y - by) - (cy - by) * (X - bx));
            const double e2 = sg * ((ax - cx) * (Y - cy) - (ay - cy) * (X - cx));
            if (e0 >= 0 && e1 >= 0 && e2 >= 0) atomicMin(tri_id + (size_t)py * width + px, t);
        }
    }
}

// depth = h / (n . (x_n, y_n, 1)) of the claiming triangle's plane (datas rows: unit normal, height); 0 and id -1 outside.
__global__ void __launch_bounds__(256) mesh_depth_kernel(int width, int height, double fx, double fy, double cx, double cy,
                                                         const double *__restrict__ datas, int32_t *tri_id, double *depth) {
    const int n = width * height;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        const int t = tri_id[i];
        double d = 0.0;
        if (t == 0x7F7F7F7F) tri_id[i] = -1;                      // the fill pattern of mvosr_depth_from_mesh (byte memset)
        else {
            const double xn = ((double)(i % width) - cx) / fx, yn = ((double)(i / width) - cy) / fy;
            d = datas[4 * t + 3] / (datas[4 * t] * xn + datas[4 * t + 1] * yn + datas[4 * t + 2]);
        }
        depth[i] = d;
    }
}

// ---- pose from the essential matrix -------------------------------------------------------------------
// Eigenvectors of the symmetric 3x3 matrix a (cyclic Jacobi, float64), columns of v, eigenvalues in w (unsorted).
__device__ inline void jacobi_eig3(double a[3][3], double v[3][3], double w[3]) {
    for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) v[i][j] = i == j ? 1.0 : 0.0;
    for (int sweep = 0; sweep < 12; ++sweep) {
        const double off = fabs(a[0][1]) + fabs(a[0][2]) + fabs(a[1][2]);
        if (off < 1.0e-300) break;
        for (int p = 0; p < 2; ++p) for (int q = p + 1; q < 3; ++q) {
            if (a[p][q] == 0.0) continue;
            const double theta = (a[q][q] - a[p][p]) / (2.0 * a[p][q]);
            const double t = (theta >= 0 ? 1.0 : -1.0) / (fabs(theta) + sqrt(theta * theta + 1.0));
            const double c = 1.0 / sqrt(t * t + 1.0), sn = t * c;
            for (int k = 0; k < 3; ++k) { const double akp = a[k][p], akq = a[k][q]; a[k][p] = c * akp - sn * akq; a[k][q] = sn * akp + c * akq; }
            for (int k = 0; k < 3; ++k) { const double apk = a[p][k], aqk = a[q][k]; a[p][k] = c * apk - sn * aqk; a[q][k] = sn * apk + c * aqk; }
            for (int k = 0; k < 3; ++k) { const double vkp = v[k][p], vkq = v[k][q]; v[k][p] = c * vkp - sn * vkq; v[k][q] = sn * vkp + c * vkq; }
        }
    }
    for (int i = 0; i < 3; ++i) w[i] = a[i][i];
}

// decomposeEssentialMat: the two rotations U W V^T, U W^T V^T and the translation direction u3 (|t| = 1) of E = U diag(s,s,0) V^T
// with det U = det V = +1.  V from the eigenvectors of E^T E (the two rotations do not depend on the choice inside the
// eigenspace of the double singular value), U = E V / s on the first two columns, u3 = u1 x u2.
__device__ inline void decompose_essential(const double *E, double R1[9], double R2[9], double t[3]) {
    double a[3][3], v[3][3], w[3];
    for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) a[i][j] = E[0 + i] * E[0 + j] + E[3 + i] * E[3 + j] + E[6 + i] * E[6 + j];
    jacobi_eig3(a, v, w);
    int o0 = 0, o1 = 1, o2 = 2;                                   // eigenvalues descending: o2 = the null direction
    if (w[o0] < w[o1]) { int x = o0; o0 = o1; o1 = x; }
    if (w[o1] < w[o2]) { int x = o1; o1 = o2; o2 = x; }
    if (w[o0] < w[o1]) { int x = o0; o0 = o1; o1 = x; }
    double v1[3] = { v[0][o0], v[1][o0], v[2][o0] }, v2[3] = { v[0][o1], v[1][o1], v[2][o1] }, v3[3];
    v3[0] = v1[1] * v2[2] - v1[2] * v2[1]; v3[1] = v1[2] * v2[0] - v1[0] * v2[2]; v3[2] = v1[0] * v2[1] - v1[1] * v2[0];   // det V = +1
    double u1[3], u2[3], u3[3];
    for (int i = 0; i < 3; ++i) { u1[i] = E[3 * i] * v1[0] + E[3 * i + 1] * v1[1] + E[3 * i + 2] * v1[2]; u2[i] = E[3 * i] * v2[0] + E[3 * i + 1] * v2[1] + E[3 * i + 2] * v2[2]; }
    double n1 = sqrt(u1[0] * u1[0] + u1[1] * u1[1] + u1[2] * u1[2]);
    for (int i = 0; i < 3; ++i) u1[i] /= n1;
    const double d12 = u1[0] * u2[0] + u1[1] * u2[1] + u1[2] * u2[2];
    for (int i = 0; i < 3; ++i) u2[i] -= d12 * u1[i];
    double n2 = sqrt(u2[0] * u2[0] + u2[1] * u2[1] + u2[2] * u2[2]);
    for (int i = 0; i < 3; ++i) u2[i] /= n2;
    u3[0] = u1[1] * u2[2] - u1[2] * u2[1]; u3[1] = u1[2] * u2[0] - u1[0] * u2[2]; u3[2] = u1[0] * u2[1] - u1[1] * u2[0];   // det U = +1
    // U W V^T = -u2 v1^T + u1 v2^T + u3 v3^T ;  U W^T V^T = u2 v1^T - u1 v2^T + u3 v3^T   (W = [0 1 0; -1 0 0; 0 0 1])
    for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) {
        const double x = u1[i] * v2[j] - u2[i] * v1[j], y = u3[i] * v3[j];
        R1[3 * i + j] = x + y; R2[3 * i + j] = -x + y;
    }
    t[0] = u3[0]; t[1] = u3[1]; t[2] = u3[2];
}

// One CTA per frame: thread 0 decomposes E, all threads triangulate every correspondence under the four candidates
// (R1,t), (R2,t), (R1,-t), (R2,-t) and count the points that pass recoverPose's mask (in front of both cameras, nearer than
// dist); the candidate with the most points wins (first one on ties, OpenCV's order).
__global__ void __launch_bounds__(256) recover_pose_kernel(int n_frames, const int32_t *__restrict__ offsets,
        const float *__restrict__ cur_u, const float *__restrict__ cur_v, const float *__restrict__ ref_u, const float *__restrict__ ref_v,
        const uint8_t *__restrict__ e_mask, const double *__restrict__ essential, mvosr_config cfg, double *poses, int32_t *n_good) {
    __shared__ Pose cand[4];
    __shared__ int good[4];
    const int tid = threadIdx.x;
    for (int f = blockIdx.x; f < n_frames; f += gridDim.x) {
        if (tid == 0) {
            double R1[9], R2[9], t[3];
            decompose_essential(essential + 9 * (size_t)f, R1, R2, t);
            for (int k = 0; k < 4; ++k) {
                for (int i = 0; i < 9; ++i) cand[k].R[i] = (k & 1) ? R2[i] : R1[i];
                for (int i = 0; i < 3; ++i) cand[k].t[i] = (k & 2) ? -t[i] : t[i];
                good[k] = 0;
            }
        }
        __syncthreads();
        const int base = offsets[f], n = offsets[f + 1] - base;
        int c[4] = { 0, 0, 0, 0 };
        for (int i = tid; i < n; i += blockDim.x) {
            if (e_mask && !e_mask[base + i]) continue;
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                double X, Y, Z, u, v;
                c[k] += triangulate_point(cur_u[base + i], cur_v[base + i], ref_u[base + i], ref_v[base + i], cand[k],
                                          cfg.fx, cfg.fy, cfg.cx, cfg.cy, cfg.triangulation_max_depth, X, Y, Z, u, v);
            }
        }
#pragma unroll
        for (int k = 0; k < 4; ++k) {
#pragma unroll
            for (int o = 16; o; o >>= 1) c[k] += __shfl_xor_sync(0xFFFFFFFFu, c[k], o);
            if ((tid & 31) == 0 && c[k]) atomicAdd(&good[k], c[k]);
        }
        __syncthreads();
        if (tid == 0) {
            int best = 0;
            for (int k = 1; k < 4; ++k) if (good[k] > good[best]) best = k;
            double *P = poses + 12 * (size_t)f;
            for (int r = 0; r < 3; ++r) { for (int cc = 0; cc < 3; ++cc) P[4 * r + cc] = cand[best].R[3 * r + cc]; P[4 * r + 3] = cand[best].t[r]; }
            if (n_good) for (int k = 0; k < 4; ++k) n_good[4 * f + k] = good[k];
        }
        __syncthreads();
    }
}

// recoverPose's per-correspondence mask under a GIVEN pose (in front of both cameras, nearer than dist), optionally ANDed with
// the essential-matrix inlier mask: mask_bool & mask_e_bool of src/thirdparty/MonocularVO/visual_odometry.py:134-136.  The
// test is triangulate_point itself, so the mask marks exactly the correspondences triangulate_kernel keeps (in their order).
__global__ void __launch_bounds__(256) pose_mask_kernel(int n_frames, const int32_t *__restrict__ offsets,
        const float *__restrict__ cur_u, const float *__restrict__ cur_v, const float *__restrict__ ref_u, const float *__restrict__ ref_v,
        const uint8_t *__restrict__ e_mask, const double *__restrict__ poses, mvosr_config cfg, uint8_t *mask_out) {
    for (int f = blockIdx.x; f < n_frames; f += gridDim.x) {
        const int base = offsets[f], n = offsets[f + 1] - base;
        Pose pose;
        for (int i = 0; i < 9; ++i) pose.R[i] = poses[12 * (size_t)f + (i / 3) * 4 + (i % 3)];
        pose.t[0] = poses[12 * (size_t)f + 3]; pose.t[1] = poses[12 * (size_t)f + 7]; pose.t[2] = poses[12 * (size_t)f + 11];
        for (int i = threadIdx.x; i < n; i += blockDim.x) {
            double X, Y, Z, u, v;
            bool ok = triangulate_point(cur_u[base + i], cur_v[base + i], ref_u[base + i], ref_v[base + i], pose,
                                        cfg.fx, cfg.fy, cfg.cx, cfg.cy, cfg.triangulation_max_depth, X, Y, Z, u, v);
            if (e_mask) ok = ok && e_mask[base + i] != 0;
            mask_out[base + i] = ok ? 1 : 0;
        }
    }
}

}  // namespace mvosr
