// find_essential_kernel: five-point RANSAC per frame (first half; the numerics are in five_point.cuh).
//   replaces  cv2.findEssentialMat(px_cur, px_ref, K, RANSAC, 0.999, threshold)   src/thirdparty/MonocularVO/visual_odometry.py:100-102,129-130
// One CTA per frame.  A round gives every thread one hypothesis: five positions from the Philox stream, the minimal solver (up
// to ten candidates, kept in the thread's local memory); the round's candidates are then numbered consecutively and scored 128
// at a time, one per thread (threads own 0..10 of them: scored in place most lanes would idle), all threads walking the
// frame's correspondences together -- staged tile by tile in shared memory in normalised float64 coordinates, so every lane
// reads the same point (a broadcast) -- and counting the Sampson inliers of their candidate.  The winner is the candidate with
// the most inliers; ties go to the lowest (hypothesis, candidate) pair.  `hypotheses` is the maximum (OpenCV's maxIters); with confidence > 0 the loop stops after the
// first round of FP5_ROUND hypotheses at whose end  hypotheses tried >= log(1 - confidence) / log(1 - w^5),  w = best inlier
// ratio -- OpenCV's adaptive count (prob = 0.999 in the reference's call), evaluated per round instead of per sample so that
// it is a property of the stream definition, not of the schedule.  confidence = 0: all hypotheses.  The last pass writes the
// winner's inlier mask.
#pragma once
#include <stdint.h>
#include "../../include/mvosr.h"
#include "five_point.cuh"
#include "five_point_tables.h"

namespace mvosr {

__constant__ fp5::Tables c_fp5_tables = MVOSR_FP5_TABLES_INIT;

constexpr int FP5_ROUND = 128;           // hypotheses per round: part of the stream definition (the stopping rule is checked per round)
constexpr int FP5_THREADS = FP5_ROUND;
#ifndef MVOSR_FP5_MIN_BLOCKS
#define MVOSR_FP5_MIN_BLOCKS 4       // CTAs per SM the register allocation aims at (128 registers); sweep with MVOSR_NVCC_EXTRA=-DMVOSR_FP5_MIN_BLOCKS=2|3
#endif
constexpr int FP5_TILE = 512;            // correspondences staged per tile (16 KB of shared memory)

__global__ void __launch_bounds__(FP5_THREADS, MVOSR_FP5_MIN_BLOCKS) find_essential_kernel(int n_frames, const int32_t *__restrict__ offsets,
        const float *__restrict__ cur_u, const float *__restrict__ cur_v, const float *__restrict__ ref_u, const float *__restrict__ ref_v,
        double fx, double fy, double cx, double cy, int hypotheses, double threshold_px, double confidence, uint64_t seed,
        const int32_t *__restrict__ frame_index, int seq_id,
        double *essential, uint8_t *e_mask, int32_t *n_inliers, int32_t *best_hyp, int32_t *hyps_used) {
    __shared__ double s_pt[FP5_TILE][4];
    __shared__ double s_cand[FP5_THREADS][9];                       // one chunk of the round's compacted candidate list
    __shared__ uint32_t s_cid[FP5_THREADS];                         // hypothesis * 16 + candidate of each entry
    __shared__ int s_ns[FP5_THREADS];
    __shared__ unsigned long long s_key[FP5_THREADS / 32];
    __shared__ unsigned long long s_best_key;
    __shared__ double s_best_E[9];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const double thr = threshold_px / (0.5 * (fx + fy)), thr2 = thr * thr;
    for (int f = blockIdx.x; f < n_frames; f += gridDim.x) {
        const int base = offsets[f], n = offsets[f + 1] - base;
        const uint32_t fidx = frame_index ? (uint32_t)frame_index[f] : (uint32_t)f;
        if (tid == 0) {
            s_best_key = 0ull;
            for (int i = 0; i < 9; ++i) s_best_E[i] = 0.0;
        }
        __syncthreads();
        int used = 0;
        if (n >= 5) {                                               // uniform over the CTA
            unsigned long long best = 0ull;                         // the best key so far: the same value in every thread
            for (int h0 = 0; h0 < hypotheses; h0 += FP5_THREADS) {
                // -- solve: one hypothesis per thread, candidates in the thread's local memory
                const int hyp = h0 + tid;
                double E[10][9];
                int ns = 0;
                if (hyp < hypotheses) {
                    int idx[5];
                    fp5::sample5(seed, (uint32_t)hyp, fidx, (uint32_t)seq_id, (uint32_t)n, idx);
                    double x1[10], x2[10];
                    for (int k = 0; k < 5; ++k) {
                        const int p = base + idx[k];
                        x1[2 * k] = ((double)cur_u[p] - cx) / fx; x1[2 * k + 1] = ((double)cur_v[p] - cy) / fy;
                        x2[2 * k] = ((double)ref_u[p] - cx) / fx; x2[2 * k + 1] = ((double)ref_v[p] - cy) / fy;
                    }
                    ns = fp5::solve(x1, x2, c_fp5_tables, E);
                }
                // -- compact: threads end up with 0..10 candidates each; scoring them where they are would leave most lanes of a
                //    warp idle most of the time, so the round's candidates are numbered consecutively (hypothesis-major) and
                //    scored 128 at a time, one per thread
                s_ns[tid] = ns;
                __syncthreads();
                int my_first = 0, total = 0;
                for (int t = 0; t < FP5_THREADS; ++t) { const int v = s_ns[t]; if (t < tid) my_first += v; total += v; }
                __syncthreads();                                    // s_ns is rewritten by the next round, and a round without candidates has no other barrier
                for (int c0 = 0; c0 < total; c0 += FP5_THREADS) {
                    // (the previous chunk's entries were copied to registers before the barriers of its tile loop: free to overwrite)
                    for (int k = 0; k < ns; ++k) {
                        const int slot = my_first + k - c0;
                        if (slot >= 0 && slot < FP5_THREADS) {
                            for (int i = 0; i < 9; ++i) s_cand[slot][i] = E[k][i];
                            s_cid[slot] = (uint32_t)(hyp * 16 + k);
                        }
                    }
                    __syncthreads();
                    const bool have = c0 + tid < total;
                    double e[9];
                    for (int i = 0; i < 9; ++i) e[i] = have ? s_cand[tid][i] : 0.0;
                    const uint32_t cid = have ? s_cid[tid] : 0u;
                    // -- score: all threads walk the frame's correspondences together, tile by tile
                    int cnt = 0;
                    for (int t0 = 0; t0 < n; t0 += FP5_TILE) {
                        const int m = min(FP5_TILE, n - t0);
                        __syncthreads();                            // the previous tile has been consumed
                        for (int i = tid; i < m; i += FP5_THREADS) {
                            const int p = base + t0 + i;
                            s_pt[i][0] = ((double)cur_u[p] - cx) / fx; s_pt[i][1] = ((double)cur_v[p] - cy) / fy;
                            s_pt[i][2] = ((double)ref_u[p] - cx) / fx; s_pt[i][3] = ((double)ref_v[p] - cy) / fy;
                        }
                        __syncthreads();
                        if (have)
                            for (int i = 0; i < m; ++i) cnt += fp5::sampson_inlier(e, s_pt[i][0], s_pt[i][1], s_pt[i][2], s_pt[i][3], thr2) ? 1 : 0;
                    }
                    // -- select: most inliers, ties to the lowest (hypothesis, candidate) pair
                    const unsigned long long key = have && cnt > 0 ? ((unsigned long long)(uint32_t)cnt << 32) | (unsigned long long)(0xFFFFFFFFu - cid) : 0ull;
                    unsigned long long wmax = key;
#pragma unroll
                    for (int o = 16; o; o >>= 1) {
                        const unsigned long long other = __shfl_xor_sync(0xFFFFFFFFu, wmax, o);
                        if (other > wmax) wmax = other;
                    }
                    if (lane == 0) s_key[warp] = wmax;
                    __syncthreads();
                    unsigned long long bmax = s_key[0];
                    for (int w = 1; w < FP5_THREADS / 32; ++w) if (s_key[w] > bmax) bmax = s_key[w];
                    if (key != 0ull && key == bmax && bmax > best) {  // keys are unique: exactly one thread
                        s_best_key = bmax;
                        for (int i = 0; i < 9; ++i) s_best_E[i] = e[i];
                    }
                    if (bmax > best) best = bmax;
                    __syncthreads();                                // s_key is rewritten by the next chunk
                }
                used = min(hypotheses, h0 + FP5_THREADS);
                if (fp5::enough_hypotheses(used, (int)(best >> 32), n, confidence)) break;   // uniform
            }
        }
        __syncthreads();
        const unsigned long long win = s_best_key;
        double e[9];
        for (int i = 0; i < 9; ++i) e[i] = s_best_E[i];
        if (e_mask)
            for (int i = tid; i < n; i += FP5_THREADS) {
                const int p = base + i;
                const double ax = ((double)cur_u[p] - cx) / fx, ay = ((double)cur_v[p] - cy) / fy;
                const double bx = ((double)ref_u[p] - cx) / fx, by = ((double)ref_v[p] - cy) / fy;
                e_mask[p] = (win != 0ull && fp5::sampson_inlier(e, ax, ay, bx, by, thr2)) ? 1 : 0;
            }
        if (tid == 0) {
            for (int i = 0; i < 9; ++i) essential[9 * (size_t)f + i] = e[i];
            if (n_inliers) n_inliers[f] = (int32_t)(win >> 32);
            if (best_hyp) best_hyp[f] = win != 0ull ? (int32_t)((0xFFFFFFFFu - (uint32_t)(win & 0xFFFFFFFFull)) >> 4) : -1;
            if (hyps_used) hyps_used[f] = used;
        }
        __syncthreads();                                            // s_best_* are reset by thread 0 for the next frame
    }
}

}  // namespace mvosr
