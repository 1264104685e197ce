// TEST INFRASTRUCTURE: compiles the __host__ __device__ numerics of csrc/five_point.cuh for the host (g++, no CUDA), so that
// the CPU suite can check the very functions find_essential_kernel runs against oracle/five_point_plan.py without a GPU.
// fp5_host_ransac_frame replays the kernel's selection rule (most inliers, ties to the lowest (hypothesis, candidate) pair)
// sequentially.  Nothing in the product links or loads this file.
#include <stdint.h>
#include <string.h>
#include "../../mvoscalerecovery_b200/csrc/five_point.cuh"
#include "../../mvoscalerecovery_b200/csrc/five_point_tables.h"

using namespace mvosr;
static const fp5::Tables g_tables = MVOSR_FP5_TABLES_INIT;

extern "C" {

int fp5_host_solve(const double *x1, const double *x2, double *E_out /* [10][9] */) {
    double E[10][9];
    const int n = fp5::solve(x1, x2, g_tables, E);
    memcpy(E_out, E, sizeof(double) * 9 * (size_t)n);
    return n;
}

void fp5_host_sample5(uint64_t seed, uint32_t hyp, uint32_t frame, uint32_t seq, uint32_t n, int32_t *idx) {
    int i[5];
    fp5::sample5(seed, hyp, frame, seq, n, i);
    for (int k = 0; k < 5; ++k) idx[k] = i[k];
}

void fp5_host_philox(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t k0, uint32_t k1, uint32_t *out) {
    fp5::philox(c0, c1, c2, c3, k0, k1, out);
}

void fp5_host_ransac_frame(int32_t n, const float *cur_u, const float *cur_v, const float *ref_u, const float *ref_v,
                           double fx, double fy, double cx, double cy, int32_t hypotheses, double threshold_px, double confidence, uint64_t seed,
                           uint32_t frame, uint32_t seq, double *essential, uint8_t *e_mask, int32_t *n_inliers, int32_t *best_hyp, int32_t *hyps_used) {
    const double thr = threshold_px / (0.5 * (fx + fy)), thr2 = thr * thr;
    unsigned long long win = 0ull;
    double best[9] = { 0, 0, 0, 0, 0, 0, 0, 0, 0 };
    int used = 0;
    if (n >= 5)
        for (int hyp = 0; hyp < hypotheses; ++hyp) {
            if (hyp > 0 && hyp % 128 == 0 && fp5::enough_hypotheses(hyp, (int)(win >> 32), n, confidence)) break;   // the rule is checked per round of 128
            used = hyp + 1;
            int idx[5];
            fp5::sample5(seed, (uint32_t)hyp, frame, seq, (uint32_t)n, idx);
            double x1[10], x2[10], E[10][9];
            for (int k = 0; k < 5; ++k) {
                const int p = idx[k];
                x1[2 * k] = ((double)cur_u[p] - cx) / fx; x1[2 * k + 1] = ((double)cur_v[p] - cy) / fy;
                x2[2 * k] = ((double)ref_u[p] - cx) / fx; x2[2 * k + 1] = ((double)ref_v[p] - cy) / fy;
            }
            const int ns = fp5::solve(x1, x2, g_tables, E);
            int bk = -1, bc = 0;
            for (int k = 0; k < ns; ++k) {
                int c = 0;
                for (int i = 0; i < n; ++i)
                    c += fp5::sampson_inlier(E[k], ((double)cur_u[i] - cx) / fx, ((double)cur_v[i] - cy) / fy,
                                             ((double)ref_u[i] - cx) / fx, ((double)ref_v[i] - cy) / fy, thr2) ? 1 : 0;
                if (c > bc) { bc = c; bk = k; }
            }
            if (bk < 0) continue;
            const unsigned long long key = ((unsigned long long)(uint32_t)bc << 32) | (unsigned long long)(0xFFFFFFFFu - (uint32_t)(hyp * 16 + bk));
            if (key > win) { win = key; memcpy(best, E[bk], sizeof(best)); }
        }
    for (int i = 0; i < n; ++i)
        e_mask[i] = (win != 0ull && fp5::sampson_inlier(best, ((double)cur_u[i] - cx) / fx, ((double)cur_v[i] - cy) / fy,
                                                        ((double)ref_u[i] - cx) / fx, ((double)ref_v[i] - cy) / fy, thr2)) ? 1 : 0;
    memcpy(essential, best, sizeof(best));
    *hyps_used = used;
    *n_inliers = (int32_t)(win >> 32);
    *best_hyp = win != 0ull ? (int32_t)((0xFFFFFFFFu - (uint32_t)(win & 0xFFFFFFFFull)) >> 4) : -1;
}

}  // extern "C"
