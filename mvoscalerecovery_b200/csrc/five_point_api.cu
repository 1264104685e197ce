// libmvosr.so, second translation unit: the five-point essential-matrix RANSAC.  It is compiled with -fmad=false
// (mvoscalerecovery_b200/build.py): without fused multiply-adds every operation of the solver and of the Sampson test is an
// individually rounded IEEE-754 double operation, so the kernel's results -- winning hypothesis, inlier mask, essential matrix --
// are bit-identical to the same arithmetic evaluated on any host, which is what lets the parity tests demand equality instead
// of a tolerance (tests/test_gpu_zz_essential.py).  The cost is one extra rounding per multiply-add of an FP64 kernel whose
// time goes into dependent chains, not throughput (profiles/README.md).
#include <cuda_runtime.h>
#include <stdint.h>
#include "handle.h"
#include "five_point_kernel.cuh"

using namespace mvosr;

extern "C" {

int mvosr_find_essential_frames(mvosr_handle *h, int32_t n_frames, const int32_t *offsets,
                                const float *cur_u, const float *cur_v, const float *ref_u, const float *ref_v,
                                int32_t hypotheses, double threshold_px, double confidence, uint64_t seed, const int32_t *frame_index, int32_t seq_id,
                                double *essential, uint8_t *e_mask_out, int32_t *n_inliers, int32_t *best_hyp, int32_t *hyps_used, void *stream) {
    if (!h || n_frames < 0 || !offsets || !cur_u || !cur_v || !ref_u || !ref_v || !essential || hypotheses < 1 || hypotheses > (1 << 24) ||
        !(threshold_px > 0.0) || !(confidence >= 0.0))
        return MVOSR_E_INVALID;
    if (n_frames == 0) return MVOSR_OK;
    CK(cudaSetDevice(h->device));
    const int grid = min(n_frames, 16 * h->num_sms);
    find_essential_kernel<<<grid, FP5_THREADS, 0, (cudaStream_t)stream>>>(n_frames, offsets, cur_u, cur_v, ref_u, ref_v,
        h->cfg.fx, h->cfg.fy, h->cfg.cx, h->cfg.cy, hypotheses, threshold_px, confidence, seed, frame_index, seq_id, essential, e_mask_out, n_inliers, best_hyp, hyps_used);
    CK(cudaGetLastError());
    h->launches += 1;
    return MVOSR_OK;
}

}  // extern "C"
