// TEST INFRASTRUCTURE: runs the SOURCE of find_essential_kernel (csrc/five_point_kernel.cuh) on the host, one OS thread per
// CUDA thread, so that the CPU suite exercises the kernel's own control flow -- tile staging, barriers, the warp-shuffle key
// reduction, the winner hand-over between rounds, the mask pass -- and not only the numerics it calls.  The CUDA vocabulary
// the kernel uses is mapped onto pthreads below: __syncthreads = a barrier over the CTA's threads, __shfl_xor_sync = an
// exchange through a per-warp slot array between two warp barriers, __shared__ = static storage (CTAs run one after the
// other), __constant__ = a const global.  Nothing in the product links or loads this file.
#include <pthread.h>
#include <stdint.h>
#include <algorithm>
#include <vector>

struct Dim3 { int x; };
static thread_local Dim3 threadIdx, blockIdx;
static Dim3 blockDim, gridDim;
static pthread_barrier_t g_cta_barrier, g_warp_barrier[64];
static unsigned long long g_warp_slots[64][32];

static inline void __syncthreads() { pthread_barrier_wait(&g_cta_barrier); }
static inline unsigned long long __shfl_xor_sync(unsigned, unsigned long long v, int lane_mask) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    g_warp_slots[warp][lane] = v;
    pthread_barrier_wait(&g_warp_barrier[warp]);
    const unsigned long long r = g_warp_slots[warp][lane ^ lane_mask];
    pthread_barrier_wait(&g_warp_barrier[warp]);
    return r;
}
using std::min;
#define __global__
#define __launch_bounds__(...)
#define __restrict__
#define __shared__ static
#define __constant__ static const

#include "../../mvoscalerecovery_b200/csrc/five_point_kernel.cuh"

namespace {
struct Args {
    int n_frames; const int32_t *offsets; const float *cu, *cv, *ru, *rv; double fx, fy, cx, cy; int hyps; double thr, conf; uint64_t seed;
    const int32_t *frame_index; int seq; double *E; uint8_t *mask; int32_t *cnt, *hyp, *used;
};
struct ThreadArg { const Args *a; int tid, bid; };

void *run_thread(void *p) {
    const ThreadArg *t = (const ThreadArg *)p;
    threadIdx.x = t->tid; blockIdx.x = t->bid;
    const Args &a = *t->a;
    mvosr::find_essential_kernel(a.n_frames, a.offsets, a.cu, a.cv, a.ru, a.rv, a.fx, a.fy, a.cx, a.cy, a.hyps, a.thr, a.conf, a.seed, a.frame_index, a.seq,
                                 a.E, a.mask, a.cnt, a.hyp, a.used);
    return nullptr;
}
}  // namespace

extern "C" int fp5_emu_find_essential(int32_t n_frames, const int32_t *offsets, const float *cu, const float *cv, const float *ru, const float *rv,
                                      double fx, double fy, double cx, double cy, int32_t hyps, double thr, double conf, uint64_t seed,
                                      const int32_t *frame_index, int32_t seq, double *E, uint8_t *mask, int32_t *cnt, int32_t *hyp, int32_t *used, int32_t grid) {
    const int T = mvosr::FP5_THREADS;
    const Args a = { n_frames, offsets, cu, cv, ru, rv, fx, fy, cx, cy, hyps, thr, conf, seed, frame_index, seq, E, mask, cnt, hyp, used };
    blockDim.x = T; gridDim.x = grid;
    pthread_barrier_init(&g_cta_barrier, nullptr, T);
    for (int w = 0; w < T / 32; ++w) pthread_barrier_init(&g_warp_barrier[w], nullptr, 32);
    pthread_attr_t attr;
    pthread_attr_init(&attr);
    pthread_attr_setstacksize(&attr, 1 << 20);
    for (int b = 0; b < grid; ++b) {                               // CTAs one after the other: __shared__ is static storage
        std::vector<pthread_t> th(T);
        std::vector<ThreadArg> ta(T);
        for (int t = 0; t < T; ++t) { ta[t] = { &a, t, b }; if (pthread_create(&th[t], &attr, run_thread, &ta[t])) return -1; }
        for (int t = 0; t < T; ++t) pthread_join(th[t], nullptr);
    }
    pthread_barrier_destroy(&g_cta_barrier);
    for (int w = 0; w < T / 32; ++w) pthread_barrier_destroy(&g_warp_barrier[w]);
    return 0;
}
