// TEST INFRASTRUCTURE: runs the SOURCE of bucket_kernel (csrc/bucket_kernel.cuh) on the host, one OS thread per CUDA thread
// (same pthread mapping as fp5_kernel_emu.cpp: __syncthreads = barrier, __shared__ = static storage, atomicOr = a host atomic).
// Nothing in the product links or loads this file.
#include <pthread.h>
#include <stdint.h>
#include <algorithm>
#include <vector>

struct Dim3 { int x; };
static thread_local Dim3 threadIdx, blockIdx;
static Dim3 blockDim, gridDim;
static pthread_barrier_t g_cta_barrier;
static inline void __syncthreads() { pthread_barrier_wait(&g_cta_barrier); }
static inline int atomicOr(int *p, int v) { return __atomic_fetch_or(p, v, __ATOMIC_RELAXED); }
using std::min;
#define __global__
#define __launch_bounds__(...)
#define __restrict__
#define __shared__ static

#include "../../mvoscalerecovery_b200/csrc/bucket_kernel.cuh"

namespace {
struct Args { int n_frames; const int32_t *off; const float *u, *v; int bs, dens; uint64_t seed; const int32_t *fi; int seq; int32_t *idx, *n_out; uint8_t *st; };
struct ThreadArg { const Args *a; int tid, bid; };
void *run_thread(void *p) {
    const ThreadArg *t = (const ThreadArg *)p;
    threadIdx.x = t->tid; blockIdx.x = t->bid;
    const Args &a = *t->a;
    mvosr::bucket_kernel(a.n_frames, a.off, a.u, a.v, a.bs, a.dens, a.seed, a.fi, a.seq, a.idx, a.n_out, a.st);
    return nullptr;
}
}  // namespace

extern "C" int bucket_emu(int32_t n_frames, const int32_t *off, const float *u, const float *v, int32_t bs, int32_t dens, uint64_t seed,
                          const int32_t *fi, int32_t seq, int32_t *idx, int32_t *n_out, uint8_t *st, int32_t grid) {
    const int T = mvosr::BUCKET_THREADS;
    const Args a = { n_frames, off, u, v, bs, dens, seed, fi, seq, idx, n_out, st };
    blockDim.x = T; gridDim.x = grid;
    pthread_barrier_init(&g_cta_barrier, nullptr, T);
    pthread_attr_t attr;
    pthread_attr_init(&attr);
    pthread_attr_setstacksize(&attr, 1 << 18);
    for (int b = 0; b < grid; ++b) {
        std::vector<pthread_t> th(T);
        std::vector<ThreadArg> ta(T);
        for (int t = 0; t < T; ++t) { ta[t] = { &a, t, b }; if (pthread_create(&th[t], &attr, run_thread, &ta[t])) return -1; }
        for (int t = 0; t < T; ++t) pthread_join(th[t], nullptr);
    }
    pthread_barrier_destroy(&g_cta_barrier);
    return 0;
}
