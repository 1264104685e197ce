// bucket_kernel: feature bucketing of the VO front-end -- replaces bucket(features, bucket_size=30, density=2) of
// src/detector.py:65-95 (also FeatureDetector.bucket, :18-47): features are binned into bucket_size x bucket_size pixel cells,
// every cell keeps at most `density` of its features, chosen at random, and the survivors are listed cell by cell (rows of cells
// top to bottom, cells left to right).  The reference shuffles each cell with numpy's global RNG (not reproducible); here the
// order inside a cell is DEFINED by a Philox key per feature:
//     r_i = word 0 of Philox4x32-10(counter = (i, frame, seq, 3), key = seed);  a cell keeps its `density` smallest (r_i, i)
// (a uniformly random subset, like the shuffle's first `density` entries).  Cell of a feature = (int(u) / bucket_size,
// int(v) / bucket_size), as the reference computes it; coordinates must be non-negative pixels.
// One CTA per frame: 64-bit keys  cell << 44 | r << 12 | i  sorted by a bitonic network in shared memory (cell-major, then r),
// an element survives when the one `density` places before it belongs to another cell, survivors are compacted in order.
// Only __syncthreads is used, so tests/host_sim can run this source under the pthread emulation.
#pragma once
#include <stdint.h>
#include "five_point.cuh"          // fp5::philox (host/device)

namespace mvosr {

constexpr int BUCKET_THREADS = 256;
constexpr int BUCKET_CAP = 4096;         // features per frame (12 index bits of the key; 32 KB of shared memory)

__global__ void __launch_bounds__(BUCKET_THREADS) bucket_kernel(int n_frames, const int32_t *__restrict__ offsets,
        const float *__restrict__ fu, const float *__restrict__ fv, int bucket_size, int density, uint64_t seed,
        const int32_t *__restrict__ frame_index, int seq_id, int32_t *out_index, int32_t *n_out, uint8_t *status) {
    __shared__ unsigned long long s_key[BUCKET_CAP];
    __shared__ int s_cnt[BUCKET_THREADS];
    __shared__ int s_bad;
    const int tid = threadIdx.x;
    for (int f = blockIdx.x; f < n_frames; f += gridDim.x) {
        const int base = offsets[f], n = offsets[f + 1] - base;
        const uint32_t fidx = frame_index ? (uint32_t)frame_index[f] : (uint32_t)f;
        if (tid == 0) s_bad = 0;
        __syncthreads();
        if (n > BUCKET_CAP || n <= 0) {                             // uniform: too large for the key layout, or empty
            if (tid == 0) { n_out[f] = 0; if (status) status[f] = n > BUCKET_CAP ? 2 : 0; }
            __syncthreads();
            continue;
        }
        int N = 1;
        while (N < n) N <<= 1;
        for (int i = tid; i < N; i += BUCKET_THREADS) {
            unsigned long long key = ~0ull;                         // padding sorts last
            if (i < n) {
                const float u = fu[base + i], v = fv[base + i];
                const bool ok = u >= 0.0f && v >= 0.0f && u < 1024.0f * (float)bucket_size && v < 1023.0f * (float)bucket_size;
                const int cu = ok ? (int)u / bucket_size : 0, cv = ok ? (int)v / bucket_size : 0;
                if (!ok) atomicOr(&s_bad, 1);                       // negative, non-finite or beyond 1024 x 1023 cells: the frame is rejected
                uint32_t r[4];
                fp5::philox((uint32_t)i, fidx, (uint32_t)seq_id, 3u, (uint32_t)seed, (uint32_t)(seed >> 32), r);
                key = ((unsigned long long)(uint32_t)(cv * 1024 + cu) << 44) | ((unsigned long long)r[0] << 12) | (unsigned long long)i;
            }
            s_key[i] = key;
        }
        __syncthreads();
        if (s_bad) {                                                // uniform after the barrier
            if (tid == 0) { n_out[f] = 0; if (status) status[f] = 1; }
            __syncthreads();
            continue;
        }
        for (int k = 2; k <= N; k <<= 1)                            // bitonic sort, ascending
            for (int j = k >> 1; j > 0; j >>= 1) {
                for (int i = tid; i < N; i += BUCKET_THREADS) {
                    const int p = i ^ j;
                    if (p > i) {
                        const unsigned long long a = s_key[i], b = s_key[p];
                        if (((i & k) == 0) == (a > b)) { s_key[i] = b; s_key[p] = a; }
                    }
                }
                __syncthreads();
            }
        // survivors: fewer than `density` predecessors in the same cell; every thread owns a contiguous range of positions
        const int per = (n + BUCKET_THREADS - 1) / BUCKET_THREADS, p0 = min(n, tid * per), p1 = min(n, p0 + per);
        int c = 0;
        for (int p = p0; p < p1; ++p) c += (p < density || (s_key[p - density] >> 44) != (s_key[p] >> 44)) ? 1 : 0;
        s_cnt[tid] = c;
        __syncthreads();
        int before = 0, total = 0;
        for (int t = 0; t < BUCKET_THREADS; ++t) { const int v = s_cnt[t]; if (t < tid) before += v; total += v; }
        for (int p = p0; p < p1; ++p)
            if (p < density || (s_key[p - density] >> 44) != (s_key[p] >> 44)) out_index[base + before++] = (int32_t)(s_key[p] & 0xFFFull);
        if (tid == 0) { n_out[f] = total; if (status) status[f] = 0; }
        __syncthreads();                                            // s_key, s_cnt and s_bad are reused by the next frame
    }
}

}  // namespace mvosr
