// Philox4x32-10 counter-based generator and the RANSAC position stream (device side).
//
// The reference draws its RANSAC samples with CPython random.sample after random.seed(None)
// (src/thirdparty/Ransac/ransac.py:6,10) and is therefore not reproducible; parity needs a
// defined stream.  Definition (also implemented, independently, by the test harness):
//   key = (seed lo, seed hi); counter = (hypothesis, frame, sequence, 0)
//   i0 = (r0*N)>>32 ; i1 = (r1*(N-1))>>32, +1 if >= i0 ; i2 = (r2*(N-2))>>32, +1 for each of the
//   two earlier positions (taken in ascending order) it is >= to.   Three distinct positions.
#pragma once
#include <stdint.h>

namespace mvosr {

__device__ __forceinline__ void philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3,
                                              uint32_t k0, uint32_t k1, uint32_t out[4]) {
#pragma unroll
    for (int r = 0; r < 10; ++r) {
        uint32_t hi0 = __umulhi(0xD2511F53u, c0), lo0 = 0xD2511F53u * c0;
        uint32_t hi1 = __umulhi(0xCD9E8D57u, c2), lo1 = 0xCD9E8D57u * c2;
        uint32_t n0 = hi1 ^ c1 ^ k0, n2 = hi0 ^ c3 ^ k1;
        c0 = n0; c1 = lo1; c2 = n2; c3 = lo0;
        k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
    }
    out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}

__device__ __forceinline__ void sample3_positions(uint64_t seed, uint32_t hyp, uint32_t frame, uint32_t seq,
                                                  uint32_t n, uint32_t &i0, uint32_t &i1, uint32_t &i2) {
    uint32_t r[4];
    philox4x32_10(hyp, frame, seq, 0u, (uint32_t)seed, (uint32_t)(seed >> 32), r);
    i0 = __umulhi(r[0], n);
    i1 = __umulhi(r[1], n - 1u);
    if (i1 >= i0) ++i1;
    i2 = __umulhi(r[2], n - 2u);
    uint32_t lo = min(i0, i1), hi = max(i0, i1);
    if (i2 >= lo) ++i2;
    if (i2 >= hi) ++i2;
}

}  // namespace mvosr
