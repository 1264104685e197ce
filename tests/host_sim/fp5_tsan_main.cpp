// TEST INFRASTRUCTURE: driver of the ThreadSanitizer run of the kernel emulation (scripts/tsan_five_point_kernel.sh): synthetic frames of
// 700 / 3 / 1300 / 40 / 513 correspondences (several tiles, a tile boundary, a frame without a model), 300 hypotheses (three rounds, the last partial), 2 CTAs.
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <math.h>
#include <vector>
extern "C" int fp5_emu_find_essential(int32_t, const int32_t *, const float *, const float *, const float *, const float *, double, double, double, double,
                                      int32_t, double, double, uint64_t, const int32_t *, int32_t, double *, uint8_t *, int32_t *, int32_t *, int32_t *, int32_t);
int main() {
    const double fx = 718.856, fy = 718.856, cx = 607.1928, cy = 185.2157;
    const int F = 5; const int lens[F] = { 700, 3, 1300, 40, 513 };
    std::vector<int32_t> off(F + 1, 0);
    for (int f = 0; f < F; ++f) off[f + 1] = off[f] + lens[f];
    const int M = off[F];
    std::vector<float> cu(M), cv(M), ru(M), rv(M);
    srand(1);
    auto U = [] { return rand() / (double)RAND_MAX; };
    for (int i = 0; i < M; ++i) {
        const double X = -8 + 16 * U(), Y = 1.7 - 3 * U() * (i % 3 == 0), Z = 5 + 35 * U();
        const double a = 0.01, tx = 0.05, ty = -0.02, tz = 0.99;      // small yaw + forward motion
        const double X2 = cos(a) * X + sin(a) * Z + tx, Y2 = Y + ty, Z2 = -sin(a) * X + cos(a) * Z + tz;
        cu[i] = (float)(fx * X / Z + cx); cv[i] = (float)(fy * Y / Z + cy);
        ru[i] = (float)(fx * X2 / Z2 + cx); rv[i] = (float)(fy * Y2 / Z2 + cy);
        if (i % 5 == 0) { ru[i] = (float)(1241 * U()); rv[i] = (float)(376 * U()); }
    }
    std::vector<double> E(9 * F); std::vector<uint8_t> mask(M); std::vector<int32_t> cnt(F), hyp(F), used(F);
    const int rc = fp5_emu_find_essential(F, off.data(), cu.data(), cv.data(), ru.data(), rv.data(), fx, fy, cx, cy, 300, 0.5, 0.0, 99, nullptr, 0,
                                          E.data(), mask.data(), cnt.data(), hyp.data(), used.data(), 2);
    for (int f = 0; f < F; ++f) printf("frame %d n %d inliers %d hyp %d used %d\n", f, lens[f], cnt[f], hyp[f], used[f]);
    return rc;
}
