// TEST INFRASTRUCTURE: driver of the ThreadSanitizer run of the bucket kernel emulation (scripts/tsan_five_point_kernel.sh).
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <vector>
extern "C" int bucket_emu(int32_t, const int32_t *, const float *, const float *, int32_t, int32_t, uint64_t, const int32_t *, int32_t, int32_t *, int32_t *, uint8_t *, int32_t);
int main() {
    const int F = 4; const int lens[F] = { 1500, 0, 4096, 37 };
    std::vector<int32_t> off(F + 1, 0);
    for (int f = 0; f < F; ++f) off[f + 1] = off[f] + lens[f];
    std::vector<float> u(off[F]), v(off[F]);
    srand(2);
    for (int i = 0; i < off[F]; ++i) { u[i] = 1241.0f * rand() / RAND_MAX; v[i] = 376.0f * rand() / RAND_MAX; }
    u[off[3] + 5] = -2.0f;
    std::vector<int32_t> idx(off[F]), n_out(F); std::vector<uint8_t> st(F);
    int rc = bucket_emu(F, off.data(), u.data(), v.data(), 30, 2, 9, nullptr, 0, idx.data(), n_out.data(), st.data(), 2);
    for (int f = 0; f < F; ++f) printf("frame %d n %d kept %d status %d\n", f, lens[f], n_out[f], st[f]);
    return rc;
}
