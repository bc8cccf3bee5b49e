// trig_check.cpp -- host cross-check of simplepanorama_b200/csrc/glibc_trig.cuh against the libm of this machine.
// Built and run by tests/test_trig_port.py (g++ -O2 -ffp-contract=off -mfma).  The same header is what the CUDA warp
// kernels include; its operations are individually rounded IEEE operations on both sides, so agreement here for every
// float is agreement on the device.
//
//   trig_check <stride> <pairs>
//     sinf / cosf / atanf: every float whose bit pattern is a multiple of <stride> (1 = all 2^32), bitwise compare
//     atan2f: <pairs> random (y, x) pairs (bit patterns from a 64-bit LCG: all magnitudes, signs, specials) plus a grid
//             of special values; bitwise compare (any NaN == any NaN)
// Exit code 0 and a line "mismatches 0 0 0 0" on success.
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <thread>
#include <vector>
#include <atomic>

#include "../simplepanorama_b200/csrc/glibc_trig.cuh"

static inline uint32_t bits(float f) { uint32_t u; memcpy(&u, &f, 4); return u; }
static inline float fl(uint32_t u) { float f; memcpy(&f, &u, 4); return f; }
static inline bool same(float a, float b) { return bits(a) == bits(b) || (a != a && b != b); }

int main(int argc, char **argv)
{
    const uint64_t stride = argc > 1 ? strtoull(argv[1], nullptr, 10) : 64;
    const uint64_t pairs = argc > 2 ? strtoull(argv[2], nullptr, 10) : 20000000ull;
    unsigned nt = std::thread::hardware_concurrency();
    if (nt == 0) nt = 4;
    std::atomic<uint64_t> bad[4];
    for (auto &b : bad) b = 0;
    uint32_t first_bad[4] = {0, 0, 0, 0};
    std::vector<std::thread> th;
    for (unsigned t = 0; t < nt; ++t)
        th.emplace_back([&, t] {
            volatile float sink;
            for (uint64_t u = (uint64_t)t * stride; u < (1ull << 32); u += (uint64_t)nt * stride) {
                const float x = fl((uint32_t)u);
                if (!same(gtrig::sinf_glibc(x), sinf(x))) { if (!bad[0]++) first_bad[0] = (uint32_t)u; }
                if (!same(gtrig::cosf_glibc(x), cosf(x))) { if (!bad[1]++) first_bad[1] = (uint32_t)u; }
                if (!same(gtrig::atanf_glibc(x), atanf(x))) { if (!bad[2]++) first_bad[2] = (uint32_t)u; }
            }
            uint64_t s = 0x9E3779B97F4A7C15ull * (t + 1);
            for (uint64_t i = t; i < pairs; i += nt) {
                s = s * 6364136223846793005ull + 1442695040888963407ull;
                uint32_t a = (uint32_t)(s >> 32);
                s = s * 6364136223846793005ull + 1442695040888963407ull;
                uint32_t b = (uint32_t)(s >> 32);
                // half of the pairs: magnitudes like the stereographic projector's (|u|, |v| within a few decades)
                if (i & 1) { a = (a & 0x80ffffffu) | ((0x78u + ((a >> 24) & 15u)) << 23); b = (b & 0x80ffffffu) | ((0x78u + ((b >> 24) & 15u)) << 23); }
                const float y = fl(a), x = fl(b);
                if (!same(gtrig::atan2f_glibc(y, x), atan2f(y, x))) { if (!bad[3]++) first_bad[3] = a; }
            }
            (void)sink;
        });
    for (auto &t : th) t.join();
    const float sp[] = {0.f, -0.f, 1.f, -1.f, INFINITY, -INFINITY, NAN, 1e-38f, -1e-38f, 1e38f, -1e38f, 0.5f, 2.f, 1e-30f, 1e30f, 3.f, -7.f};
    for (float y : sp)
        for (float x : sp)
            if (!same(gtrig::atan2f_glibc(y, x), atan2f(y, x))) { if (!bad[3]++) first_bad[3] = bits(y); }
    printf("mismatches %llu %llu %llu %llu\n", (unsigned long long)bad[0], (unsigned long long)bad[1], (unsigned long long)bad[2],
           (unsigned long long)bad[3]);
    for (int k = 0; k < 4; ++k)
        if (bad[k]) printf("first mismatch of %s at bits 0x%08x\n", k == 0 ? "sinf" : k == 1 ? "cosf" : k == 2 ? "atanf" : "atan2f", first_bad[k]);
    return (bad[0] || bad[1] || bad[2] || bad[3]) ? 1 : 0;
}
