// AddressSanitizer / UBSan harness of the host packers of libncb (pack.cu: dense tensor -> CSR, planar
// rows -> CSR with 2 or 3 variables): random inputs, every output checked against a scalar restatement.
//
//   g++ -x c++ -std=c++17 -O1 -g -fsanitize=address,undefined -I include -I nanocompore_b200/csrc \
//       nanocompore_b200/csrc/pack.cu tests/csrc/asan_pack_harness.cpp -lpthread -o build/asan/pack && ./build/asan/pack
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <random>
#include <vector>
#include "ncb.h"

static bool num(float x) { return x == x; }

int main() {
    std::mt19937 g(7);
    std::uniform_real_distribution<float> u(0.f, 1.f);
    int checked = 0;
    for (int rep = 0; rep < 60; ++rep) {
        const int64_t P = (int64_t)(g() % 1400), R = (int64_t)(g() % 40);
        const int motor = rep % 3 == 0 ? (int)(g() % 9) : 0, threads = 1 + (int)(g() % 5), V = motor > 0 ? 3 : 2;
        std::vector<std::vector<float>> I(R, std::vector<float>(P)), D(R, std::vector<float>(P));
        std::vector<const float *> ip(R), dp(R);
        std::vector<uint8_t> lab(R);
        for (int64_t r = 0; r < R; ++r) {
            for (int64_t p = 0; p < P; ++p) {
                I[r][p] = u(g) < 0.4f ? nanf("") : u(g);
                D[r][p] = u(g) < 0.4f ? nanf("") : u(g);
            }
            ip[r] = I[r].data(); dp[r] = D[r].data(); lab[r] = (uint8_t)(g() % 8);
        }
        std::vector<int64_t> counts(P + 1, -1), off(P + 1, 0);
        if (ncb_pack_rows_count(ip.data(), dp.data(), R, P, motor, counts.data(), threads) != NCB_OK) return 1;
        for (int64_t p = 0; p < P; ++p) off[p + 1] = off[p] + counts[p];
        std::vector<float> sig((size_t)off[P] * V + 1, -7.f);
        std::vector<uint8_t> lo((size_t)off[P] + 1, 255);
        if (ncb_pack_rows_fill(ip.data(), dp.data(), lab.data(), R, P, motor, off.data(), sig.data(), lo.data(), threads) != NCB_OK) return 1;
        for (int64_t p = 0; p < P; ++p) {
            int64_t o = off[p];
            for (int64_t r = 0; r < R; ++r) {
                const float m = (motor > 0 && p + motor < P) ? D[r][p + motor] : nanf("");
                if (!(num(I[r][p]) || num(D[r][p]) || num(m))) continue;
                const float *s = &sig[(size_t)o * V];
                auto same = [](float a, float b) { return (a != a && b != b) || a == b; };
                if (!same(s[0], I[r][p]) || !same(s[1], D[r][p]) || (V == 3 && !same(s[2], m)) || lo[o] != lab[r]) {
                    printf("mismatch rep %d p %ld r %ld\n", rep, (long)p, (long)r);
                    return 2;
                }
                ++o;
            }
            if (o != off[p + 1]) { printf("count mismatch rep %d p %ld\n", rep, (long)p); return 3; }
            ++checked;
        }
        // dense packer on the same data
        if (R > 0 && P > 0) {
            std::vector<float> dense((size_t)P * R * 2);
            for (int64_t p = 0; p < P; ++p)
                for (int64_t r = 0; r < R; ++r) { dense[(p * R + r) * 2] = I[r][p]; dense[(p * R + r) * 2 + 1] = D[r][p]; }
            std::vector<int64_t> c2(P), o2(P + 1, 0);
            if (ncb_pack_count(dense.data(), P, (int32_t)R, 2, c2.data(), threads) != NCB_OK) return 1;
            for (int64_t p = 0; p < P; ++p) o2[p + 1] = o2[p] + c2[p];
            std::vector<float> s2((size_t)o2[P] * 2 + 1);
            std::vector<uint8_t> l2((size_t)o2[P] + 1);
            if (ncb_pack_fill(dense.data(), lab.data(), P, (int32_t)R, 2, o2.data(), s2.data(), l2.data(), threads) != NCB_OK) return 1;
            if (motor == 0 && (o2[P] != off[P])) { printf("dense / rows disagree rep %d\n", rep); return 4; }
        }
    }
    printf("ok: %d positions checked\n", checked);
    return 0;
}
