// rt_xorwow_skip.h — cuRAND XORWOW subsequences without cuRAND's tables: curand_init(seed, subsequence, 0).
//
// The reference's upstream per-pixel seeding is curand_init(1984, pixel_index, 0) (main.cu:90; commented out at HEAD in
// favour of :93 because cuRAND's skip-ahead made render_init cost twice the whole frame, main.cu:91-92).  Subsequence s
// starts 2^67 * s draws into the seed's stream (XORWOW_SEQUENCE_SPACING, curand_precalc.h:54).  The five xorshift
// words advance linearly over GF(2) and the Weyl counter does not move (362437 * 2^67 = 0 mod 2^32,
// curand_kernel.h:697), so the skip is the bit matrix T^(2^67 * s) applied to the words.  cuRAND carries
// T^(2^67 * 4^k) as 218 KB of precalculated constants; here the matrices are DERIVED at first use from the step
// function: T column by column, 67 squarings, then one matrix per BIT of s (powers of T commute, so the product is the
// one cuRAND forms from base-4 digits).  A device kernel then applies them to every pixel in parallel.
#pragma once
#include <stdint.h>
#include <string.h>

#include <vector>

#include "rt_math.cuh"

namespace rt {

constexpr int kSkipBits = 40;            // subsequence numbers below 2^40
constexpr int kSkipRowWords = 5;
constexpr int kSkipRows = 160;
constexpr size_t kSkipMatrixWords = (size_t)kSkipRows * kSkipRowWords;

// out = M applied to in; M stored as the images of the 160 unit vectors (bit j of word i -> row i*32+j)
RT_HD void skip_apply(const uint32_t *M, const uint32_t in[5], uint32_t out[5]) {
    uint32_t r0 = 0, r1 = 0, r2 = 0, r3 = 0, r4 = 0;
    for (int i = 0; i < 5; i++) {
        uint32_t bits = in[i];
        const uint32_t *row = M + (size_t)i * 32 * kSkipRowWords;
        for (int j = 0; j < 32; j++, row += kSkipRowWords)
            if (bits >> j & 1u) { r0 ^= row[0]; r1 ^= row[1]; r2 ^= row[2]; r3 ^= row[3]; r4 ^= row[4]; }
    }
    out[0] = r0; out[1] = r1; out[2] = r2; out[3] = r3; out[4] = r4;
}

// host: the kSkipBits matrices T^(2^(67+k)), k = 0 .. kSkipBits-1
inline std::vector<uint32_t> make_skip_tables() {
    std::vector<uint32_t> T(kSkipMatrixWords), tmp(kSkipMatrixWords), all((size_t)kSkipBits * kSkipMatrixWords);
    for (int i = 0; i < kSkipRows; i++) {           // one draw from each unit vector (curand_kernel.h:569-586)
        xorwow s;
        s.d = 0; s.v0 = s.v1 = s.v2 = s.v3 = s.v4 = 0;
        uint32_t *w[5] = {&s.v0, &s.v1, &s.v2, &s.v3, &s.v4};
        *w[i / 32] = 1u << (i & 31);
        xorwow_next(s);
        for (int k = 0; k < 5; k++) T[(size_t)i * 5 + k] = *w[k];
    }
    auto square = [&](std::vector<uint32_t> &M) {
        for (int i = 0; i < kSkipRows; i++) skip_apply(M.data(), &M[(size_t)i * 5], &tmp[(size_t)i * 5]);
        M.swap(tmp);
    };
    for (int q = 0; q < 67; q++) square(T);
    for (int k = 0; k < kSkipBits; k++) {
        memcpy(&all[(size_t)k * kSkipMatrixWords], T.data(), kSkipMatrixWords * 4);
        square(T);
    }
    return all;
}

// {d, v0..v4} after curand_init(seed, subsequence, 0)
RT_HD void xorwow_seed_subsequence(xorwow &s, unsigned long long seed, unsigned long long subsequence, const uint32_t *tables) {
    xorwow_seed(s, seed);
    uint32_t v[5] = {s.v0, s.v1, s.v2, s.v3, s.v4};
    for (int k = 0; k < kSkipBits && (subsequence >> k); k++)
        if (subsequence >> k & 1ull) skip_apply(tables + (size_t)k * kSkipMatrixWords, v, v);
    s.v0 = v[0]; s.v1 = v[1]; s.v2 = v[2]; s.v3 = v[3]; s.v4 = v[4];
}

}  // namespace rt
