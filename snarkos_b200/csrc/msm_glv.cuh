// msm_glv.cuh -- GLV decomposition for BLS12-377 G1: every 253-bit scalar becomes two 127-bit ones.
//
// G1 has the endomorphism phi(x, y) = (beta x, y) = lambda * (x, y), beta a primitive cube root of unity in Fq and
// lambda = u^2 - 1 (u = 0x8508c00000000001 the curve parameter; lambda^2 + lambda + 1 = u^4 - u^2 + 1 = r).
// Since lambda ~ 2^126.1 ~ sqrt(r), plain division splits a scalar: k = k1 + k2 * lambda with k1 = k mod lambda,
// k2 = k div lambda, both in [0, 2^127), so  k P = k1 P + k2 phi(P)  -- an MSM over 2n points with half-length scalars:
// the same number of bucket additions, but half the windows: half the buckets to reduce and half the dependent
// doublings of the window fold (what a small call mostly waits for), and one window less at 2^24 (7 x 19 bits
// instead of 15 x 18: 14 n instead of 15 n additions).
// The constants are derived and checked in tools/gen_glv_constants.py (beta from lambda * G on the oracle).
#pragma once
#include "msm_core.cuh"

#define MSM_GLV_BITS 127

// lambda = 0x452217cc900000010a11800000000000
B200_HOSTDEV uint32_t msm_glv_lambda(int i) {
    const uint32_t v[4] = {0x00000000u, 0x0a118000u, 0x90000001u, 0x452217ccu};
    return v[i];
}
// m = floor(2^254 / lambda) (128 bits)
B200_HOSTDEV uint32_t msm_glv_recip(int i) {
    const uint32_t v[4] = {0xabe4060bu, 0x1fdcbb4cu, 0xa7f4dc58u, 0xecfdeaa5u};
    return v[i];
}
// beta in Montgomery form
B200_HD fq_t msm_glv_beta() {
    const uint32_t v[12] = {0xa5847973u, 0xdacd106du, 0xbac2a79au, 0xd8fe2454u, 0xfd832edcu, 0x1ada4fd6u,
                            0x9d150908u, 0xfb986844u, 0xea32285eu, 0xd63eb8aeu, 0x6f873fd0u, 0x0167d6a3u};
    fq_t b;
    B200_UNROLL
    for (int i = 0; i < 12; i++) b.v[i] = v[i];
    return b;
}

// k (8 x u32, < r) -> k1 = k mod lambda, k2 = k div lambda (4 x u32 each, < 2^127)
B200_HD void msm_glv_split(const uint32_t* k, uint32_t* k1, uint32_t* k2) {
    // q = floor(k * m / 2^254): at most 1 below the true quotient (checked on 2 * 10^5 random scalars and the edges)
    uint32_t prod[13];
    for (int i = 0; i < 13; i++) prod[i] = 0;
    for (int i = 0; i < 8; i++) {
        unsigned long long carry = 0;
        for (int j = 0; j < 4; j++) {
            const unsigned long long t = (unsigned long long)k[i] * msm_glv_recip(j) + prod[i + j] + carry;
            prod[i + j] = (uint32_t)t;
            carry = t >> 32;
        }
        prod[i + 4] = (uint32_t)carry;
    }
    uint32_t q[4];
    for (int j = 0; j < 4; j++) q[j] = (prod[7 + j] >> 30) | (prod[8 + j] << 2);
    // rem = k - q * lambda, exact in 128 bits (0 <= rem < 3 lambda < 2^129 would not fit: the quotient is off by at
    // most one, so rem < 2 lambda < 2^128)
    uint32_t ql[4] = {0, 0, 0, 0};
    for (int i = 0; i < 4; i++) {
        unsigned long long carry = 0;
        for (int j = 0; i + j < 4; j++) {
            const unsigned long long t = (unsigned long long)q[i] * msm_glv_lambda(j) + ql[i + j] + carry;
            ql[i + j] = (uint32_t)t;
            carry = t >> 32;
        }
    }
    uint32_t rem[4];
    unsigned long long borrow = 0;
    for (int i = 0; i < 4; i++) {
        const unsigned long long t = (unsigned long long)k[i] - ql[i] - borrow;
        rem[i] = (uint32_t)t;
        borrow = (t >> 32) & 1u;
    }
    for (int fix = 0; fix < 2; fix++) {
        // rem >= lambda ?
        uint32_t d[4];
        unsigned long long b = 0;
        for (int i = 0; i < 4; i++) {
            const unsigned long long t = (unsigned long long)rem[i] - msm_glv_lambda(i) - b;
            d[i] = (uint32_t)t;
            b = (t >> 32) & 1u;
        }
        if (b) break;                                  // rem < lambda
        for (int i = 0; i < 4; i++) rem[i] = d[i];
        unsigned long long c = 1;
        for (int i = 0; i < 4; i++) {
            const unsigned long long t = (unsigned long long)q[i] + c;
            q[i] = (uint32_t)t;
            c = t >> 32;
        }
    }
    for (int i = 0; i < 4; i++) { k1[i] = rem[i]; k2[i] = q[i]; }
}
