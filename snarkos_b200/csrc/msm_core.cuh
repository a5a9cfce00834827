// msm_core.cuh -- host/device-shared pieces of the Pippenger MSM over BLS12-377 G1.
//
// GPU counterpart of snarkVM VariableBase::msm -> batched::msm / standard::msm
// [UPSTREAM snarkvm-algorithms 1.0.0 @ dea322b: algorithms/src/msm/variable_base/{mod,batched,standard}.rs;
//  SURVEY.md 8a rows a1-a2].  Same group element, different schedule:
//     snarkVM : c = ln(n) + 2 unsigned windows, rayon over windows, affine batch-add or Jacobian buckets
//     here    : signed c-bit digits (half the buckets), counting sort of (window,bucket) -> point lists,
//               one thread per bucket accumulating in XYZZ, segmented running-sum reduction, Horner fold.
#pragma once
#include "ec.cuh"

#define MSM_SCALAR_BITS 253

struct MsmShape {
    uint32_t c;          // window width in bits
    uint32_t nwin;       // floor(253 / c) + 1: always room for the last signed-digit carry
    uint32_t nbuckets;   // per window: 2^(c-1), bucket b holds digit magnitude b + 1
};
B200_HOSTDEV MsmShape msm_shape(uint32_t c, uint32_t bits = MSM_SCALAR_BITS) {
    MsmShape s;
    s.c = c;
    s.nwin = bits / c + 1;
    s.nbuckets = 1u << (c - 1);
    return s;
}

// Scalars cross the ABI canonical (< r), which is what snarkVM's to_bigint() produces.  A non-canonical 256-bit k >= r
// still denotes (k mod r) * P on a point of order r, and both the window count (253 bits) and the GLV split (127-bit
// halves) assume k < r -- so k is reduced here instead of being silently truncated: at most 13 subtractions
// (2^256 / r < 14), none for canonical input (one compare of the top limb).
B200_HD void msm_scalar_reduce(uint32_t* s) {
    if (s[7] < FrP::mod(7)) return;
    for (int it = 0; it < 14; it++) {
        uint32_t d[8];
        uint32_t borrow = 0;
        for (int i = 0; i < 8; i++) {
            const unsigned long long t = (unsigned long long)s[i] - FrP::mod(i) - borrow;
            d[i] = (uint32_t)t;
            borrow = (uint32_t)(t >> 32) & 1u;
        }
        if (borrow) return;                            // s < r
        for (int i = 0; i < 8; i++) s[i] = d[i];
    }
}

// bits [lo, lo + c) of a 256-bit little-endian scalar held as 8 x u32 (c <= 24)
B200_HD uint32_t msm_window_bits(const uint32_t* s, uint32_t lo, uint32_t c) {
    uint32_t limb = lo >> 5, off = lo & 31;
    if (limb >= 8) return 0;
    uint64_t v = s[limb];
    if (limb + 1 < 8) v |= (uint64_t)s[limb + 1] << 32;
    return (uint32_t)(v >> off) & ((1u << c) - 1);
}

// Signed digit of window w given the carry from window w-1.  Returns magnitude in [0, 2^(c-1)] and the
// sign; updates carry.   sum_w digit_w * 2^(c*w) == scalar.
B200_HD uint32_t msm_signed_digit(const uint32_t* s, uint32_t w, uint32_t c, uint32_t& carry, uint32_t& neg) {
    uint32_t raw = msm_window_bits(s, w * c, c) + carry;
    uint32_t half = 1u << (c - 1);
    if (raw > half) {
        carry = 1;
        neg = 1;
        return (1u << c) - raw;
    }
    carry = 0;
    neg = 0;
    return raw;
}

// 96-byte packed affine point used on the device: x | y (Montgomery), infinity = (0, 0)
struct g1_packed_t { uint4 w[6]; };
// The packed BASES (gathered at random by index) sit in 128-byte records, x | y | 32 B pad: an L2 miss on B200 fetches
// the whole 128-byte line from HBM (ncu: 165 B per 48-byte and 202 B per 96-byte gather from 96-byte records), so a
// record that never straddles a line costs one line per gather.  Point LISTS (read in order) are two planes of
// 48-byte coordinates (x plane, y plane).
#define G1_BASE_U4 8u
#define G1_BASE_BYTES (16u * G1_BASE_U4)
B200_HD g1_packed_t g1_load_packed(const uint4* p) {
    g1_packed_t r;
    B200_UNROLL
    for (int k = 0; k < 6; k++) r.w[k] = p[k];
    return r;
}
// affine point of a 128-byte base record with three 256-bit loads (x | y = bytes 0..95 of a 32-byte aligned record)
B200_HD g1_affine_t g1_load_base(const uint4* rec) {
    uint32_t w[24];
    ptx::ld_global_256(rec, w);
    ptx::ld_global_256(rec + 2, w + 8);
    ptx::ld_global_256(rec + 4, w + 16);
    g1_affine_t a;
    B200_UNROLL
    for (int k = 0; k < 12; k++) { a.x.v[k] = w[k]; a.y.v[k] = w[12 + k]; }
    return a;
}
// x coordinate only: one 256-bit and one 128-bit load
B200_HD fq_t g1_load_base_x(const uint4* rec) {
    uint32_t w[8];
    ptx::ld_global_256(rec, w);
    const uint4 t = rec[2];
    fq_t x;
    B200_UNROLL
    for (int k = 0; k < 8; k++) x.v[k] = w[k];
    x.v[8] = t.x; x.v[9] = t.y; x.v[10] = t.z; x.v[11] = t.w;
    return x;
}
B200_HD g1_packed_t g1_load_planes(const uint4* x, const uint4* y) {
    g1_packed_t r;
    B200_UNROLL
    for (int k = 0; k < 3; k++) { r.w[k] = x[k]; r.w[3 + k] = y[k]; }
    return r;
}
B200_HD void g1_store_packed(uint4* p, const g1_packed_t& v) {
    B200_UNROLL
    for (int k = 0; k < 6; k++) p[k] = v.w[k];
}

B200_HD g1_affine_t g1_unpack(const g1_packed_t& p) {
    g1_affine_t a;
    a.x.v[0] = p.w[0].x; a.x.v[1] = p.w[0].y; a.x.v[2] = p.w[0].z; a.x.v[3] = p.w[0].w;
    a.x.v[4] = p.w[1].x; a.x.v[5] = p.w[1].y; a.x.v[6] = p.w[1].z; a.x.v[7] = p.w[1].w;
    a.x.v[8] = p.w[2].x; a.x.v[9] = p.w[2].y; a.x.v[10] = p.w[2].z; a.x.v[11] = p.w[2].w;
    a.y.v[0] = p.w[3].x; a.y.v[1] = p.w[3].y; a.y.v[2] = p.w[3].z; a.y.v[3] = p.w[3].w;
    a.y.v[4] = p.w[4].x; a.y.v[5] = p.w[4].y; a.y.v[6] = p.w[4].z; a.y.v[7] = p.w[4].w;
    a.y.v[8] = p.w[5].x; a.y.v[9] = p.w[5].y; a.y.v[10] = p.w[5].z; a.y.v[11] = p.w[5].w;
    return a;
}

// XYZZ bucket image in global memory: 12 x uint4 (X | Y | ZZ | ZZZ)
struct g1_xyzz_mem_t { uint4 w[12]; };

B200_HD void fq_to_u4x3(const fq_t& a, uint4* w) {
    w[0].x = a.v[0]; w[0].y = a.v[1]; w[0].z = a.v[2]; w[0].w = a.v[3];
    w[1].x = a.v[4]; w[1].y = a.v[5]; w[1].z = a.v[6]; w[1].w = a.v[7];
    w[2].x = a.v[8]; w[2].y = a.v[9]; w[2].z = a.v[10]; w[2].w = a.v[11];
}
B200_HD fq_t fq_from_u4x3(const uint4* w) {
    fq_t a;
    a.v[0] = w[0].x; a.v[1] = w[0].y; a.v[2] = w[0].z; a.v[3] = w[0].w;
    a.v[4] = w[1].x; a.v[5] = w[1].y; a.v[6] = w[1].z; a.v[7] = w[1].w;
    a.v[8] = w[2].x; a.v[9] = w[2].y; a.v[10] = w[2].z; a.v[11] = w[2].w;
    return a;
}
B200_HD void g1_xyzz_store(g1_xyzz_mem_t* dst, const g1_xyzz_t& p) {
    fq_to_u4x3(p.X, dst->w + 0);
    fq_to_u4x3(p.Y, dst->w + 3);
    fq_to_u4x3(p.ZZ, dst->w + 6);
    fq_to_u4x3(p.ZZZ, dst->w + 9);
}
B200_HD g1_xyzz_t g1_xyzz_load(const g1_xyzz_mem_t* src) {
    g1_xyzz_t p;
    p.X = fq_from_u4x3(src->w + 0);
    p.Y = fq_from_u4x3(src->w + 3);
    p.ZZ = fq_from_u4x3(src->w + 6);
    p.ZZZ = fq_from_u4x3(src->w + 9);
    return p;
}

// Segment of the running-sum bucket reduction: buckets [s0, s0 + len) of one window carry the weights
// s0+1 .. s0+len.  Returns  sum_i (s0 + 1 + i) * B[s0 + i]  =  (sum_i (i+1) B[s0+i])  +  s0 * (sum_i B[s0+i]).
B200_HD g1_xyzz_t msm_reduce_segment(const g1_xyzz_mem_t* buckets, uint32_t s0, uint32_t len) {
    g1_xyzz_t running = g1_xyzz_infinity();
    g1_xyzz_t acc = g1_xyzz_infinity();
    for (uint32_t i = len; i-- > 0;) {
        g1_xyzz_t b = g1_xyzz_load(buckets + s0 + i);
        g1_add(running, b);
        g1_add(acc, running);
    }
    if (s0) {
        g1_xyzz_t t = g1_mul_u64(running, s0);
        g1_add(acc, t);
    }
    return acc;
}

// Horner fold of the per-window sums, high window first: total = sum_w 2^(c*w) * W_w
B200_HD g1_xyzz_t msm_fold_windows(const g1_xyzz_mem_t* wsum, uint32_t nwin, uint32_t c) {
    g1_xyzz_t total = g1_xyzz_infinity();
    for (uint32_t w = nwin; w-- > 0;) {
        g1_dbl_k(total, c);
        g1_xyzz_t s = g1_xyzz_load(wsum + w);
        g1_add(total, s);
    }
    return total;
}
