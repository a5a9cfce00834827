// ec_coop.cuh -- the XYZZ group law with FOUR LANES PER POINT, for the latency-bound tails of an MSM.
//
// The window fold is one dependent chain per MSM: 127 (GLV) or 253 doublings and one addition per window, 9 and 14 Fq
// products each, executed by ONE thread -- ~1.3 us per product on a lone warp (every IMAD.WIDE of the carry chain
// waits for the previous one), 1.7 ms for the 256 small MSMs of a block's verification, 1.0 ms at 2^24.  The other 31
// lanes of the warp idle.  SIMT makes different lanes running DIFFERENT products free, so here a quad of lanes holds
// the same point (replicated), every lane computes one of the independent products of a level of the formula's
// dependency graph, and the quad exchanges the results with shuffles:
//     doubling  9 products in 3 levels     addition  14 products in 4 levels
// Same formulas and the same values as ec.cuh (dbl-2008-s-1, add-2008-s); the exceptional cases (operand at infinity,
// equal points, opposite points) are quad-uniform branches, shuffles use the quad's own mask.
#pragma once
#include "ec.cuh"

struct Quad {
    uint32_t mask, base, q;
};
__device__ __forceinline__ Quad quad_here() {
    const uint32_t lane = threadIdx.x & 31u;
    Quad Q;
    Q.base = lane & ~3u;
    Q.q = lane & 3u;
    Q.mask = 0xFu << Q.base;
    return Q;
}
__device__ __forceinline__ fq_t quad_pick(const Quad& Q, const fq_t& a0, const fq_t& a1, const fq_t& a2, const fq_t& a3) {
    fq_t r;
#pragma unroll
    for (int i = 0; i < 12; i++) {
        const uint32_t lo = (Q.q & 1u) ? a1.v[i] : a0.v[i];
        const uint32_t hi = (Q.q & 1u) ? a3.v[i] : a2.v[i];
        r.v[i] = (Q.q & 2u) ? hi : lo;
    }
    return r;
}
// r_j = a_j * b_j for j < 4, lane j of the quad computing product j; every lane receives all four results
__device__ __forceinline__ void quad_mul4(const Quad& Q, const fq_t& a0, const fq_t& b0, const fq_t& a1, const fq_t& b1,
                                          const fq_t& a2, const fq_t& b2, const fq_t& a3, const fq_t& b3, fq_t& r0, fq_t& r1,
                                          fq_t& r2, fq_t& r3) {
    const fq_t x = quad_pick(Q, a0, a1, a2, a3), y = quad_pick(Q, b0, b1, b2, b3);
    const fq_t r = fp_mul(x, y);
#pragma unroll
    for (int i = 0; i < 12; i++) {
        r0.v[i] = __shfl_sync(Q.mask, r.v[i], Q.base + 0);
        r1.v[i] = __shfl_sync(Q.mask, r.v[i], Q.base + 1);
        r2.v[i] = __shfl_sync(Q.mask, r.v[i], Q.base + 2);
        r3.v[i] = __shfl_sync(Q.mask, r.v[i], Q.base + 3);
    }
}

// p <- 2p, all four lanes hold p on entry and on exit
__device__ __forceinline__ void g1_dbl_quad(const Quad& Q, g1_xyzz_t& p) {
    if (g1_xyzz_is_infinity(p)) return;                       // quad-uniform
    const fq_t U = fp_dbl(p.Y);
    fq_t V, XX, d0, d1;
    quad_mul4(Q, U, U, p.X, p.X, U, U, p.X, p.X, V, XX, d0, d1);
    const fq_t M = fp_add(fp_dbl(XX), XX);
    fq_t W, S, MM, ZZ3;
    quad_mul4(Q, U, V, p.X, V, M, M, V, p.ZZ, W, S, MM, ZZ3);
    const fq_t X3 = fp_sub(MM, fp_dbl(S));
    fq_t T1, T2, ZZZ3;
    quad_mul4(Q, M, fp_sub(S, X3), W, p.Y, W, p.ZZZ, W, p.ZZZ, T1, T2, ZZZ3, d0);
    p.X = X3;
    p.Y = fp_sub(T1, T2);
    p.ZZ = ZZ3;
    p.ZZZ = ZZZ3;
}

// acc <- acc + q
__device__ __forceinline__ void g1_add_quad(const Quad& Q, g1_xyzz_t& acc, const g1_xyzz_t& q) {
    if (g1_xyzz_is_infinity(q)) return;
    if (g1_xyzz_is_infinity(acc)) { acc = q; return; }
    fq_t U1, U2, S1, S2;
    quad_mul4(Q, acc.X, q.ZZ, q.X, acc.ZZ, acc.Y, q.ZZZ, q.Y, acc.ZZZ, U1, U2, S1, S2);
    const fq_t Pp = fp_sub(U2, U1), R = fp_sub(S2, S1);
    if (fp_is_zero(Pp)) {
        if (fp_is_zero(R)) g1_dbl_quad(Q, acc);
        else acc = g1_xyzz_infinity();
        return;
    }
    fq_t PP, RR, ZZa, ZZZa;
    quad_mul4(Q, Pp, Pp, R, R, acc.ZZ, q.ZZ, acc.ZZZ, q.ZZZ, PP, RR, ZZa, ZZZa);
    fq_t PPP, Qq, ZZ3, d0;
    quad_mul4(Q, Pp, PP, U1, PP, ZZa, PP, ZZa, PP, PPP, Qq, ZZ3, d0);
    const fq_t X3 = fp_sub(fp_sub(RR, PPP), fp_dbl(Qq));
    fq_t T1, T2, ZZZ3;
    quad_mul4(Q, R, fp_sub(Qq, X3), S1, PPP, ZZZa, PPP, ZZZa, PPP, T1, T2, ZZZ3, d0);
    acc.X = X3;
    acc.Y = fp_sub(T1, T2);
    acc.ZZ = ZZ3;
    acc.ZZZ = ZZZ3;
}
