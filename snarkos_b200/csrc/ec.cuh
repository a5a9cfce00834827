// ec.cuh -- BLS12-377 G1 (y^2 = x^3 + 1 over Fq) group law in XYZZ coordinates.
//
// GPU counterpart of snarkVM's short_weierstrass_jacobian::{Affine, Projective} add_assign_mixed /
// double_in_place / add_assign  [UPSTREAM snarkvm-curves 1.0.0 @ dea322b:
// curves/src/templates/short_weierstrass_jacobian/{affine,projective}.rs; SURVEY.md 8a row a8].
// snarkVM accumulates in Jacobian (and in affine with batch inversion inside batched::msm); the group
// element computed is the same, only the internal representation differs, and results cross the ABI as
// Jacobian (X, Y, Z) via g1_xyzz_to_jacobian -- compared after affine normalisation.
//
// XYZZ: (X, Y, ZZ, ZZZ) with x = X/ZZ, y = Y/ZZZ, ZZ^3 = ZZZ^2; infinity <=> ZZ = 0.
// The textbook formulas are incomplete (P + P gives ZZ = 0), so every entry point branches on the
// exceptional cases: operand at infinity, equal points (-> double), opposite points (-> infinity).
#pragma once
#include "field.cuh"

struct g1_affine_t { fq_t x, y; };           // (0, 0) encodes infinity (not on the curve since b = 1)
struct g1_xyzz_t { fq_t X, Y, ZZ, ZZZ; };

B200_HD g1_affine_t g1_affine_infinity() {
    g1_affine_t a;
    a.x = fp_zero<FqP>();
    a.y = fp_zero<FqP>();
    return a;
}
B200_HD bool g1_affine_is_infinity(const g1_affine_t& a) { return fp_is_zero(a.x) && fp_is_zero(a.y); }
B200_HD g1_affine_t g1_affine_neg(const g1_affine_t& a) {
    g1_affine_t r;
    r.x = a.x;
    r.y = fp_neg(a.y);
    return r;
}
B200_HD g1_xyzz_t g1_xyzz_infinity() {
    g1_xyzz_t p;
    p.X = fp_one<FqP>();
    p.Y = fp_one<FqP>();
    p.ZZ = fp_zero<FqP>();
    p.ZZZ = fp_zero<FqP>();
    return p;
}
B200_HD bool g1_xyzz_is_infinity(const g1_xyzz_t& p) { return fp_is_zero(p.ZZ); }
B200_HD g1_xyzz_t g1_xyzz_from_affine(const g1_affine_t& a) {
    if (g1_affine_is_infinity(a)) return g1_xyzz_infinity();
    g1_xyzz_t p;
    p.X = a.x;
    p.Y = a.y;
    p.ZZ = fp_one<FqP>();
    p.ZZZ = fp_one<FqP>();
    return p;
}

// p <- 2p            (dbl-2008-s-1 for a = 0: 6M + 4S... written as 2S + 7M here, a = 0)
B200_HD void g1_dbl(g1_xyzz_t& p) {
    if (g1_xyzz_is_infinity(p)) return;
    fq_t U = fp_dbl(p.Y);
    fq_t V = fp_sqr(U);
    fq_t W = fp_mul(U, V);
    fq_t S = fp_mul(p.X, V);
    fq_t XX = fp_sqr(p.X);
    fq_t M = fp_add(fp_dbl(XX), XX);
    fq_t X3 = fp_sub(fp_sqr(M), fp_dbl(S));
    fq_t Y3 = fp_sub(fp_mul(M, fp_sub(S, X3)), fp_mul(W, p.Y));
    p.ZZ = fp_mul(V, p.ZZ);
    p.ZZZ = fp_mul(W, p.ZZZ);
    p.X = X3;
    p.Y = Y3;
}

// affine doubling straight into XYZZ (ZZ = ZZZ = 1 on input)
B200_HD g1_xyzz_t g1_dbl_affine(const g1_affine_t& a) {
    g1_xyzz_t r;
    fq_t U = fp_dbl(a.y);
    fq_t V = fp_sqr(U);
    fq_t W = fp_mul(U, V);
    fq_t S = fp_mul(a.x, V);
    fq_t XX = fp_sqr(a.x);
    fq_t M = fp_add(fp_dbl(XX), XX);
    r.X = fp_sub(fp_sqr(M), fp_dbl(S));
    r.Y = fp_sub(fp_mul(M, fp_sub(S, r.X)), fp_mul(W, a.y));
    r.ZZ = V;
    r.ZZZ = W;
    return r;
}

// acc <- acc + a   (mixed addition, madd-2008-s: 8M + 2S)
B200_HD void g1_madd(g1_xyzz_t& acc, const g1_affine_t& a) {
    if (g1_affine_is_infinity(a)) return;
    if (g1_xyzz_is_infinity(acc)) {
        acc.X = a.x;
        acc.Y = a.y;
        acc.ZZ = fp_one<FqP>();
        acc.ZZZ = fp_one<FqP>();
        return;
    }
    fq_t U2 = fp_mul(a.x, acc.ZZ);
    fq_t S2 = fp_mul(a.y, acc.ZZZ);
    fq_t Pp = fp_sub(U2, acc.X);
    fq_t R = fp_sub(S2, acc.Y);
    if (fp_is_zero(Pp)) {
        if (fp_is_zero(R)) acc = g1_dbl_affine(a);
        else acc = g1_xyzz_infinity();
        return;
    }
    fq_t PP = fp_sqr(Pp);
    fq_t PPP = fp_mul(Pp, PP);
    fq_t Q = fp_mul(acc.X, PP);
    fq_t X3 = fp_sub(fp_sub(fp_sqr(R), PPP), fp_dbl(Q));
    fq_t Y3 = fp_sub(fp_mul(R, fp_sub(Q, X3)), fp_mul(acc.Y, PPP));
    acc.ZZ = fp_mul(acc.ZZ, PP);
    acc.ZZZ = fp_mul(acc.ZZZ, PPP);
    acc.X = X3;
    acc.Y = Y3;
}

// acc <- acc + q   (add-2008-s: 12M + 2S)
B200_HD void g1_add(g1_xyzz_t& acc, const g1_xyzz_t& q) {
    if (g1_xyzz_is_infinity(q)) return;
    if (g1_xyzz_is_infinity(acc)) { acc = q; return; }
    fq_t U1 = fp_mul(acc.X, q.ZZ);
    fq_t U2 = fp_mul(q.X, acc.ZZ);
    fq_t S1 = fp_mul(acc.Y, q.ZZZ);
    fq_t S2 = fp_mul(q.Y, acc.ZZZ);
    fq_t Pp = fp_sub(U2, U1);
    fq_t R = fp_sub(S2, S1);
    if (fp_is_zero(Pp)) {
        if (fp_is_zero(R)) g1_dbl(acc);
        else acc = g1_xyzz_infinity();
        return;
    }
    fq_t PP = fp_sqr(Pp);
    fq_t PPP = fp_mul(Pp, PP);
    fq_t Q = fp_mul(U1, PP);
    fq_t X3 = fp_sub(fp_sub(fp_sqr(R), PPP), fp_dbl(Q));
    fq_t Y3 = fp_sub(fp_mul(R, fp_sub(Q, X3)), fp_mul(S1, PPP));
    acc.ZZ = fp_mul(fp_mul(acc.ZZ, q.ZZ), PP);
    acc.ZZZ = fp_mul(fp_mul(acc.ZZZ, q.ZZZ), PPP);
    acc.X = X3;
    acc.Y = Y3;
}

// k * p, left-to-right double-and-add on a 64-bit scalar
B200_HD g1_xyzz_t g1_mul_u64(const g1_xyzz_t& p, uint64_t k) {
    g1_xyzz_t acc = g1_xyzz_infinity();
    bool started = false;
    for (int i = 63; i >= 0; i--) {
        if (started) g1_dbl(acc);
        if ((k >> i) & 1) {
            if (started) g1_add(acc, p);
            else { acc = p; started = true; }
        }
    }
    return acc;
}

// XYZZ -> Jacobian without inversion: Z = ZZZ, Xj = X * ZZ^2, Yj = Y * ZZZ^2   
// infinity -> (1, 1, 0) in Montgomery form, snarkVM's Projective::zero() convention.
B200_HD void g1_xyzz_to_jacobian(const g1_xyzz_t& p, fq_t& X, fq_t& Y, fq_t& Z) {
    if (g1_xyzz_is_infinity(p)) {
        X = fp_one<FqP>();
        Y = fp_one<FqP>();
        Z = fp_zero<FqP>();
        return;
    }
    X = fp_mul(p.X, fp_sqr(p.ZZ));
    Y = fp_mul(p.Y, fp_sqr(p.ZZZ));
    Z = p.ZZZ;
}

// Jacobian doubling for a = 0 (dbl-2009-l: 2M + 5S = 7 products against the 9 of the XYZZ doubling), used where a
// long run of doublings sits on ONE thread's critical path (the window fold: 253 - c dependent doublings).
// Z = 0 (infinity) maps to Z = 0.
B200_HD void g1_jac_dbl(fq_t& X, fq_t& Y, fq_t& Z) {
    const fq_t A = fp_sqr(X);
    const fq_t B = fp_sqr(Y);
    const fq_t C = fp_sqr(B);
    const fq_t XB = fp_add(X, B);
    const fq_t D = fp_dbl(fp_sub(fp_sub(fp_sqr(XB), A), C));
    const fq_t E = fp_add(fp_dbl(A), A);
    const fq_t F = fp_sqr(E);
    const fq_t Z3 = fp_dbl(fp_mul(Y, Z));
    X = fp_sub(F, fp_dbl(D));
    const fq_t C8 = fp_dbl(fp_dbl(fp_dbl(C)));
    Y = fp_sub(fp_mul(E, fp_sub(D, X)), C8);
    Z = Z3;
}
// p <- 2^k p through Jacobian coordinates (4 products in, 2 out, 7 per doubling)
B200_HD void g1_dbl_k(g1_xyzz_t& p, uint32_t k) {
    if (g1_xyzz_is_infinity(p) || k == 0) return;
    fq_t X, Y, Z;
    g1_xyzz_to_jacobian(p, X, Y, Z);
    for (uint32_t i = 0; i < k; i++) g1_jac_dbl(X, Y, Z);
    if (fp_is_zero(Z)) { p = g1_xyzz_infinity(); return; }       // only for a point of order 2^j (y = 0 on the way)
    p.X = X;
    p.Y = Y;
    p.ZZ = fp_sqr(Z);
    p.ZZZ = fp_mul(p.ZZ, Z);
}

// XYZZ -> affine (one Fermat inversion; utility paths only)
B200_HD g1_affine_t g1_xyzz_to_affine(const g1_xyzz_t& p) {
    if (g1_xyzz_is_infinity(p)) return g1_affine_infinity();
    // 1/ZZZ = i3; 1/ZZ = i3^2 * ZZ^2 ... since ZZ^3 = ZZZ^2:  1/ZZ = ZZ^2 / ZZZ^2
    fq_t i3 = fp_inv(p.ZZZ);
    fq_t i2 = fp_mul(fp_sqr(i3), fp_sqr(p.ZZ));
    g1_affine_t a;
    a.x = fp_mul(p.X, i2);
    a.y = fp_mul(p.Y, i3);
    return a;
}
