// CPU-side self test shim (TEST ONLY): compiles snarkos_b200/csrc/field.cuh and ec.cuh with plain g++,
// where ptx_ops.cuh emulates every PTX carry-chain instruction bit-exactly.  Lets the non-GPU test suite
// verify the exact instruction sequences of the device field / curve arithmetic against the oracle.
// This object is never part of the shipped library and is not a CPU fallback.
#include <cstddef>
#include <cstring>

#include "../../snarkos_b200/csrc/field.cuh"
#include "../../snarkos_b200/csrc/ec.cuh"

template <class P> static Fp<P> ld(const uint32_t* p) { Fp<P> r; memcpy(r.v, p, 4 * P::N); return r; }
template <class P> static void st(uint32_t* p, const Fp<P>& a) { memcpy(p, a.v, 4 * P::N); }

template <class P> static Fp<P> mul_via_wide(const Fp<P>& a, const Fp<P>& b) { uint32_t T[2 * P::N]; fp_mul_wide<P>(T, a, b); return fp_redc_wide<P>(T); }
template <class P> static Fp<P> mms(const Fp<P>& a, const Fp<P>& b) { return fp_mul_mul_sub<P>(a, b, b, a); }           // a*b - b*a = 0
template <class P> static Fp<P> mms2(const Fp<P>& a, const Fp<P>& b) { return fp_mul_mul_sub<P>(a, a, b, b); }         // a^2 - b^2

extern "C" {
#define BINOP(NAME, P, FN)                                                                    \
    void NAME(uint32_t* out, const uint32_t* a, const uint32_t* b, size_t n) {                \
        for (size_t i = 0; i < n; i++) st<P>(out + P::N * i, FN(ld<P>(a + P::N * i), ld<P>(b + P::N * i))); \
    }
BINOP(host_fr_mul, FrP, fp_mul<FrP>)
BINOP(host_fr_mul_cc, FrP, fp_mul_cc<FrP>)
BINOP(host_fq_mul_cc, FqP, fp_mul_cc<FqP>)
BINOP(host_fr_add, FrP, fp_add<FrP>)
BINOP(host_fr_sub, FrP, fp_sub<FrP>)
BINOP(host_fq_mul, FqP, fp_mul<FqP>)
BINOP(host_fq_add, FqP, fp_add<FqP>)
BINOP(host_fq_sub, FqP, fp_sub<FqP>)
#define UNOP(NAME, P, FN)                                                                     \
    void NAME(uint32_t* out, const uint32_t* a, size_t n) {                                   \
        for (size_t i = 0; i < n; i++) st<P>(out + P::N * i, FN(ld<P>(a + P::N * i)));          \
    }
BINOP(host_fr_mul_wide, FrP, mul_via_wide<FrP>)
BINOP(host_fq_mul_wide, FqP, mul_via_wide<FqP>)
BINOP(host_fr_mms0, FrP, mms<FrP>)
BINOP(host_fq_mms0, FqP, mms<FqP>)
BINOP(host_fr_mms2, FrP, mms2<FrP>)
BINOP(host_fq_mms2, FqP, mms2<FqP>)
UNOP(host_fr_sqr, FrP, fp_sqr_sos<FrP>)
UNOP(host_fq_sqr, FqP, fp_sqr_sos<FqP>)
UNOP(host_fr_neg, FrP, fp_neg<FrP>)
UNOP(host_fq_neg, FqP, fp_neg<FqP>)
UNOP(host_fr_inv, FrP, fp_inv<FrP>)
UNOP(host_fq_inv, FqP, fp_inv<FqP>)
UNOP(host_fr_inv_gcd, FrP, fp_inv_gcd<FrP>)
UNOP(host_fq_inv_gcd, FqP, fp_inv_gcd<FqP>)
UNOP(host_fr_to_mont, FrP, fp_to_mont<FrP>)
UNOP(host_fr_from_mont, FrP, fp_from_mont<FrP>)

// ---- curve: XYZZ accumulate / add / double / scalar mul, results as Jacobian 144 B
static g1_affine_t ld_aff(const uint8_t* p) {
    g1_affine_t a;
    memcpy(a.x.v, p, 48);
    memcpy(a.y.v, p + 48, 48);
    if (p[96]) a = g1_affine_infinity();
    return a;
}
static void st_jac(uint8_t* out, const g1_xyzz_t& p) {
    fq_t X, Y, Z;
    g1_xyzz_to_jacobian(p, X, Y, Z);
    memcpy(out, X.v, 48); memcpy(out + 48, Y.v, 48); memcpy(out + 96, Z.v, 48);
}
// out = sum_i sign_i * P_i accumulated left to right with g1_madd (signs: 0 = +, 1 = -)
void host_g1_madd_chain(uint8_t* out_jac, const uint8_t* pts, const uint8_t* neg, size_t n, size_t stride) {
    g1_xyzz_t acc = g1_xyzz_infinity();
    for (size_t i = 0; i < n; i++) {
        g1_affine_t a = ld_aff(pts + i * stride);
        if (neg && neg[i]) a = g1_affine_neg(a);
        g1_madd(acc, a);
    }
    st_jac(out_jac, acc);
}
// pairwise tree of g1_add over XYZZ leaves
void host_g1_add_tree(uint8_t* out_jac, const uint8_t* pts, size_t n, size_t stride) {
    g1_xyzz_t* buf = new g1_xyzz_t[n ? n : 1];
    for (size_t i = 0; i < n; i++) buf[i] = g1_xyzz_from_affine(ld_aff(pts + i * stride));
    if (n == 0) buf[0] = g1_xyzz_infinity();
    for (size_t m = n; m > 1; m = (m + 1) / 2) {
        for (size_t i = 0; i < m / 2; i++) { g1_xyzz_t t = buf[2 * i]; g1_add(t, buf[2 * i + 1]); buf[i] = t; }
        if (m & 1) buf[m / 2] = buf[m - 1];
    }
    st_jac(out_jac, buf[0]);
    delete[] buf;
}
void host_g1_mul_u64(uint8_t* out_jac, const uint8_t* pt, uint64_t k) {
    g1_xyzz_t p = g1_xyzz_from_affine(ld_aff(pt));
    st_jac(out_jac, g1_mul_u64(p, k));
}
void host_g1_dbl(uint8_t* out_jac, const uint8_t* pt) {
    g1_xyzz_t p = g1_xyzz_from_affine(ld_aff(pt));
    g1_dbl(p);
    st_jac(out_jac, p);
}
// 2^k * P through the Jacobian doubling run used by the window fold
void host_g1_dbl_k(uint8_t* out_jac, const uint8_t* pt, uint32_t k) {
    g1_xyzz_t p = g1_xyzz_from_affine(ld_aff(pt));
    g1_dbl(p);                       // start from a non-trivial XYZZ representation (ZZ, ZZZ != 1)
    g1_dbl_k(p, k);
    st_jac(out_jac, p);
}
// affine normalisation on the "device" code path: Jacobian-free (x, y) from XYZZ of k * P
void host_g1_mul_u64_affine(uint8_t* out_affine, const uint8_t* pt, uint64_t k, size_t stride) {
    g1_xyzz_t p = g1_mul_u64(g1_xyzz_from_affine(ld_aff(pt)), k);
    g1_affine_t a = g1_xyzz_to_affine(p);
    memset(out_affine, 0, stride);
    if (g1_affine_is_infinity(a)) { out_affine[96] = 1; return; }
    memcpy(out_affine, a.x.v, 48);
    memcpy(out_affine + 48, a.y.v, 48);
}
}
