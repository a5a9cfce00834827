/*
 * oracle.c -- CPU restatement (TEST INFRASTRUCTURE ONLY) of the snarkVM hot path that snarkOS
 * reaches through VM::execute / Ledger::check_* (SURVEY.md section 8a):
 *
 *   VariableBase::msm over BLS12-377 G1        [UPSTREAM snarkvm-algorithms 1.0.0 @ dea322b,
 *                                               algorithms/src/msm/variable_base/{mod,batched,standard}.rs]
 *       oracle_msm_batched = batched::msm (what BLS12-377 G1 is dispatched to: affine pair additions with
 *       one inversion per batch), oracle_msm = standard::msm (Jacobian buckets; the independent cross-check)
 *   EvaluationDomain::{fft,ifft,coset_fft,coset_ifft}_in_place over Fr
 *                                              [UPSTREAM algorithms/src/fft/domain.rs]
 *   Fp256 / Fp384 Montgomery arithmetic on 64-bit limbs
 *                                              [UPSTREAM fields/src/{fp_256,fp_384}.rs]
 *   short-Weierstrass Jacobian add / mixed add / double (a = 0, b = 1)
 *                                              [UPSTREAM curves/src/templates/short_weierstrass_jacobian/]
 *
 * The snarkVM sources are an un-vendored git dependency of /root/reference (Cargo.toml:44-49,
 * Cargo.lock:3578-3606) and no Rust toolchain exists in the build image, so this file restates
 * the published algorithms with snarkVM's data conventions (64-bit little-endian limbs,
 * Montgomery bases / NTT data, canonical scalars, 104-byte affine stride).
 *
 * PARITY STATUS: parity unpinned at the snarkOS boundary (the reference carries no known-answer
 * vectors for this path); pinned instead against the first-principles vectors of SURVEY.md
 * appendix A (tests/golden/kat.json) and against the Python big-int oracle (oracle/bls12_377.py).
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
 * load this library.  The product (snarkos_b200/) never links or calls it.
 *
 * Build: see oracle/Makefile  (gcc -O3 -fopenmp -shared -fPIC).
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#include "bls12_377_constants.h"

typedef unsigned __int128 u128;
typedef uint64_t u64;

/* ------------------------------------------------------------------------------------------
 * Generic Montgomery arithmetic on N 64-bit limbs (CIOS; both moduli have spare top bits)
 * ---------------------------------------------------------------------------------------- */
#define DEFINE_FIELD(PFX, N, MOD, INV)                                                          \
    static inline int PFX##_geq(const u64 *a, const u64 *b) {                                   \
        for (int i = N - 1; i >= 0; i--) {                                                      \
            if (a[i] > b[i]) return 1;                                                          \
            if (a[i] < b[i]) return 0;                                                          \
        }                                                                                       \
        return 1;                                                                               \
    }                                                                                           \
    static inline int PFX##_is_zero(const u64 *a) {                                             \
        u64 t = 0;                                                                              \
        for (int i = 0; i < N; i++) t |= a[i];                                                  \
        return t == 0;                                                                          \
    }                                                                                           \
    static inline int PFX##_eq(const u64 *a, const u64 *b) {                                    \
        u64 t = 0;                                                                              \
        for (int i = 0; i < N; i++) t |= a[i] ^ b[i];                                           \
        return t == 0;                                                                          \
    }                                                                                           \
    static inline void PFX##_sub_nored(u64 *r, const u64 *a, const u64 *b) {                    \
        u64 borrow = 0;                                                                         \
        for (int i = 0; i < N; i++) {                                                           \
            u128 t = (u128)a[i] - b[i] - borrow;                                                \
            r[i] = (u64)t;                                                                      \
            borrow = (u64)(t >> 64) & 1;                                                        \
        }                                                                                       \
    }                                                                                           \
    static inline void PFX##_add(u64 *r, const u64 *a, const u64 *b) {                          \
        u64 carry = 0;                                                                          \
        for (int i = 0; i < N; i++) {                                                           \
            u128 t = (u128)a[i] + b[i] + carry;                                                 \
            r[i] = (u64)t;                                                                      \
            carry = (u64)(t >> 64);                                                             \
        }                                                                                       \
        if (PFX##_geq(r, MOD)) PFX##_sub_nored(r, r, MOD);                                      \
    }                                                                                           \
    static inline void PFX##_sub(u64 *r, const u64 *a, const u64 *b) {                          \
        u64 borrow = 0;                                                                         \
        u64 t2[N];                                                                              \
        for (int i = 0; i < N; i++) {                                                           \
            u128 t = (u128)a[i] - b[i] - borrow;                                                \
            t2[i] = (u64)t;                                                                     \
            borrow = (u64)(t >> 64) & 1;                                                        \
        }                                                                                       \
        if (borrow) {                                                                           \
            u64 carry = 0;                                                                      \
            for (int i = 0; i < N; i++) {                                                       \
                u128 t = (u128)t2[i] + MOD[i] + carry;                                          \
                t2[i] = (u64)t;                                                                 \
                carry = (u64)(t >> 64);                                                         \
            }                                                                                   \
        }                                                                                       \
        memcpy(r, t2, sizeof(t2));                                                              \
    }                                                                                           \
    static inline void PFX##_neg(u64 *r, const u64 *a) {                                        \
        if (PFX##_is_zero(a)) { memset(r, 0, N * 8); return; }                                  \
        PFX##_sub_nored(r, MOD, a);                                                             \
    }                                                                                           \
    static inline void PFX##_dbl(u64 *r, const u64 *a) { PFX##_add(r, a, a); }                  \
    /* "no-carry" CIOS [UPSTREAM fields/src/fp_256.rs / fp_384.rs mul_assign: the modulus leaves its top bit     \
     * spare, so product and reduction of one row run as two interleaved carry chains and t never needs an        \
     * extra limb]; fully unrolled (N is a constant) so that t lives in registers */                              \
    static inline void PFX##_mul(u64 *r, const u64 *a, const u64 *b) {                          \
        u64 t[N];                                                                               \
        _Pragma("GCC unroll 8") for (int j = 0; j < N; j++) t[j] = 0;                           \
        _Pragma("GCC unroll 8") for (int i = 0; i < N; i++) {                                   \
            u128 p = (u128)a[0] * b[i] + t[0];                                                  \
            u64 c1 = (u64)(p >> 64);                                                            \
            const u64 k = (u64)p * (u64)(INV);                                                  \
            u128 q = (u128)k * MOD[0] + (u64)p;                                                 \
            u64 c2 = (u64)(q >> 64);                                                            \
            _Pragma("GCC unroll 8") for (int j = 1; j < N; j++) {                               \
                p = (u128)a[j] * b[i] + t[j] + c1;                                              \
                c1 = (u64)(p >> 64);                                                            \
                q = (u128)k * MOD[j] + (u64)p + c2;                                             \
                c2 = (u64)(q >> 64);                                                            \
                t[j - 1] = (u64)q;                                                              \
            }                                                                                   \
            t[N - 1] = c1 + c2;                                                                 \
        }                                                                                       \
        if (PFX##_geq(t, MOD)) PFX##_sub_nored(t, t, MOD);                                      \
        memcpy(r, t, N * 8);                                                                    \
    }                                                                                           \
    static inline void PFX##_sqr(u64 *r, const u64 *a) { PFX##_mul(r, a, a); }                  \
    /* r = a^e, e given as `en` 64-bit limbs (plain binary exponent), a in Montgomery form */   \
    static void PFX##_pow(u64 *r, const u64 *a, const u64 *e, int en, const u64 *one) {         \
        u64 acc[N];                                                                             \
        memcpy(acc, one, N * 8);                                                                \
        int started = 0;                                                                        \
        for (int i = en * 64 - 1; i >= 0; i--) {                                                \
            if (started) PFX##_sqr(acc, acc);                                                   \
            if ((e[i / 64] >> (i % 64)) & 1) {                                                  \
                PFX##_mul(acc, acc, a);                                                         \
                started = 1;                                                                    \
            }                                                                                   \
        }                                                                                       \
        memcpy(r, acc, N * 8);                                                                  \
    }                                                                                           \
    /* Fermat inversion a^(m-2); a = 0 -> 0 */                                                  \
    static void PFX##_inv(u64 *r, const u64 *a, const u64 *one) {                               \
        u64 e[N];                                                                               \
        u64 two[N];                                                                             \
        memset(two, 0, sizeof(two));                                                            \
        two[0] = 2;                                                                             \
        PFX##_sub_nored(e, MOD, two);                                                           \
        PFX##_pow(r, a, e, N, one);                                                             \
    }

static const u64 FR_INV64 = FR_INV;
static const u64 FQ_INV64 = FQ_INV;
DEFINE_FIELD(fr, 4, FR_MODULUS, FR_INV64)
DEFINE_FIELD(fq, 6, FQ_MODULUS, FQ_INV64)

/* ------------------------------------------------------------------------------------------
 * G1 Jacobian arithmetic (a = 0):  dbl-2009-l, madd-2007-bl, add-2007-bl   
 * ---------------------------------------------------------------------------------------- */
typedef struct { u64 x[6], y[6], z[6]; } jac_t;     /* z == 0  <=> infinity */
typedef struct { u64 x[6], y[6]; int inf; } aff_t;

static void jac_set_inf(jac_t *p) {
    memcpy(p->x, FQ_ONE, 48);
    memcpy(p->y, FQ_ONE, 48);
    memset(p->z, 0, 48);
}

static void jac_double(jac_t *r, const jac_t *p) {
    if (fq_is_zero(p->z)) { *r = *p; return; }
    u64 A[6], B[6], C[6], D[6], E[6], F[6], t[6], X3[6], Y3[6], Z3[6];
    fq_sqr(A, p->x);
    fq_sqr(B, p->y);
    fq_sqr(C, B);
    fq_add(t, p->x, B);
    fq_sqr(t, t);
    fq_sub(t, t, A);
    fq_sub(t, t, C);
    fq_dbl(D, t);
    fq_dbl(E, A);
    fq_add(E, E, A);
    fq_sqr(F, E);
    fq_dbl(t, D);
    fq_sub(X3, F, t);
    fq_sub(t, D, X3);
    fq_mul(Y3, E, t);
    fq_dbl(t, C); fq_dbl(t, t); fq_dbl(t, t);
    fq_sub(Y3, Y3, t);
    fq_mul(Z3, p->y, p->z);
    fq_dbl(Z3, Z3);
    memcpy(r->x, X3, 48); memcpy(r->y, Y3, 48); memcpy(r->z, Z3, 48);
}

static void jac_add_affine(jac_t *r, const jac_t *p, const aff_t *q) {
    if (q->inf) { *r = *p; return; }
    if (fq_is_zero(p->z)) {
        memcpy(r->x, q->x, 48); memcpy(r->y, q->y, 48); memcpy(r->z, FQ_ONE, 48);
        return;
    }
    u64 Z1Z1[6], U2[6], S2[6], H[6], HH[6], I[6], J[6], rr[6], V[6], t[6], X3[6], Y3[6], Z3[6];
    fq_sqr(Z1Z1, p->z);
    fq_mul(U2, q->x, Z1Z1);
    fq_mul(S2, q->y, p->z);
    fq_mul(S2, S2, Z1Z1);
    fq_sub(H, U2, p->x);
    fq_sub(rr, S2, p->y);
    if (fq_is_zero(H)) {
        if (fq_is_zero(rr)) { jac_double(r, p); return; }
        jac_set_inf(r);
        return;
    }
    fq_sqr(HH, H);
    fq_dbl(I, HH); fq_dbl(I, I);
    fq_mul(J, H, I);
    fq_dbl(rr, rr);
    fq_mul(V, p->x, I);
    fq_sqr(X3, rr);
    fq_sub(X3, X3, J);
    fq_dbl(t, V);
    fq_sub(X3, X3, t);
    fq_sub(t, V, X3);
    fq_mul(Y3, rr, t);
    fq_mul(t, p->y, J);
    fq_dbl(t, t);
    fq_sub(Y3, Y3, t);
    fq_add(Z3, p->z, H);
    fq_sqr(Z3, Z3);
    fq_sub(Z3, Z3, Z1Z1);
    fq_sub(Z3, Z3, HH);
    memcpy(r->x, X3, 48); memcpy(r->y, Y3, 48); memcpy(r->z, Z3, 48);
}

static void jac_add(jac_t *r, const jac_t *p, const jac_t *q) {
    if (fq_is_zero(q->z)) { *r = *p; return; }
    if (fq_is_zero(p->z)) { *r = *q; return; }
    u64 Z1Z1[6], Z2Z2[6], U1[6], U2[6], S1[6], S2[6], H[6], I[6], J[6], rr[6], V[6], t[6];
    u64 X3[6], Y3[6], Z3[6];
    fq_sqr(Z1Z1, p->z);
    fq_sqr(Z2Z2, q->z);
    fq_mul(U1, p->x, Z2Z2);
    fq_mul(U2, q->x, Z1Z1);
    fq_mul(S1, p->y, q->z); fq_mul(S1, S1, Z2Z2);
    fq_mul(S2, q->y, p->z); fq_mul(S2, S2, Z1Z1);
    fq_sub(H, U2, U1);
    fq_sub(rr, S2, S1);
    if (fq_is_zero(H)) {
        if (fq_is_zero(rr)) { jac_double(r, p); return; }
        jac_set_inf(r);
        return;
    }
    fq_dbl(I, H); fq_sqr(I, I);
    fq_mul(J, H, I);
    fq_dbl(rr, rr);
    fq_mul(V, U1, I);
    fq_sqr(X3, rr);
    fq_sub(X3, X3, J);
    fq_dbl(t, V);
    fq_sub(X3, X3, t);
    fq_sub(t, V, X3);
    fq_mul(Y3, rr, t);
    fq_mul(t, S1, J);
    fq_dbl(t, t);
    fq_sub(Y3, Y3, t);
    fq_add(Z3, p->z, q->z);
    fq_sqr(Z3, Z3);
    fq_sub(Z3, Z3, Z1Z1);
    fq_sub(Z3, Z3, Z2Z2);
    fq_mul(Z3, Z3, H);
    memcpy(r->x, X3, 48); memcpy(r->y, Y3, 48); memcpy(r->z, Z3, 48);
}

static void jac_to_affine(aff_t *r, const jac_t *p) {
    if (fq_is_zero(p->z)) { memset(r, 0, sizeof(*r)); r->inf = 1; return; }
    u64 zi[6], zi2[6];
    fq_inv(zi, p->z, FQ_ONE);
    fq_sqr(zi2, zi);
    fq_mul(r->x, p->x, zi2);
    fq_mul(zi2, zi2, zi);
    fq_mul(r->y, p->y, zi2);
    r->inf = 0;
}

static void load_affine(aff_t *a, const uint8_t *src) {
    memcpy(a->x, src, 48);
    memcpy(a->y, src + 48, 48);
    a->inf = src[96] != 0;
}

static void store_affine(uint8_t *dst, const aff_t *a, size_t stride) {
    memset(dst, 0, stride);
    if (a->inf) { dst[96] = 1; return; }
    memcpy(dst, a->x, 48);
    memcpy(dst + 48, a->y, 48);
}

/* ------------------------------------------------------------------------------------------
 * Exported helpers (ctypes): element-wise field ops for pinning against the Python oracle
 * ---------------------------------------------------------------------------------------- */
void oracle_fr_mul(u64 *out, const u64 *a, const u64 *b, size_t n) {
    for (size_t i = 0; i < n; i++) fr_mul(out + 4 * i, a + 4 * i, b + 4 * i);
}
void oracle_fq_mul(u64 *out, const u64 *a, const u64 *b, size_t n) {
    for (size_t i = 0; i < n; i++) fq_mul(out + 6 * i, a + 6 * i, b + 6 * i);
}
void oracle_fr_add(u64 *out, const u64 *a, const u64 *b, size_t n) {
    for (size_t i = 0; i < n; i++) fr_add(out + 4 * i, a + 4 * i, b + 4 * i);
}
void oracle_fr_sub(u64 *out, const u64 *a, const u64 *b, size_t n) {
    for (size_t i = 0; i < n; i++) fr_sub(out + 4 * i, a + 4 * i, b + 4 * i);
}
void oracle_fq_add(u64 *out, const u64 *a, const u64 *b, size_t n) {
    for (size_t i = 0; i < n; i++) fq_add(out + 6 * i, a + 6 * i, b + 6 * i);
}
void oracle_fq_sub(u64 *out, const u64 *a, const u64 *b, size_t n) {
    for (size_t i = 0; i < n; i++) fq_sub(out + 6 * i, a + 6 * i, b + 6 * i);
}
void oracle_fq_inv(u64 *out, const u64 *a, size_t n) {
    for (size_t i = 0; i < n; i++) fq_inv(out + 6 * i, a + 6 * i, FQ_ONE);
}
void oracle_fr_inv(u64 *out, const u64 *a, size_t n) {
    for (size_t i = 0; i < n; i++) fr_inv(out + 4 * i, a + 4 * i, FR_ONE);
}
/* canonical <-> Montgomery */
void oracle_fr_to_mont(u64 *out, const u64 *a, size_t n) {
    for (size_t i = 0; i < n; i++) fr_mul(out + 4 * i, a + 4 * i, FR_R2);
}
void oracle_fr_from_mont(u64 *out, const u64 *a, size_t n) {
    u64 one[4] = {1, 0, 0, 0};
    for (size_t i = 0; i < n; i++) fr_mul(out + 4 * i, a + 4 * i, one);
}

/* Jacobian (144 B) -> affine (stride bytes) */
void oracle_g1_to_affine(uint8_t *out_affine, const uint8_t *in_jac, size_t n, size_t stride) {
    for (size_t i = 0; i < n; i++) {
        jac_t p;
        memcpy(&p, in_jac + 144 * i, 144);
        aff_t a;
        jac_to_affine(&a, &p);
        store_affine(out_affine + stride * i, &a, stride);
    }
}

/* Jacobian (144 B) -> compressed G1Affine (48 B): canonical x little-endian, flags in the last byte -- bit 7 set iff
 * y > -y (as canonical integers), bit 6 = infinity (x serialised as 0)
 * [UPSTREAM curves/src/templates/short_weierstrass_jacobian/affine.rs serialize_with_mode(Compress::Yes);
 *  utilities/src/serialize/flags.rs SWFlags::u8_bitmask] -- recalled, see oracle/bls12_377.py g1_compress */
void oracle_g1_compress(uint8_t *out48, const uint8_t *in_jac, size_t n) {
    static const u64 one_plain[6] = {1, 0, 0, 0, 0, 0};
    for (size_t i = 0; i < n; i++) {
        jac_t p;
        memcpy(&p, in_jac + 144 * i, 144);
        aff_t a;
        jac_to_affine(&a, &p);
        uint8_t *o = out48 + 48 * i;
        memset(o, 0, 48);
        if (a.inf) { o[47] |= 1u << 6; continue; }
        u64 x[6], y[6], ny[6];
        fq_mul(x, a.x, one_plain);                     /* out of Montgomery form */
        fq_mul(y, a.y, one_plain);
        fq_sub_nored(ny, FQ_MODULUS, y);                  /* p - y; y != 0 on this curve's prime-order subgroup */
        memcpy(o, x, 48);
        if (!fq_is_zero(y) && fq_geq(y, ny) && !fq_eq(y, ny)) o[47] |= 1u << 7;
    }
}

/* out[i] = k[i] * base  (64-bit scalars, MSB-first double-and-add), affine out */
void oracle_g1_mul_u64(uint8_t *out_affine, const uint8_t *base_affine, const u64 *k, size_t n,
                       size_t stride) {
    aff_t b;
    load_affine(&b, base_affine);
#pragma omp parallel for schedule(static)
    for (long i = 0; i < (long)n; i++) {
        jac_t acc;
        jac_set_inf(&acc);
        for (int bit = 63; bit >= 0; bit--) {
            jac_double(&acc, &acc);
            if ((k[i] >> bit) & 1) jac_add_affine(&acc, &acc, &b);
        }
        aff_t a;
        jac_to_affine(&a, &acc);
        store_affine(out_affine + stride * (size_t)i, &a, stride);
    }
}

int oracle_g1_is_on_curve(const uint8_t *affine) {
    aff_t a;
    load_affine(&a, affine);
    if (a.inf) return 1;
    u64 l[6], r[6];
    fq_sqr(l, a.y);
    fq_sqr(r, a.x);
    fq_mul(r, r, a.x);
    fq_add(r, r, FQ_ONE);
    return fq_eq(l, r);
}

/* ------------------------------------------------------------------------------------------
 * VariableBase::msm restated (standard::msm shape: window c = ln(n) + 2, one task per window,
 * buckets in Jacobian with mixed adds, running-sum reduction, fold high -> low).
 * n < 15 takes the bit-serial double-and-add path like batched::msm's small-size guard.
 * Output: Jacobian 144 B (X, Y, Z Montgomery).   nthreads <= 0 -> all cores.
 * ---------------------------------------------------------------------------------------- */
static inline unsigned get_window(const u64 *s, int lo, int c) {
    /* bits [lo, lo + c) of a 256-bit little-endian scalar */
    int limb = lo >> 6, off = lo & 63;
    u64 v = s[limb] >> off;
    if (off + c > 64 && limb + 1 < 4) v |= s[limb + 1] << (64 - off);
    return (unsigned)(v & ((1ull << c) - 1));
}

int oracle_msm_window_bits(size_t n) {
    if (n < 32) return 1;
    return (int)log((double)n) + 2;            /* snarkVM: ln_without_floats(n) + 2 */
}

void oracle_msm(uint8_t *out_jac, const uint8_t *bases, size_t n, size_t stride,
                const u64 *scalars, int nthreads) {
    jac_t total;
    jac_set_inf(&total);
    if (n == 0) { memcpy(out_jac, &total, 144); return; }
#ifdef _OPENMP
    if (nthreads <= 0) nthreads = omp_get_max_threads();
#else
    nthreads = 1;
#endif
    if (n < 15) {
        for (size_t i = 0; i < n; i++) {
            aff_t b;
            load_affine(&b, bases + i * stride);
            jac_t acc;
            jac_set_inf(&acc);
            const u64 *s = scalars + 4 * i;
            for (int bit = 255; bit >= 0; bit--) {
                jac_double(&acc, &acc);
                if ((s[bit >> 6] >> (bit & 63)) & 1) jac_add_affine(&acc, &acc, &b);
            }
            jac_add(&total, &total, &acc);
        }
        memcpy(out_jac, &total, 144);
        return;
    }
    int c = oracle_msm_window_bits(n);
    int nwin = (253 + c - 1) / c;
    jac_t *wsum = (jac_t *)malloc(sizeof(jac_t) * nwin);
    size_t nb = ((size_t)1 << c) - 1;
#pragma omp parallel for schedule(dynamic, 1) num_threads(nthreads)
    for (int w = 0; w < nwin; w++) {
        jac_t *buckets = (jac_t *)malloc(sizeof(jac_t) * nb);
        for (size_t b = 0; b < nb; b++) jac_set_inf(&buckets[b]);
        for (size_t i = 0; i < n; i++) {
            unsigned d = get_window(scalars + 4 * i, w * c, c);
            if (!d) continue;
            aff_t p;
            load_affine(&p, bases + i * stride);
            jac_add_affine(&buckets[d - 1], &buckets[d - 1], &p);
        }
        jac_t running, acc;
        jac_set_inf(&running);
        jac_set_inf(&acc);
        for (size_t b = nb; b-- > 0;) {
            jac_add(&running, &running, &buckets[b]);
            jac_add(&acc, &acc, &running);
        }
        wsum[w] = acc;
        free(buckets);
    }
    for (int w = nwin - 1; w >= 0; w--) {
        for (int k = 0; k < c; k++) jac_double(&total, &total);
        jac_add(&total, &total, &wsum[w]);
    }
    free(wsum);
    memcpy(out_jac, &total, 144);
}


/* ------------------------------------------------------------------------------------------
 * batched::msm restated -- the variant VariableBase::msm dispatches BLS12-377 G1 to
 *   [UPSTREAM algorithms/src/msm/variable_base/batched.rs: msm / batched_window / batch_add /
 *    batch_add_in_place, curves/.../short_weierstrass_jacobian/affine.rs: batch_add_loop_1 / batch_add_loop_2]
 *
 *   n < 15              bit-serial double-and-add over all bases
 *   c = 1 (n < 32) else ln(n) + 2, unsigned c-bit windows, 2^c - 1 buckets, one task per window (rayon there,
 *   OpenMP here -- so at most ceil(253 / c) threads are ever busy, exactly like upstream)
 *   per window: (bucket, index) positions sorted by bucket; rounds of PAIRWISE AFFINE additions, flushed in batches of
 *   batch_size / 2 pairs that share ONE field inversion (Montgomery's trick: ~6 Fq products per addition instead of
 *   the ~11 of a Jacobian mixed addition); the first round reads the bases, later rounds work in place; then the
 *   running-sum reduction in Jacobian coordinates and the fold over windows, high to low.
 * Only the order inside a bucket may differ from upstream (stable counting sort instead of sort_unstable); the group
 * element is the same.
 * ---------------------------------------------------------------------------------------- */
static size_t batched_batch_size(size_t n) { return n < 500000 ? 300 : 3000; }    /* upstream get_batch_size (x86_64) */

typedef struct { u64 x[6], y[6]; } affp_t;            /* affine, (0, 0) = infinity */
static inline int affp_is_inf(const affp_t *p) { return fq_is_zero(p->x) && fq_is_zero(p->y); }

/* out[i] = a[i] + b[i] for cnt pairs with one shared inversion.  Exceptional pairs are resolved on the spot. */
static void batched_pairs_add(affp_t *out, const affp_t *a, const affp_t *b, size_t cnt, u64 (*den)[6], u64 (*pre)[6]) {
    u64 acc[6];
    memcpy(acc, FQ_ONE, 48);
    /* loop 1: denominators and running products */
    for (size_t i = 0; i < cnt; i++) {
        memcpy(pre[i], acc, 48);
        if (affp_is_inf(&a[i]) || affp_is_inf(&b[i])) { memset(den[i], 0, 48); continue; }
        if (fq_eq(a[i].x, b[i].x)) {
            if (fq_eq(a[i].y, b[i].y) && !fq_is_zero(a[i].y)) fq_dbl(den[i], a[i].y);   /* doubling: 2 y */
            else { memset(den[i], 0, 48); continue; }                                   /* P + (-P) */
        } else {
            fq_sub(den[i], b[i].x, a[i].x);
        }
        fq_mul(acc, acc, den[i]);
    }
    u64 inv[6];
    fq_inv(inv, acc, FQ_ONE);
    /* loop 2: walk back, peel the inverses off, finish the additions */
    for (size_t i = cnt; i-- > 0;) {
        if (fq_is_zero(den[i])) {
            if (affp_is_inf(&a[i])) out[i] = b[i];
            else if (affp_is_inf(&b[i])) out[i] = a[i];
            else memset(&out[i], 0, sizeof(affp_t));
            continue;
        }
        u64 dinv[6], lam[6], num[6], x3[6], y3[6], t[6];
        fq_mul(dinv, inv, pre[i]);
        fq_mul(inv, inv, den[i]);
        if (fq_eq(a[i].x, b[i].x)) {                   /* doubling: lambda = 3 x^2 / 2 y */
            fq_sqr(t, a[i].x);
            fq_dbl(num, t);
            fq_add(num, num, t);
        } else {
            fq_sub(num, b[i].y, a[i].y);
        }
        fq_mul(lam, num, dinv);
        fq_sqr(x3, lam);
        fq_sub(x3, x3, a[i].x);
        fq_sub(x3, x3, b[i].x);
        fq_sub(t, a[i].x, x3);
        fq_mul(y3, lam, t);
        fq_sub(y3, y3, a[i].y);
        memcpy(out[i].x, x3, 48);
        memcpy(out[i].y, y3, 48);
    }
}

static void batched_window(jac_t *res, const uint8_t *bases, size_t n, size_t stride, const u64 *scalars, int w_start, int c) {
    const size_t nb = ((size_t)1 << c) - 1;
    const size_t batch = batched_batch_size(n) / 2;
    /* positions sorted by bucket (counting sort): idx[off[b] .. off[b + 1]) = scalar indices of bucket b */
    uint32_t *off = (uint32_t *)calloc(nb + 2, sizeof(uint32_t));
    uint32_t *idx = (uint32_t *)malloc(sizeof(uint32_t) * (n ? n : 1));
    for (size_t i = 0; i < n; i++) {
        unsigned d = get_window(scalars + 4 * i, w_start, c);
        if (d) off[d + 1]++;
    }
    for (size_t b = 1; b <= nb + 1; b++) off[b] += off[b - 1];
    uint32_t *cur = (uint32_t *)malloc(sizeof(uint32_t) * (nb + 2));
    memcpy(cur, off, sizeof(uint32_t) * (nb + 2));
    for (size_t i = 0; i < n; i++) {
        unsigned d = get_window(scalars + 4 * i, w_start, c);
        if (d) idx[cur[d]++] = (uint32_t)i;
    }
    const size_t total = off[nb + 1];
    /* lists: cnt[b] points of bucket b at pts[start[b]..]; the first round builds them from the bases */
    affp_t *pts = (affp_t *)malloc(sizeof(affp_t) * (total / 2 + nb + 1));
    uint32_t *start = (uint32_t *)malloc(sizeof(uint32_t) * (nb + 1));
    uint32_t *cnt = (uint32_t *)malloc(sizeof(uint32_t) * (nb + 1));
    affp_t *pa = (affp_t *)malloc(sizeof(affp_t) * batch), *pb = (affp_t *)malloc(sizeof(affp_t) * batch);
    affp_t *po = (affp_t *)malloc(sizeof(affp_t) * batch);
    u64(*den)[6] = (u64(*)[6])malloc(48 * batch);
    u64(*pre)[6] = (u64(*)[6])malloc(48 * batch);
    size_t *dst = (size_t *)malloc(sizeof(size_t) * batch);
    size_t pend = 0, wr = 0;
    int all_ones = 1;
#define FLUSH()                                                          \
    do {                                                                 \
        if (pend) {                                                      \
            batched_pairs_add(po, pa, pb, pend, den, pre);               \
            for (size_t q_ = 0; q_ < pend; q_++) pts[dst[q_]] = po[q_];  \
            pend = 0;                                                    \
        }                                                                \
    } while (0)
    for (size_t b = 1; b <= nb; b++) {                 /* round 1: bases -> lists (batch_add_write) */
        const uint32_t lo = off[b], m = off[b + 1] - lo;
        start[b] = (uint32_t)wr;
        cnt[b] = (m + 1) / 2;
        if (cnt[b] > 1) all_ones = 0;
        for (uint32_t j = 0; j + 1 < m; j += 2) {
            aff_t A, B;
            load_affine(&A, bases + (size_t)idx[lo + j] * stride);
            load_affine(&B, bases + (size_t)idx[lo + j + 1] * stride);
            memset(&pa[pend], 0, sizeof(affp_t));
            memset(&pb[pend], 0, sizeof(affp_t));
            if (!A.inf) { memcpy(pa[pend].x, A.x, 48); memcpy(pa[pend].y, A.y, 48); }
            if (!B.inf) { memcpy(pb[pend].x, B.x, 48); memcpy(pb[pend].y, B.y, 48); }
            dst[pend++] = wr++;
            if (pend == batch) FLUSH();
        }
        if (m & 1) {                                   /* the odd one out is copied through (upstream: !0u32 instruction) */
            aff_t A;
            load_affine(&A, bases + (size_t)idx[lo + m - 1] * stride);
            memset(&pts[wr], 0, sizeof(affp_t));
            if (!A.inf) { memcpy(pts[wr].x, A.x, 48); memcpy(pts[wr].y, A.y, 48); }
            wr++;
        }
    }
    FLUSH();
    while (!all_ones) {                                /* later rounds in place (batch_add_in_place) */
        all_ones = 1;
        for (size_t b = 1; b <= nb; b++) {
            const uint32_t m = cnt[b], s0 = start[b];
            if (m < 2) continue;
            /* results go to the front of the bucket's own range: slot j/2 <- slots j, j + 1.  The operands are copied
             * out when the pair is queued and a queued result lands at or below the slots already consumed, so
             * nothing still to be read in this round is overwritten; every queue is flushed before the next round */
            for (uint32_t j = 0; j + 1 < m; j += 2) {
                pa[pend] = pts[s0 + j];
                pb[pend] = pts[s0 + j + 1];
                dst[pend++] = s0 + j / 2;
                if (pend == batch) FLUSH();
            }
            if (m & 1) pts[s0 + m / 2] = pts[s0 + m - 1];
            cnt[b] = (m + 1) / 2;
            if (cnt[b] > 1) all_ones = 0;
        }
        FLUSH();
    }
#undef FLUSH
    /* running-sum reduction: res = sum_b b * bucket_b */
    jac_t running, acc;
    jac_set_inf(&running);
    jac_set_inf(&acc);
    for (size_t b = nb; b >= 1; b--) {
        if (cnt[b]) {
            aff_t q;
            q.inf = affp_is_inf(&pts[start[b]]);
            memcpy(q.x, pts[start[b]].x, 48);
            memcpy(q.y, pts[start[b]].y, 48);
            jac_add_affine(&running, &running, &q);
        }
        jac_add(&acc, &acc, &running);
    }
    *res = acc;
    free(off); free(idx); free(cur); free(pts); free(start); free(cnt); free(pa); free(pb); free(po); free(den); free(pre); free(dst);
}

void oracle_msm_batched(uint8_t *out_jac, const uint8_t *bases, size_t n, size_t stride, const u64 *scalars, int nthreads) {
    jac_t total;
    jac_set_inf(&total);
    if (n < 15) { oracle_msm(out_jac, bases, n, stride, scalars, nthreads); return; }   /* same bit-serial path */
#ifdef _OPENMP
    if (nthreads <= 0) nthreads = omp_get_max_threads();
#else
    nthreads = 1;
#endif
    const int c = oracle_msm_window_bits(n);
    const int nwin = (253 + c - 1) / c;
    jac_t *wsum = (jac_t *)malloc(sizeof(jac_t) * nwin);
#pragma omp parallel for schedule(dynamic, 1) num_threads(nthreads)
    for (int w = 0; w < nwin; w++) batched_window(&wsum[w], bases, n, stride, scalars, w * c, c);
    for (int w = nwin - 1; w >= 0; w--) {
        for (int k = 0; k < c; k++) jac_double(&total, &total);
        jac_add(&total, &total, &wsum[w]);
    }
    free(wsum);
    memcpy(out_jac, &total, 144);
}

/* out[i] = (k0 + i) * G as affine images: n DISTINCT points in ~1 us each (one mixed addition per point, one shared
 * inversion per chunk) -- inputs for the CPU arm of bench.py at full size without any GPU code. */
void oracle_g1_sequence(uint8_t *out_affine, u64 k0, size_t n, size_t stride, int nthreads) {
#ifdef _OPENMP
    if (nthreads <= 0) nthreads = omp_get_max_threads();
#else
    nthreads = 1;
#endif
    aff_t g;
    memcpy(g.x, G1_GEN_X, 48);
    memcpy(g.y, G1_GEN_Y, 48);
    g.inf = 0;
    const size_t chunk = 4096;
    const long nchunks = (long)((n + chunk - 1) / chunk);
#pragma omp parallel for schedule(dynamic, 4) num_threads(nthreads)
    for (long ci = 0; ci < nchunks; ci++) {
        const size_t lo = (size_t)ci * chunk, hi = lo + chunk < n ? lo + chunk : n, m = hi - lo;
        jac_t *J = (jac_t *)malloc(sizeof(jac_t) * m);
        u64(*pre)[6] = (u64(*)[6])malloc(48 * m);
        jac_t acc;
        jac_set_inf(&acc);
        const u64 k = k0 + lo;
        for (int bit = 63; bit >= 0; bit--) {
            jac_double(&acc, &acc);
            if ((k >> bit) & 1) jac_add_affine(&acc, &acc, &g);
        }
        u64 prod[6];
        memcpy(prod, FQ_ONE, 48);
        for (size_t i = 0; i < m; i++) {
            J[i] = acc;
            memcpy(pre[i], prod, 48);
            if (!fq_is_zero(acc.z)) fq_mul(prod, prod, acc.z);
            jac_add_affine(&acc, &acc, &g);
        }
        u64 inv[6];
        fq_inv(inv, prod, FQ_ONE);
        for (size_t i = m; i-- > 0;) {
            aff_t a;
            if (fq_is_zero(J[i].z)) { memset(&a, 0, sizeof(a)); a.inf = 1; }
            else {
                u64 zi[6], zi2[6];
                fq_mul(zi, inv, pre[i]);
                fq_mul(inv, inv, J[i].z);
                fq_sqr(zi2, zi);
                fq_mul(a.x, J[i].x, zi2);
                fq_mul(zi2, zi2, zi);
                fq_mul(a.y, J[i].y, zi2);
                a.inf = 0;
            }
            store_affine(out_affine + stride * (lo + i), &a, stride);
        }
        free(J);
        free(pre);
    }
}

/* ------------------------------------------------------------------------------------------
 * EvaluationDomain over Fr: in-order radix-2 (i)FFT with a precomputed root table
 * (io_helper / oi_helper + derange restated as bit-reverse + DIT), Montgomery data in place.
 *   direction: 0 forward, 1 inverse;  coset: 0 / 1 (generator 22)     SURVEY appendix A.1
 * ---------------------------------------------------------------------------------------- */
static void fr_pow_u64(u64 *r, const u64 *a, u64 e) {
    u64 ee[1] = {e};
    fr_pow(r, a, ee, 1, FR_ONE);
}

static void domain_group_gen(u64 *gen, int log_n, int inverse) {
    /* group_gen = TWO_ADIC_ROOT ^ (2^(47 - log_n)) */
    memcpy(gen, inverse ? FR_TWO_ADIC_ROOT_INV : FR_TWO_ADIC_ROOT, 32);
    for (int i = 0; i < FR_TWO_ADICITY - log_n; i++) fr_sqr(gen, gen);
}

static inline size_t bitrev(size_t x, int bits) {
    size_t r = 0;
    for (int i = 0; i < bits; i++) { r = (r << 1) | (x & 1); x >>= 1; }
    return r;
}

/* FFTPrecomputation restated [UPSTREAM algorithms/src/fft/domain.rs: precompute_fft / precompute_ifft]: snarkVM builds
 * the n/2 powers of the root ONCE per domain and callers keep it (Varuna holds it in the proving key's domains), so the
 * table -- and size_inv, and the coset powers -- are cached here per (log_n, direction) and built in parallel; a timed
 * oracle_ntt call after the first one does transforms only. */
typedef struct { u64 *roots; u64 *cpow; u64 size_inv[4]; int ready; } ntt_pc_t;
static ntt_pc_t g_ntt_pc[64][2];

static void powers_parallel(u64 *out, const u64 *base, size_t count, int nthreads) {
    /* out[i] = base^i: chunk heads by square-and-multiply, then one multiplication per element */
    const size_t chunk = 1 << 12;
    const long nchunks = (long)((count + chunk - 1) / chunk);
    (void)nthreads;
#pragma omp parallel for schedule(static) num_threads(nthreads)
    for (long ci = 0; ci < nchunks; ci++) {
        const size_t lo = (size_t)ci * chunk, hi = lo + chunk < count ? lo + chunk : count;
        u64 e[1] = {lo};
        fr_pow(out + 4 * lo, base, e, 1, FR_ONE);
        for (size_t i = lo + 1; i < hi; i++) fr_mul(out + 4 * i, out + 4 * (i - 1), base);
    }
}

static ntt_pc_t *ntt_precomputation(int log_n, int direction, int need_coset, int nthreads) {
    ntt_pc_t *pc = &g_ntt_pc[log_n][direction];
#pragma omp critical(oracle_ntt_pc)
    {
        const size_t n = (size_t)1 << log_n;
        if (!pc->ready) {
            u64 gen[4];
            domain_group_gen(gen, log_n, direction);
            const size_t half = n > 1 ? n / 2 : 1;
            pc->roots = (u64 *)malloc(32 * half);
            powers_parallel(pc->roots, gen, half, nthreads);
            u64 nn[4] = {n, 0, 0, 0};
            fr_mul(nn, nn, FR_R2);
            fr_inv(pc->size_inv, nn, FR_ONE);
            pc->ready = 1;
        }
        if (need_coset && !pc->cpow) {
            pc->cpow = (u64 *)malloc(32 * n);
            powers_parallel(pc->cpow, direction ? FR_GENERATOR_INV : FR_GENERATOR, n, nthreads);
        }
    }
    return pc;
}

static void ntt_one(u64 *a, int log_n, const u64 *roots /* n/2 powers of omega */, int par,
                    int nthreads) {
    size_t n = (size_t)1 << log_n;
    (void)par; (void)nthreads;
#pragma omp parallel for schedule(static) if (par) num_threads(nthreads)
    for (long ii = 0; ii < (long)n; ii++) {
        size_t i = (size_t)ii, j = bitrev(i, log_n);
        if (i < j) {
            u64 t[4];
            memcpy(t, a + 4 * i, 32); memcpy(a + 4 * i, a + 4 * j, 32); memcpy(a + 4 * j, t, 32);
        }
    }
    /* upstream compacts the roots a layer needs into a contiguous run when the stride through the table gets large
     * (cache behaviour of the short-distance layers); same here */
    u64 *compact = (u64 *)malloc(32 * (n > 1 ? n / 2 : 1));
    for (int s = 0; s < log_n; s++) {
        size_t m = (size_t)1 << s;
        size_t tw_stride = n >> (s + 1);
        const u64 *tw = roots;
        size_t tws = tw_stride;
        if (tw_stride > 1 && m >= 2) {
#pragma omp parallel for schedule(static) if (par && m >= 4096) num_threads(nthreads)
            for (long j = 0; j < (long)m; j++) memcpy(compact + 4 * j, roots + 4 * ((size_t)j * tw_stride), 32);
            tw = compact;
            tws = 1;
        }
        /* butterfly q of this stage: block k = (q / m) * 2m, offset j = q % m */
#pragma omp parallel for schedule(static) if (par) num_threads(nthreads)
        for (long q = 0; q < (long)(n / 2); q++) {
            size_t j = (size_t)q & (m - 1);
            size_t k = (((size_t)q) >> s) << (s + 1);
            u64 t[4], u[4];
            if (j) fr_mul(t, a + 4 * (k + j + m), tw + 4 * (j * tws));
            else memcpy(t, a + 4 * (k + j + m), 32);
            memcpy(u, a + 4 * (k + j), 32);
            fr_add(a + 4 * (k + j), u, t);
            fr_sub(a + 4 * (k + j + m), u, t);
        }
    }
    free(compact);
}

void oracle_ntt(u64 *data, int log_n, size_t batch, size_t batch_stride_elems, int direction,
                int coset, int nthreads) {
    size_t n = (size_t)1 << log_n;
#ifdef _OPENMP
    if (nthreads <= 0) nthreads = omp_get_max_threads();
#else
    nthreads = 1;
#endif
    const ntt_pc_t *pc = ntt_precomputation(log_n, direction, coset, nthreads);
    const u64 *roots = pc->roots, *cpow = pc->cpow, *size_inv = pc->size_inv;
    int outer = (batch >= (size_t)nthreads) || nthreads == 1;   /* parallel over polys, else inside one */
#pragma omp parallel for schedule(dynamic, 1) if (outer) num_threads(nthreads)
    for (long b = 0; b < (long)batch; b++) {
        u64 *a = data + 4 * (size_t)b * batch_stride_elems;
        if (coset && !direction) {
#pragma omp parallel for schedule(static) if (!outer) num_threads(nthreads)
            for (long i = 0; i < (long)n; i++) fr_mul(a + 4 * i, a + 4 * i, cpow + 4 * i);
        }
        ntt_one(a, log_n, roots, !outer, nthreads);
        if (direction) {
#pragma omp parallel for schedule(static) if (!outer) num_threads(nthreads)
            for (long i = 0; i < (long)n; i++) fr_mul(a + 4 * i, a + 4 * i, size_inv);
        }
        if (coset && direction) {
#pragma omp parallel for schedule(static) if (!outer) num_threads(nthreads)
            for (long i = 0; i < (long)n; i++) fr_mul(a + 4 * i, a + 4 * i, cpow + 4 * i);
        }
    }
    (void)fr_pow_u64;
}

/* many independent MSMs, one task per MSM (rayon over transactions in the reference's block verification) */
void oracle_msm_many(uint8_t *out_jac, const uint8_t *bases, const u64 *scalars, const u64 *offsets, size_t nmsm,
                     size_t stride, int nthreads) {
#ifdef _OPENMP
    if (nthreads <= 0) nthreads = omp_get_max_threads();
#else
    nthreads = 1;
#endif
#pragma omp parallel for schedule(dynamic, 1) num_threads(nthreads)
    for (long m = 0; m < (long)nmsm; m++) {
        const size_t lo = offsets[m], hi = offsets[m + 1];
        oracle_msm_batched(out_jac + 144 * (size_t)m, bases + lo * stride, hi - lo, stride, scalars + 4 * lo, 1);
    }
}

int oracle_num_threads(void) {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}
