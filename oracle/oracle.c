/*
 * oracle.c -- CPU restatement (TEST INFRASTRUCTURE ONLY) of the snarkVM hot path that snarkOS
 * reaches through VM::execute / Ledger::check_* (SURVEY.md section 8a):
 *
 *   VariableBase::msm over BLS12-377 G1        [UPSTREAM snarkvm-algorithms 1.0.0 @ dea322b,
 *                                               algorithms/src/msm/variable_base/{mod,standard}.rs]
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
    static inline void PFX##_mul(u64 *r, const u64 *a, const u64 *b) {                          \
        u64 t[N + 2];                                                                           \
        memset(t, 0, sizeof(t));                                                                \
        for (int i = 0; i < N; i++) {                                                           \
            u64 c = 0;                                                                          \
            for (int j = 0; j < N; j++) {                                                       \
                u128 p = (u128)a[j] * b[i] + t[j] + c;                                          \
                t[j] = (u64)p;                                                                  \
                c = (u64)(p >> 64);                                                             \
            }                                                                                   \
            u128 s = (u128)t[N] + c;                                                            \
            t[N] = (u64)s;                                                                      \
            t[N + 1] = (u64)(s >> 64);                                                          \
            u64 m = t[0] * (u64)(INV);                                                          \
            u128 p = (u128)m * MOD[0] + t[0];                                                   \
            c = (u64)(p >> 64);                                                                 \
            for (int j = 1; j < N; j++) {                                                       \
                p = (u128)m * MOD[j] + t[j] + c;                                                \
                t[j - 1] = (u64)p;                                                              \
                c = (u64)(p >> 64);                                                             \
            }                                                                                   \
            s = (u128)t[N] + c;                                                                 \
            t[N - 1] = (u64)s;                                                                  \
            t[N] = t[N + 1] + (u64)(s >> 64);                                                   \
        }                                                                                       \
        if (t[N] || PFX##_geq(t, MOD)) PFX##_sub_nored(t, t, MOD);                              \
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
    for (int s = 0; s < log_n; s++) {
        size_t m = (size_t)1 << s;
        size_t tw_stride = n >> (s + 1);
        /* butterfly q of this stage: block k = (q / m) * 2m, offset j = q % m */
#pragma omp parallel for schedule(static) if (par) num_threads(nthreads)
        for (long q = 0; q < (long)(n / 2); q++) {
            size_t j = (size_t)q & (m - 1);
            size_t k = (((size_t)q) >> s) << (s + 1);
            u64 t[4], u[4];
            fr_mul(t, a + 4 * (k + j + m), roots + 4 * (j * tw_stride));
            memcpy(u, a + 4 * (k + j), 32);
            fr_add(a + 4 * (k + j), u, t);
            fr_sub(a + 4 * (k + j + m), u, t);
        }
    }
}

void oracle_ntt(u64 *data, int log_n, size_t batch, size_t batch_stride_elems, int direction,
                int coset, int nthreads) {
    size_t n = (size_t)1 << log_n;
#ifdef _OPENMP
    if (nthreads <= 0) nthreads = omp_get_max_threads();
#else
    nthreads = 1;
#endif
    u64 gen[4];
    domain_group_gen(gen, log_n, direction);
    size_t half = n > 1 ? n / 2 : 1;
    u64 *roots = (u64 *)malloc(32 * half);
    memcpy(roots, FR_ONE, 32);
    for (size_t i = 1; i < half; i++) fr_mul(roots + 4 * i, roots + 4 * (i - 1), gen);
    u64 size_inv[4], nn[4] = {n, 0, 0, 0};
    fr_mul(nn, nn, FR_R2);
    fr_inv(size_inv, nn, FR_ONE);
    /* coset powers g^i (forward: before the FFT) or g^-i (inverse: after the iFFT) */
    u64 *cpow = NULL;
    if (coset) {
        cpow = (u64 *)malloc(32 * n);
        memcpy(cpow, FR_ONE, 32);
        const u64 *g = direction ? FR_GENERATOR_INV : FR_GENERATOR;
        for (size_t i = 1; i < n; i++) fr_mul(cpow + 4 * i, cpow + 4 * (i - 1), g);
    }
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
    free(roots);
    if (cpow) free(cpow);
    (void)fr_pow_u64;
}

int oracle_num_threads(void) {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}
