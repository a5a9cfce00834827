// field.cuh -- Montgomery arithmetic for BLS12-377 Fr (8 x 32-bit limbs) and Fq (12 x 32-bit limbs).
//
// Replaces (on the GPU) snarkVM's Fp256<FrParameters> / Fp384<FqParameters> mul_assign,
// square_in_place, add/sub/neg  [UPSTREAM snarkvm-fields 1.0.0 @ dea322b: fields/src/fp_256.rs,
// fp_384.rs; SURVEY.md section 8a row a7].  Same values, different limb width: a snarkVM u64 limb is
// two consecutive u32 limbs here (little endian), so the byte image of an element is identical and
// buffers cross the C ABI unchanged.  Every function returns a FULLY REDUCED value in [0, m).
//
// Multiplication: word-serial Montgomery with the partial products split into an "even" and an "odd"
// accumulator so that every {lo,hi} product pair lands on a 64-bit aligned register pair and the
// whole row is one carry chain of IMAD.WIDE.U32.X.  Because -m^-1 mod 2^32 = 0xffffffff for both
// moduli, the per-row quotient digit is just the negation of the low limb (mul by 0xffffffff).
// After each row the value is divided by 2^32, which swaps the roles of the two accumulators; the
// single limb that falls out of alignment is added with add.cc and its carry is consumed by the
// next row's first chain, so nothing ever ripples.  Cost: 2*N^2 + O(N) multiply-adds.
#pragma once
#include "bls12_377_constants.h"
#include "ptx_ops.cuh"

#ifdef __CUDACC__
#define B200_UNROLL _Pragma("unroll")
#else
#define B200_UNROLL
#endif

// ---------------------------------------------------------------------------------------------
// Field parameter packs (limb i of each constant as an immediate after unrolling)
// ---------------------------------------------------------------------------------------------
struct FrP {
    static constexpr int N = FR_LIMBS;
    static constexpr int MODBITS = 253;
    B200_HDM static uint32_t mod(int i) {
        const uint32_t m[8] = {0x00000001u, 0x0a118000u, 0xd0000001u, 0x59aa76feu,
                               0x5c37b001u, 0x60b44d1eu, 0x9a2ca556u, 0x12ab655eu};
        return m[i];
    }
    B200_HDM static uint32_t one(int i) {
        const uint32_t m[8] = {0xfffffff3u, 0x7d1c7fffu, 0x6ffffff2u, 0x7257f50fu,
                               0x512c0feeu, 0x16d81575u, 0x2bbb9a9du, 0x0d4bda32u};
        return m[i];
    }
    B200_HDM static uint32_t r2(int i) {
        const uint32_t m[8] = {0xb861857bu, 0x25d577bau, 0x8860591fu, 0xcc2c27b5u,
                               0xe5dc8593u, 0xa7cc008fu, 0xeff1c939u, 0x011fdae7u};
        return m[i];
    }
};

struct FqP {
    static constexpr int N = FQ_LIMBS;
    static constexpr int MODBITS = 377;
    B200_HDM static uint32_t mod(int i) {
        const uint32_t m[12] = {0x00000001u, 0x8508c000u, 0x30000000u, 0x170b5d44u,
                                0xba094800u, 0x1ef3622fu, 0x00f5138fu, 0x1a22d9f3u,
                                0x6ca1493bu, 0xc63b05c0u, 0x17c510eau, 0x01ae3a46u};
        return m[i];
    }
    B200_HDM static uint32_t one(int i) {
        const uint32_t m[12] = {0xffffff68u, 0x02cdffffu, 0x7fffffb1u, 0x51409f83u,
                                0x8a7d3ff2u, 0x9f7db3a9u, 0x6e7c6305u, 0x7b4e97b7u,
                                0x803c84e8u, 0x4cf495bfu, 0xe2fdf49au, 0x008d6661u};
        return m[i];
    }
    B200_HDM static uint32_t r2(int i) {
        const uint32_t m[12] = {0x9400cd22u, 0xb786686cu, 0xb00431b1u, 0x0329fcaau,
                                0x62d6b46du, 0x22a5f111u, 0x827dc3acu, 0xbfdf7d03u,
                                0x41790bf9u, 0x837e92f0u, 0x1e914b88u, 0x006dfccbu};
        return m[i];
    }
};

template <class P>
struct Fp {
    uint32_t v[P::N];
};
typedef Fp<FrP> fr_t;
typedef Fp<FqP> fq_t;

// ---------------------------------------------------------------------------------------------
// basic predicates / constants
// ---------------------------------------------------------------------------------------------
template <class P> B200_HD Fp<P> fp_zero() {
    Fp<P> r;
    B200_UNROLL
    for (int i = 0; i < P::N; i++) r.v[i] = 0;
    return r;
}
template <class P> B200_HD Fp<P> fp_one() {
    Fp<P> r;
    B200_UNROLL
    for (int i = 0; i < P::N; i++) r.v[i] = P::one(i);
    return r;
}
template <class P> B200_HD Fp<P> fp_r2() {
    Fp<P> r;
    B200_UNROLL
    for (int i = 0; i < P::N; i++) r.v[i] = P::r2(i);
    return r;
}
template <class P> B200_HD bool fp_is_zero(const Fp<P>& a) {
    uint32_t t = 0;
    B200_UNROLL
    for (int i = 0; i < P::N; i++) t |= a.v[i];
    return t == 0;
}
template <class P> B200_HD bool fp_eq(const Fp<P>& a, const Fp<P>& b) {
    uint32_t t = 0;
    B200_UNROLL
    for (int i = 0; i < P::N; i++) t |= a.v[i] ^ b.v[i];
    return t == 0;
}

// r = (t >= m) ? t - m : t      for t < 2m
template <class P> B200_HD void fp_reduce_once(Fp<P>& t) {
    uint32_t d[P::N];
    d[0] = ptx::sub_cc(t.v[0], P::mod(0));
    B200_UNROLL
    for (int i = 1; i < P::N; i++) d[i] = ptx::subc_cc(t.v[i], P::mod(i));
    uint32_t borrow = ptx::subc(0u, 0u);          // 0 or 0xffffffff
    B200_UNROLL
    for (int i = 0; i < P::N; i++) t.v[i] = borrow ? t.v[i] : d[i];
}

template <class P> B200_HD Fp<P> fp_add(const Fp<P>& a, const Fp<P>& b) {
    Fp<P> r;
    r.v[0] = ptx::add_cc(a.v[0], b.v[0]);
    B200_UNROLL
    for (int i = 1; i < P::N - 1; i++) r.v[i] = ptx::addc_cc(a.v[i], b.v[i]);
    r.v[P::N - 1] = ptx::addc(a.v[P::N - 1], b.v[P::N - 1]);   // 2m < 2^(32N): no carry out
    fp_reduce_once(r);
    return r;
}

template <class P> B200_HD Fp<P> fp_sub(const Fp<P>& a, const Fp<P>& b) {
    Fp<P> r;
    r.v[0] = ptx::sub_cc(a.v[0], b.v[0]);
    B200_UNROLL
    for (int i = 1; i < P::N; i++) r.v[i] = ptx::subc_cc(a.v[i], b.v[i]);
    uint32_t mask = ptx::subc(0u, 0u);             // 0xffffffff iff a < b
    r.v[0] = ptx::add_cc(r.v[0], P::mod(0) & mask);
    B200_UNROLL
    for (int i = 1; i < P::N - 1; i++) r.v[i] = ptx::addc_cc(r.v[i], P::mod(i) & mask);
    r.v[P::N - 1] = ptx::addc(r.v[P::N - 1], P::mod(P::N - 1) & mask);
    return r;
}

template <class P> B200_HD Fp<P> fp_neg(const Fp<P>& a) {
    Fp<P> r;
    uint32_t nz = 0;
    B200_UNROLL
    for (int i = 0; i < P::N; i++) nz |= a.v[i];
    r.v[0] = ptx::sub_cc(P::mod(0), a.v[0]);
    B200_UNROLL
    for (int i = 1; i < P::N - 1; i++) r.v[i] = ptx::subc_cc(P::mod(i), a.v[i]);
    r.v[P::N - 1] = ptx::subc(P::mod(P::N - 1), a.v[P::N - 1]);
    B200_UNROLL
    for (int i = 0; i < P::N; i++) r.v[i] = nz ? r.v[i] : 0u;
    return r;
}

template <class P> B200_HD Fp<P> fp_dbl(const Fp<P>& a) { return fp_add(a, a); }

template <class P> B200_HD Fp<P> fp_select(bool c, const Fp<P>& a, const Fp<P>& b) {   // c ? a : b
    Fp<P> r;
    B200_UNROLL
    for (int i = 0; i < P::N; i++) r.v[i] = c ? a.v[i] : b.v[i];
    return r;
}

// ---------------------------------------------------------------------------------------------
// Montgomery product
// ---------------------------------------------------------------------------------------------
// One row:  acc += v * s  on limbs of `v` with parity `par` (0: even limbs into an even-aligned
// accumulator, 1: odd limbs into an odd-aligned accumulator).  The accumulator index k holds the
// product pair of v[k + par].  `first` selects whether the chain starts with carry-in.
namespace detail {

// acc[k], acc[k+1] += lo/hi(v[k + par] * s), k = 0, 2, ..., N-2; chain starts WITHOUT carry-in.
template <class P, int PAR, class V>
B200_HD void row_mad(uint32_t* acc, V v, uint32_t s) {
    B200_UNROLL
    for (int k = 0; k < P::N; k += 2) {
        acc[k] = (k == 0) ? ptx::mad_lo_cc(v(k + PAR), s, acc[k]) : ptx::madc_lo_cc(v(k + PAR), s, acc[k]);
        acc[k + 1] = ptx::madc_hi_cc(v(k + PAR), s, acc[k + 1]);
    }
}

// acc += mod * m on the EVEN limbs of the modulus, exploiting mod[0] == 1 (both moduli are 1 mod 2^32): the first
// product pair is m * 1 = {m, 0}, i.e. two additions instead of a wide multiply (one multiplier slot saved per row:
// 4 % of an Fq product, 6 % of an Fr product -- the multiplier pipe is the bound, adds ride along for free).
template <class P>
B200_HD void row_mad_mod_even(uint32_t* acc, uint32_t m) {
    acc[0] = ptx::add_cc(acc[0], m);
    acc[1] = ptx::addc_cc(acc[1], 0u);
    B200_UNROLL
    for (int k = 2; k < P::N; k += 2) {
        acc[k] = ptx::madc_lo_cc(P::mod(k), m, acc[k]);
        acc[k + 1] = ptx::madc_hi_cc(P::mod(k), m, acc[k + 1]);
    }
}

// acc[k], acc[k+1] = lo/hi(v[k + 1] * s) + acc[k+2], acc[k+3]   (odd limbs, WITH carry-in, and the
// accumulator slides down by two limbs = the division by 2^64 of the array that was "even").
// acc has N + 1 limbs on entry (acc[N] = carry limb); acc[N - 1] receives the last high word.
template <class P, class V>
B200_HD void row_mad_shift(uint32_t* acc, V v, uint32_t s) {
    B200_UNROLL
    for (int k = 0; k < P::N; k += 2) {
        acc[k] = ptx::madc_lo_cc(v(k + 1), s, acc[k + 2]);
        if (k + 3 <= P::N) acc[k + 1] = ptx::madc_hi_cc(v(k + 1), s, acc[k + 3]);
        else               acc[k + 1] = ptx::madc_hi(v(k + 1), s, 0u);   // top word: cannot carry out
    }
}

}  // namespace detail

template <class P> B200_HD Fp<P> fp_mul_cc(const Fp<P>& a, const Fp<P>& b) {
    constexpr int N = P::N;
    // X, Y: the two accumulators.  Exactly one of them is "even" (value at limb positions 0..N,
    // N + 1 limbs incl. carry limb) and the other "odd" (positions 1..N, N limbs) at any time.
    uint32_t X[N + 1], Y[N + 1];
    auto A = [&](int i) { return a.v[i]; };
    auto M = [&](int i) { return P::mod(i); };

    // ---- row 0: plain products
    B200_UNROLL
    for (int k = 0; k < N; k += 2) {
        X[k] = ptx::mul_lo(a.v[k], b.v[0]);
        X[k + 1] = ptx::mul_hi(a.v[k], b.v[0]);
        Y[k] = ptx::mul_lo(a.v[k + 1], b.v[0]);
        Y[k + 1] = ptx::mul_hi(a.v[k + 1], b.v[0]);
    }
    X[N] = 0;
    Y[N] = 0;
    {
        // quotient digit X[0] * (-mod^-1) = -X[0] since -mod^-1 = 0xffffffff.  Written as sub.cc on purpose:
        // a plain negation is folded by ptxas into a negated IMAD operand, which blocks IMAD.WIDE fusion.
        uint32_t m = ptx::sub_cc(0u, X[0]);
        detail::row_mad<P, 1>(Y, M, m);               // odd chain: no carry out (value bound)
        detail::row_mad_mod_even<P>(X, m);
        X[N] = ptx::addc(X[N], 0u);
    }
    // ---- rows 1 .. N-1; E = even-role array entering the row (E[0] == 0), O = odd-role array
    B200_UNROLL
    for (int i = 1; i < N; i++) {
        uint32_t* E = (i & 1) ? X : Y;
        uint32_t* O = (i & 1) ? Y : X;
        // divide by 2^32: O becomes even, E (from limb 2 up) becomes odd; limb E[1] is the
        // unaligned straggler: add it at position 0 and let its carry enter the odd chain.
        O[0] = ptx::add_cc(O[0], E[1]);
        detail::row_mad_shift<P>(E, A, b.v[i]);       // E[k] <- E[k+2] + a_odd * b_i   (+carry in)
        detail::row_mad<P, 0>(O, A, b.v[i]);          // O += a_even * b_i
        O[N] = ptx::addc(0u, 0u);
        uint32_t m = ptx::sub_cc(0u, O[0]);
        detail::row_mad<P, 1>(E, M, m);
        detail::row_mad_mod_even<P>(O, m);
        O[N] = ptx::addc(O[N], 0u);
    }
    // after row N-1 (odd index, N even): even-role array is Y... for i = N-1 odd: E = X, O = Y and
    // O ended as the even-role array with O[0] == 0.  Result = (O >> 32) + E.
    uint32_t* Ev = ((N - 1) & 1) ? Y : X;
    uint32_t* Od = ((N - 1) & 1) ? X : Y;
    Fp<P> r;
    r.v[0] = ptx::add_cc(Od[0], Ev[1]);
    B200_UNROLL
    for (int k = 1; k < N - 1; k++) r.v[k] = ptx::addc_cc(Od[k], Ev[k + 1]);
    r.v[N - 1] = ptx::addc(Od[N - 1], Ev[N]);
    fp_reduce_once(r);
    return r;
}

// Measured alternatives (profiles/r01_microbench_int_pipe.json, B200 @1965 MHz, per SM per clock):
//   IMAD (lo) 63.5 | IMAD.HI 27.3 | IMAD.WIDE 31.5 | IMAD.WIDE + 64-bit addend 25.2 | IADD3 126 | DFMA 58.4
// i.e. a full 32x32->64 product costs two multiplier passes however it is written, so the 2*N^2 wide
// multiply-adds of this routine ARE the integer-multiply roofline: 288 / 31.5 = 9.14 clk per Fq product
// (measured 9.33 = 98 %), 128 / 31.5 = 4.06 clk per Fr product (measured 4.19 = 97 %).  A carry-free
// variant on 29-bit limbs (13 / 9 limbs, plain IMAD.WIDE into 64-bit columns) was built and measured at
// 0.68x / 0.61x of this one (more products, and ptxas splits every accumulate into IMAD.WIDE + 2 IADD3),
// and an FP64-pipe variant does not pay either (3 FP64 ops + 4 integer adds per 52-bit product makes it
// issue-bound), so both were dropped.
template <class P> B200_HD Fp<P> fp_mul(const Fp<P>& a, const Fp<P>& b) { return fp_mul_cc(a, b); }

// ---------------------------------------------------------------------------------------------
// Separated product / reduction ("SOS"): full 2N-limb products and a stand-alone Montgomery reduction.
// MEASURED AND NOT USED ON THE HOT PATHS (kept as tested building blocks, tests/test_host_arith.py): in isolation the
// XYZZ mixed-add loop gains 4.7 % (2.77 -> 2.90 G adds/s), but the 2N-limb temporaries push msm_accumulate_kernel from
// 167 to 212 registers (12 -> 8 resident warps per SM) and its time is unchanged (84.0 vs 84.1 ms at 2^24); forced
// back to 168 registers it spills and slows to 92 ms.  fp_sqr therefore stays the interleaved product.
// Same multiplier cost as the interleaved routine for a plain product (N^2 + N(N-1) wide multiply-adds), but it
// makes two savings possible where the multiplier pipe is the bound (the MSM):
//   * squaring: the off-diagonal products are computed once and doubled -> N(N+1)/2 instead of N^2 products
//     (Fq: 78 + 132 = 210 instead of 276 multiplier slots);
//   * a*b +- c*d with ONE reduction (the Y3 = R*(Q - X3) - Y1*PPP of every point addition): 2*144 + 132 = 420
//     instead of 552 slots.
// The 2N-limb values are accumulated in two arrays by the parity of the position where a {lo,hi} product pair
// starts, so that again every pair sits on an aligned register pair and every row is one IMAD.WIDE.U32.X carry
// chain; the carry leaving a chain lands on a limb that so far holds at most another chain's carry, so one addc
// suffices and nothing ripples.
// ---------------------------------------------------------------------------------------------
namespace detail {

// T[0 .. 2N) = P0 + P1
template <int N2> B200_HD void wide_merge(uint32_t* T, const uint32_t* P0, const uint32_t* P1) {
    T[0] = ptx::add_cc(P0[0], P1[0]);
    B200_UNROLL
    for (int k = 1; k < N2 - 1; k++) T[k] = ptx::addc_cc(P0[k], P1[k]);
    T[N2 - 1] = ptx::addc(P0[N2 - 1], P1[N2 - 1]);
}

}  // namespace detail

template <class P> B200_HD void fp_mul_wide(uint32_t* T, const Fp<P>& a, const Fp<P>& b) {
    constexpr int N = P::N;
    uint32_t P0[2 * N + 2], P1[2 * N + 2];
    B200_UNROLL
    for (int k = 0; k < 2 * N + 2; k++) { P0[k] = 0; P1[k] = 0; }
    B200_UNROLL
    for (int i = 0; i < N; i++) {
        uint32_t* E = (i & 1) ? P1 : P0;             // pairs starting at positions of the parity of i
        uint32_t* O = (i & 1) ? P0 : P1;
        B200_UNROLL
        for (int j = 0; j < N; j += 2) {
            E[i + j] = (j == 0) ? ptx::mad_lo_cc(a.v[j], b.v[i], E[i + j]) : ptx::madc_lo_cc(a.v[j], b.v[i], E[i + j]);
            E[i + j + 1] = ptx::madc_hi_cc(a.v[j], b.v[i], E[i + j + 1]);
        }
        E[i + N] = ptx::addc(E[i + N], 0u);
        B200_UNROLL
        for (int j = 1; j < N; j += 2) {
            O[i + j] = (j == 1) ? ptx::mad_lo_cc(a.v[j], b.v[i], O[i + j]) : ptx::madc_lo_cc(a.v[j], b.v[i], O[i + j]);
            O[i + j + 1] = ptx::madc_hi_cc(a.v[j], b.v[i], O[i + j + 1]);
        }
        O[i + N + 1] = ptx::addc(O[i + N + 1], 0u);
    }
    detail::wide_merge<2 * N>(T, P0, P1);
}

template <class P> B200_HD void fp_sqr_wide(uint32_t* T, const Fp<P>& a) {
    constexpr int N = P::N;
    uint32_t P0[2 * N + 2], P1[2 * N + 2];
    B200_UNROLL
    for (int k = 0; k < 2 * N + 2; k++) { P0[k] = 0; P1[k] = 0; }
    // off-diagonal products a[i] * a[j], i < j, once
    B200_UNROLL
    for (int i = 0; i < N - 1; i++) {
        B200_UNROLL
        for (int q = 0; q < 2; q++) {
            const int j0 = i + 1 + (((i + 1) & 1) != q ? 1 : 0);      // first j > i with j % 2 == q
            if (j0 <= N - 1) {
                uint32_t* A = ((i + j0) & 1) ? P1 : P0;
                int last = j0;
                B200_UNROLL
                for (int j = j0; j < N; j += 2) {
                    A[i + j] = (j == j0) ? ptx::mad_lo_cc(a.v[j], a.v[i], A[i + j]) : ptx::madc_lo_cc(a.v[j], a.v[i], A[i + j]);
                    A[i + j + 1] = ptx::madc_hi_cc(a.v[j], a.v[i], A[i + j + 1]);
                    last = j;
                }
                A[i + last + 2] = ptx::addc(A[i + last + 2], 0u);
            }
        }
    }
    uint32_t S[2 * N];
    detail::wide_merge<2 * N>(S, P0, P1);
    // T = 2 * S + sum_i a[i]^2 * 2^(64 i)
    B200_UNROLL
    for (int k = 2 * N - 1; k >= 1; k--) T[k] = (S[k] << 1) | (S[k - 1] >> 31);
    T[0] = S[0] << 1;
    B200_UNROLL
    for (int i = 0; i < N; i++) {
        T[2 * i] = (i == 0) ? ptx::mad_lo_cc(a.v[i], a.v[i], T[2 * i]) : ptx::madc_lo_cc(a.v[i], a.v[i], T[2 * i]);
        T[2 * i + 1] = ptx::madc_hi_cc(a.v[i], a.v[i], T[2 * i + 1]);
    }
}

// T (2N limbs) += U (2N limbs); the caller guarantees no overflow (sums of a few products of reduced values)
template <class P> B200_HD void fp_wide_add(uint32_t* T, const uint32_t* U) {
    constexpr int N2 = 2 * P::N;
    T[0] = ptx::add_cc(T[0], U[0]);
    B200_UNROLL
    for (int k = 1; k < N2 - 1; k++) T[k] = ptx::addc_cc(T[k], U[k]);
    T[N2 - 1] = ptx::addc(T[N2 - 1], U[N2 - 1]);
}

// Montgomery reduction T / 2^(32N) mod m of a 2N-limb value T < m * 2^(32N) (fully reduced result).
// Word-serial like fp_mul_cc with the same even/odd accumulators; the a*b row is replaced by the division's shift,
// which also brings in the next high limb of T.
template <class P> B200_HD Fp<P> fp_redc_wide(const uint32_t* T) {
    constexpr int N = P::N;
    uint32_t X[N + 1], Y[N + 1];
    auto M = [&](int i) { return P::mod(i); };
    B200_UNROLL
    for (int k = 0; k < N; k++) { X[k] = T[k]; Y[k] = 0; }
    X[N] = 0;
    Y[N] = 0;
    {
        uint32_t m = ptx::sub_cc(0u, X[0]);
        detail::row_mad<P, 1>(Y, M, m);
        detail::row_mad_mod_even<P>(X, m);
        X[N] = ptx::addc(X[N], 0u);
    }
    B200_UNROLL
    for (int i = 1; i < N; i++) {
        uint32_t* E = (i & 1) ? X : Y;                // even-role array entering the row, E[0] == 0
        uint32_t* O = (i & 1) ? Y : X;
        O[0] = ptx::add_cc(O[0], E[1]);               // straggler limb; its carry rides the shift below
        B200_UNROLL
        for (int k = 0; k < N - 2; k++) E[k] = ptx::addc_cc(E[k + 2], 0u);      // divide by 2^64: E becomes odd-role
        E[N - 2] = ptx::addc_cc(E[N], T[N + i - 1]);  // next high limb of T enters at position N-1
        E[N - 1] = ptx::addc(0u, 0u);
        O[N] = 0;
        uint32_t m = ptx::sub_cc(0u, O[0]);
        detail::row_mad<P, 1>(E, M, m);
        detail::row_mad_mod_even<P>(O, m);
        O[N] = ptx::addc(O[N], 0u);
    }
    uint32_t* Ev = ((N - 1) & 1) ? Y : X;
    uint32_t* Od = ((N - 1) & 1) ? X : Y;
    Fp<P> r;
    r.v[0] = ptx::add_cc(Od[0], Ev[1]);
    B200_UNROLL
    for (int k = 1; k < N - 1; k++) r.v[k] = ptx::addc_cc(Od[k], Ev[k + 1]);
    r.v[N - 1] = ptx::addc(Od[N - 1], Ev[N]);
    r.v[N - 1] += T[2 * N - 1];                       // last high limb; the total is < 2m, so this cannot overflow
    fp_reduce_once(r);
    return r;
}

template <class P> B200_HD Fp<P> fp_sqr_sos(const Fp<P>& a) {
    uint32_t T[2 * P::N];
    fp_sqr_wide<P>(T, a);
    return fp_redc_wide<P>(T);
}
template <class P> B200_HD Fp<P> fp_sqr(const Fp<P>& a) { return fp_mul_cc(a, a); }

// a*b - c*d with a single reduction
template <class P> B200_HD Fp<P> fp_mul_mul_sub(const Fp<P>& a, const Fp<P>& b, const Fp<P>& c, const Fp<P>& d) {
    uint32_t T[2 * P::N], U[2 * P::N];
    fp_mul_wide<P>(T, a, b);
    fp_mul_wide<P>(U, fp_neg(c), d);                  // (m - c) * d = -c*d (mod m), both addends non-negative
    fp_wide_add<P>(T, U);
    return fp_redc_wide<P>(T);
}

// canonical <-> Montgomery
template <class P> B200_HD Fp<P> fp_to_mont(const Fp<P>& a) { return fp_mul(a, fp_r2<P>()); }
template <class P> B200_HD Fp<P> fp_from_mont(const Fp<P>& a) {
    Fp<P> one = fp_zero<P>();
    one.v[0] = 1;
    return fp_mul(a, one);
}

// a^e for a 64-bit exponent (left-to-right square and multiply)
template <class P> B200_HD Fp<P> fp_pow_u64(const Fp<P>& a, uint64_t e) {
    Fp<P> acc = fp_one<P>();
    bool started = false;
    for (int i = 63; i >= 0; i--) {
        if (started) acc = fp_sqr(acc);
        if ((e >> i) & 1) {
            acc = started ? fp_mul(acc, a) : a;
            started = true;
        }
    }
    return acc;
}

// Fermat inverse a^(m-2); 0 -> 0.  (Utility paths only: never on the MSM / NTT inner loops.)
template <class P> B200_HD Fp<P> fp_inv(const Fp<P>& a) {
    uint32_t e[P::N];
    e[0] = ptx::sub_cc(P::mod(0), 2u);
    B200_UNROLL
    for (int i = 1; i < P::N; i++) e[i] = ptx::subc_cc(P::mod(i), 0u);
    Fp<P> acc = fp_one<P>();
    bool started = false;
    for (int i = 32 * P::N - 1; i >= 0; i--) {
        if (started) acc = fp_sqr(acc);
        if ((e[i >> 5] >> (i & 31)) & 1) {
            acc = started ? fp_mul(acc, a) : a;
            started = true;
        }
    }
    return acc;
}

// Inverse by the binary extended Euclid (right-shift form, v kept odd): additions, shifts and selects only -- no
// multiplier pipe, and a dependent chain ~4x shorter than the ~570 products of the Fermat ladder (0.52 ms per batch
// inversion tail on one warp).  Invariant x1 * a = u * C, x2 * a = v * C (mod m) with C = R^2, so for a Montgomery
// input a = A R the result x2 = C / a = A^-1 R is the Montgomery inverse with no extra product.  0 -> 0.
// At most bits(a) + bits(m) iterations; every iteration is the same instruction stream for all lanes (predicated).
template <class P> B200_HD Fp<P> fp_inv_gcd(const Fp<P>& a) {
    constexpr int N = P::N;
    uint32_t u[N], v[N], x1[N], x2[N];
    uint32_t nz = 0;
    B200_UNROLL
    for (int i = 0; i < N; i++) {
        u[i] = a.v[i];
        v[i] = P::mod(i);
        x1[i] = P::r2(i);
        x2[i] = 0;
        nz |= u[i];
    }
    while (nz) {
        const bool odd = u[0] & 1u;
        if (odd) {
            // lt = (u < v): borrow of u - v
            uint32_t d[N];
            d[0] = ptx::sub_cc(u[0], v[0]);
            B200_UNROLL
            for (int i = 1; i < N; i++) d[i] = ptx::subc_cc(u[i], v[i]);
            const uint32_t lt = ptx::subc(0u, 0u);                 // 0xffffffff iff u < v
            if (lt) {
                // (u, v) <- (v - u, u);  (x1, x2) <- (x2 - x1, x1)
                uint32_t t[N];
                t[0] = ptx::sub_cc(v[0], u[0]);
                B200_UNROLL
                for (int i = 1; i < N - 1; i++) t[i] = ptx::subc_cc(v[i], u[i]);
                t[N - 1] = ptx::subc(v[N - 1], u[N - 1]);
                B200_UNROLL
                for (int i = 0; i < N; i++) { v[i] = u[i]; u[i] = t[i]; }
                B200_UNROLL
                for (int i = 0; i < N; i++) { const uint32_t s = x1[i]; x1[i] = x2[i]; x2[i] = s; }
            } else {
                B200_UNROLL
                for (int i = 0; i < N; i++) u[i] = d[i];
            }
            // x1 <- x1 - x2 mod m   (after the swap above x2 is the old x1)
            x1[0] = ptx::sub_cc(x1[0], x2[0]);
            B200_UNROLL
            for (int i = 1; i < N; i++) x1[i] = ptx::subc_cc(x1[i], x2[i]);
            const uint32_t mask = ptx::subc(0u, 0u);
            x1[0] = ptx::add_cc(x1[0], P::mod(0) & mask);
            B200_UNROLL
            for (int i = 1; i < N - 1; i++) x1[i] = ptx::addc_cc(x1[i], P::mod(i) & mask);
            x1[N - 1] = ptx::addc(x1[N - 1], P::mod(N - 1) & mask);
        }
        // u is even now: u <- u / 2, x1 <- x1 / 2 mod m
        B200_UNROLL
        for (int i = 0; i < N - 1; i++) u[i] = (u[i] >> 1) | (u[i + 1] << 31);
        u[N - 1] >>= 1;
        const uint32_t addm = (x1[0] & 1u) ? 0xffffffffu : 0u;
        x1[0] = ptx::add_cc(x1[0], P::mod(0) & addm);
        B200_UNROLL
        for (int i = 1; i < N - 1; i++) x1[i] = ptx::addc_cc(x1[i], P::mod(i) & addm);
        x1[N - 1] = ptx::addc(x1[N - 1], P::mod(N - 1) & addm);   // < 2m < 2^(32N): no carry out
        B200_UNROLL
        for (int i = 0; i < N - 1; i++) x1[i] = (x1[i] >> 1) | (x1[i + 1] << 31);
        x1[N - 1] >>= 1;
        nz = 0;
        B200_UNROLL
        for (int i = 0; i < N; i++) nz |= u[i];
    }
    Fp<P> r;
    B200_UNROLL
    for (int i = 0; i < N; i++) r.v[i] = x2[i];
    return r;
}
