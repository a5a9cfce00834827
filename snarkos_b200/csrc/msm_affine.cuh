// msm_affine.cuh -- bucket accumulation by rounds of pairwise AFFINE additions sharing one inversion.
//
// snarkVM's own bucket method for BLS12-377 does this on the CPU: `batched::batch_add` adds the points of a bucket
// pair by pair in affine coordinates with ONE batch inversion per round
// [UPSTREAM snarkvm-algorithms 1.0.0 @ dea322b: algorithms/src/msm/variable_base/batched.rs; SURVEY.md 8a row a2].
// An affine addition given 1/(x2 - x1) costs 2M + 1S, and Montgomery's trick prices the inversion at 3M per
// addition, so a bucket addition costs ~6.3 Fq products instead of the 10 of an XYZZ mixed addition -- the only
// lever left on a kernel that already runs at 81 % of the integer-multiply roofline.
//
// One round turns the sorted per-bucket point lists into lists of half the length:
//     out[noff[b] + j] = in[off[b] + 2j] + in[off[b] + 2j + 1]          (the odd one out is copied through)
// in three steps, all flat over the OUTPUT positions q (equal work per lane whatever the bucket sizes are):
//     denoms : every thread walks MSM_PAIRS_PER_THREAD consecutive outputs, d_q = x2 - x1 (2 y1 for a doubling, 1 for
//              the exceptional cases), stores the running product BEFORE q (pre[q]) and its total (partial[t])
//     invert : partial[] inverted as a whole (recursive Montgomery trick, Fermat only on the last few hundred values)
//     add    : every thread walks its outputs backwards: 1/d_q = I * pre[q]; I *= d_q; lambda = num / d_q;
//              x3 = lambda^2 - x1 - x2;  y3 = lambda (x1 - x3) - y1
// Round 0 gathers its operands from the packed bases through the sorted entry list (index | sign << 31); later
// rounds read the previous round's list.  After a few rounds (lists of 2..4 points left) the XYZZ chunk kernel of
// msm.cu finishes the buckets -- it also absorbs every degenerate distribution, so the rounds need no termination test.
//
// Exceptional cases (the affine formulas are as incomplete as the XYZZ ones): an operand at infinity ((0, 0)) copies
// the other; x1 == x2 is a doubling when y1 == y2 != 0 and the point at infinity otherwise.  Their denominator is 1,
// so the running products never vanish.  Both walks classify from the same loads, hence identically.
//
// Every function is B200_HD: the CUDA kernels in msm.cu call them per thread, tests/host/msm_host.cpp calls the very
// same functions on the CPU (PTX emulated bit-exactly) to check the whole schedule against the oracle without a GPU.
#pragma once
#include "msm_core.cuh"

#ifndef MSM_PAIRS_PER_THREAD
#define MSM_PAIRS_PER_THREAD 32u
#endif

struct PairRound {
    const uint4* src;            // round 0: packed bases (128-byte records x | y | pad); later: x list of the previous round
    const uint4* src_y;          // later rounds: y list (lists are two planes of 48-byte coordinates, so that the
                                 // denominator pass reads x only); unused in round 0
    const uint32_t* entries;     // round 0: sorted (index | sign << 31); nullptr afterwards
    const uint32_t* off;         // K + 1 offsets of the input lists
    const uint32_t* noff;        // K + 1 offsets of the output lists, noff[b+1] - noff[b] = ceil(count_b / 2)
    uint32_t K;
};

enum { PAIR_ADD = 0, PAIR_DBL = 1, PAIR_COPY_A = 2, PAIR_COPY_B = 3, PAIR_INF = 4 };

// bucket of output position q: the b with noff[b] <= q < noff[b + 1]   (q < noff[K])
B200_HD uint32_t pair_locate(const uint32_t* noff, uint32_t K, uint32_t q) {
    uint32_t lo = 0, hi = K;
    while (hi - lo > 1) {
        uint32_t mid = (lo + hi) >> 1;
        if (noff[mid] <= q) lo = mid; else hi = mid;
    }
    return lo;
}

B200_HD fq_t pair_load_fq(const uint4* w) { return fq_from_u4x3(w); }

// where the coordinates of input slot i live (and the sign its y takes)
struct PairSlot {
    const uint4* x;
    const uint4* y;
    uint32_t neg;
};
B200_HD PairSlot pair_slot(const PairRound& rd, uint32_t i) {
    PairSlot s;
    if (rd.entries) {
        const uint32_t id = rd.entries[i];
        s.neg = id >> 31;
        s.x = rd.src + (size_t)(id & 0x7fffffffu) * G1_BASE_U4;
        s.y = s.x + 3;
    } else {
        s.neg = 0;
        s.x = rd.src + 3 * (size_t)i;
        s.y = rd.src_y + 3 * (size_t)i;
    }
    return s;
}
B200_HD fq_t pair_slot_y(const PairSlot& s) {
    fq_t y = pair_load_fq(s.y);
    return s.neg ? fp_neg(y) : y;
}

// both coordinates of a slot, sign applied (base records: three 256-bit loads)
B200_HD void pair_load_point(const PairRound& rd, const PairSlot& s, fq_t& x, fq_t& y) {
    if (rd.entries) {
        const g1_affine_t a = g1_load_base(s.x);
        x = a.x;
        y = a.y;
    } else {
        x = pair_load_fq(s.x);
        y = pair_load_fq(s.y);
    }
    if (s.neg) y = fp_neg(y);
}

// classification of the pair (A, B) and its denominator; y's already carry their signs
B200_HD int pair_classify(const fq_t& x1, const fq_t& y1, const fq_t& x2, const fq_t& y2, bool has2) {
    if (!has2) return PAIR_COPY_A;
    if (fp_is_zero(x1) && fp_is_zero(y1)) return PAIR_COPY_B;
    if (fp_is_zero(x2) && fp_is_zero(y2)) return PAIR_COPY_A;
    if (fp_eq(x1, x2)) return (fp_eq(y1, y2) && !fp_is_zero(y1)) ? PAIR_DBL : PAIR_INF;
    return PAIR_ADD;
}

// forward walk over the outputs of one thread: input slot pair of output q (buckets advance monotonically)
struct PairWalk {
    uint32_t b, nb, ne, ob, oe;
};
B200_HD void pair_walk_to(const PairRound& rd, PairWalk& w, uint32_t q, uint32_t& i0, bool& has2) {
    while (q >= w.ne) {                                            // next non-empty bucket
        ++w.b;
        w.nb = w.ne;
        w.ne = rd.noff[w.b + 1];
        w.ob = w.oe;
        w.oe = rd.off[w.b + 1];
    }
    i0 = w.ob + 2 * (q - w.nb);
    has2 = i0 + 1 < w.oe;
}

// step 1: denominators and their running products over the outputs [t * PPT, (t + 1) * PPT).
// Measured (ncu, 2^24, round 0): a stream of random gathers with one product each, limited by the gather rate of
// the memory system (~25 G lines/s, DRAM 60 % busy) -- register prefetch of the next pair (107 registers) changes
// nothing, L2 prefetch three outputs ahead thrashes L2 and costs 50 %; the plain loop is the fastest form.
B200_HD void pair_denoms_thread(const PairRound& rd, uint32_t t, uint4* pre, uint4* partial) {
    const uint32_t total = rd.noff[rd.K];
    const unsigned long long q0l = (unsigned long long)t * MSM_PAIRS_PER_THREAD;
    fq_t acc = fp_one<FqP>();
    if (q0l >= total) {
        fq_to_u4x3(acc, partial + 3 * (size_t)t);                  // the inversion runs over the whole launch bound
        return;
    }
    const uint32_t q0 = (uint32_t)q0l;
    const uint32_t q1 = (q0l + MSM_PAIRS_PER_THREAD < total) ? q0 + MSM_PAIRS_PER_THREAD : total;
    PairWalk w;
    w.b = pair_locate(rd.noff, rd.K, q0);
    w.nb = rd.noff[w.b];
    w.ne = rd.noff[w.b + 1];
    w.ob = rd.off[w.b];
    w.oe = rd.off[w.b + 1];
    for (uint32_t q = q0; q < q1; q++) {
        uint32_t i0;
        bool has2;
        pair_walk_to(rd, w, q, i0, has2);
        fq_to_u4x3(acc, pre + 3 * (size_t)q);
        if (!has2) continue;
        const PairSlot A = pair_slot(rd, i0), B = pair_slot(rd, i0 + 1);
        const fq_t x1 = rd.entries ? g1_load_base_x(A.x) : pair_load_fq(A.x);
        const fq_t x2 = rd.entries ? g1_load_base_x(B.x) : pair_load_fq(B.x);
        fq_t d;
        if (!fp_is_zero(x1) && !fp_is_zero(x2) && !fp_eq(x1, x2)) {
            d = fp_sub(x2, x1);                                    // the only case random inputs ever see
        } else {
            const fq_t y1 = pair_slot_y(A), y2 = pair_slot_y(B);
            const int kind = pair_classify(x1, y1, x2, y2, true);
            if (kind == PAIR_ADD) d = fp_sub(x2, x1);
            else if (kind == PAIR_DBL) d = fp_dbl(y1);
            else continue;
        }
        acc = fp_mul(acc, d);
    }
    fq_to_u4x3(acc, partial + 3 * (size_t)t);
}

// step 3: the additions, walking the same outputs backwards with the inverted thread products
B200_HD void pair_add_thread(const PairRound& rd, uint32_t t, const uint4* pre, const uint4* partial_inv,
                             uint4* out_x, uint4* out_y) {
    const uint32_t total = rd.noff[rd.K];
    const unsigned long long q0l = (unsigned long long)t * MSM_PAIRS_PER_THREAD;
    if (q0l >= total) return;
    const uint32_t q0 = (uint32_t)q0l;
    const uint32_t q1 = (q0l + MSM_PAIRS_PER_THREAD < total) ? q0 + MSM_PAIRS_PER_THREAD : total;
    fq_t I = pair_load_fq(partial_inv + 3 * (size_t)t);
    uint32_t b = pair_locate(rd.noff, rd.K, q1 - 1);
    uint32_t nb = rd.noff[b], ob = rd.off[b], oe = rd.off[b + 1];
    for (uint32_t q = q1; q-- > q0;) {
        while (q < nb) {                                           // previous non-empty bucket
            --b;
            nb = rd.noff[b];
            oe = ob;
            ob = rd.off[b];
        }
        const uint32_t i0 = ob + 2 * (q - nb);
        const bool has2 = i0 + 1 < oe;
        fq_t x1, y1;
        pair_load_point(rd, pair_slot(rd, i0), x1, y1);
        fq_t x2 = x1, y2 = y1;
        if (has2) pair_load_point(rd, pair_slot(rd, i0 + 1), x2, y2);
        const int kind = pair_classify(x1, y1, x2, y2, has2);
        fq_t x3, y3;
        if (kind == PAIR_ADD || kind == PAIR_DBL) {
            fq_t d, num;
            if (kind == PAIR_ADD) {
                d = fp_sub(x2, x1);
                num = fp_sub(y2, y1);
            } else {
                d = fp_dbl(y1);
                const fq_t xx = fp_sqr(x1);
                num = fp_add(fp_dbl(xx), xx);
            }
            const fq_t inv = fp_mul(I, pair_load_fq(pre + 3 * (size_t)q));
            I = fp_mul(I, d);
            const fq_t lambda = fp_mul(num, inv);
            x3 = fp_sub(fp_sub(fp_sqr(lambda), x1), x2);
            y3 = fp_sub(fp_mul(lambda, fp_sub(x1, x3)), y1);
        } else if (kind == PAIR_COPY_A) {
            x3 = x1;
            y3 = y1;
        } else if (kind == PAIR_COPY_B) {
            x3 = x2;
            y3 = y2;
        } else {
            x3 = fp_zero<FqP>();
            y3 = fp_zero<FqP>();
        }
        fq_to_u4x3(x3, out_x + 3 * (size_t)q);
        fq_to_u4x3(y3, out_y + 3 * (size_t)q);
    }
}

// Batch inversion of Fq values that are never zero (products of denominators): chunk products up, inverse walk down.
#define MSM_INV_CHUNK 16u

B200_HD void fq_inv_up_thread(uint4* partial, const uint4* data, size_t n, size_t t) {
    const size_t lo = t * MSM_INV_CHUNK;
    if (lo >= n) return;
    const size_t hi = lo + MSM_INV_CHUNK < n ? lo + MSM_INV_CHUNK : n;
    fq_t acc = pair_load_fq(data + 3 * lo);
    for (size_t i = lo + 1; i < hi; i++) acc = fp_mul(acc, pair_load_fq(data + 3 * i));
    fq_to_u4x3(acc, partial + 3 * t);
}
B200_HD void fq_inv_down_thread(uint4* data, const uint4* partial_inv, size_t n, size_t t) {
    const size_t lo = t * MSM_INV_CHUNK;
    if (lo >= n) return;
    const size_t hi = lo + MSM_INV_CHUNK < n ? lo + MSM_INV_CHUNK : n;
    fq_t pre[MSM_INV_CHUNK];
    fq_t acc = fp_one<FqP>();
    for (size_t i = lo; i < hi; i++) {
        pre[i - lo] = acc;
        acc = fp_mul(acc, pair_load_fq(data + 3 * i));
    }
    fq_t inv = pair_load_fq(partial_inv + 3 * t);
    for (size_t i = hi; i-- > lo;) {
        const fq_t x = pair_load_fq(data + 3 * i);
        fq_to_u4x3(fp_mul(inv, pre[i - lo]), data + 3 * i);
        inv = fp_mul(inv, x);
    }
}
// tail of the recursion: n values, `nthreads` threads, each inverts the product of its strided share by Fermat
#define MSM_INV_SMALL_PER_THREAD 8u
B200_HD void fq_inv_small_thread(uint4* data, size_t n, uint32_t t, uint32_t nthreads) {
    fq_t pre[MSM_INV_SMALL_PER_THREAD];
    fq_t acc = fp_one<FqP>();
    uint32_t cnt = 0;
    for (size_t i = t; i < n; i += nthreads, cnt++) {
        pre[cnt] = acc;
        acc = fp_mul(acc, pair_load_fq(data + 3 * i));
    }
    if (cnt == 0) return;
    fq_t inv = fp_inv(acc);
    for (uint32_t c = cnt; c-- > 0;) {
        const size_t i = t + (size_t)c * nthreads;
        const fq_t x = pair_load_fq(data + 3 * i);
        fq_to_u4x3(fp_mul(inv, pre[c]), data + 3 * i);
        inv = fp_mul(inv, x);
    }
}
