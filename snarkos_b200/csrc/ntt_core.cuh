// ntt_core.cuh -- radix-2 NTT over BLS12-377 Fr as a sequence of shared-memory passes.
//
// GPU counterpart of snarkVM EvaluationDomain::{fft,ifft,coset_fft,coset_ifft}_in_place
// [UPSTREAM snarkvm-algorithms 1.0.0 @ dea322b: algorithms/src/fft/domain.rs -- in_order_fft_in_place,
//  in_order_ifft_in_place, in_order_coset_(i)fft_in_place, distribute_powers; SURVEY.md 8a rows a3-a6,
//  appendix A.1].  Natural order in, natural order out, Montgomery data, result fully reduced.
//
// Decomposition (generalised four-step / Cooley-Tukey with k passes, N = L_0 * L_1 * ... * L_{k-1}):
//   pass i < k-1 works on sub-problems of size M_i = N / (L_0..L_{i-1}) (contiguous, base p * M_i):
//       for every column m < S_i = M_i / L_i:  L_i-point NTT over x[p*M_i + t*S_i + m], t -> k_i,
//       multiply by w_{M_i}^(m * k_i), store at p*M_i + k_i*S_i + m                       (in place layout)
//   last pass: contiguous L_{k-1}-point NTTs; element k_{k-1} of sub-problem p = (k_0,..,k_{k-2})
//       goes to  k_0 + L_0*k_1 + ... + (N / L_{k-1}) * k_{k-1}                            (digit reversal)
// so the output is in natural order without a separate bit-reversal pass over HBM.
//
// One CTA owns a tile of 2^log_len x 2^log_cw elements in shared memory ("cw" = adjacent columns for the
// strided passes, adjacent k_0 rows for the last pass) so that every global access is a run of
// 2^log_cw * 32 contiguous bytes.  Inside the tile: DIF butterflies (natural -> bit-reversed), the
// bit reversal is undone for free when the tile is read back out of shared memory.
//
// Every phase is a B200_HD function of (params, tile, batch, tid, nthreads): the CUDA kernel calls the
// phases with __syncthreads() in between, the host test shim (tests/host/ntt_host.cpp) calls them for all
// tids in turn -- same code, so the index arithmetic is verified on the CPU against the oracle.
#pragma once
#include "field.cuh"

#define NTT_MAX_PASSES 4
#if defined(__CUDACC__)
#define B200_UNROLL_SHAPED _Pragma("unroll (SH::unroll)")
#else
#define B200_UNROLL_SHAPED
#endif
#define NTT_TILE_TW_LOG 12          // tile twiddle table: w_{2^12}^e, e < 2^11
#define NTT_POW_LO_LOG 13           // two-level power tables: x^e = LO[e & 8191] * HI[e >> 13]

struct NttPassParams {
    const uint4* src;               // element = 2 consecutive uint4 (32 B)
    uint4* dst;
    const uint4* tile_tw;           // w_{2^12}^e (direction-specific), e < 2^11
    const uint4* pow_lo;            // w_N^e, e < 2^13                        (inter-pass twiddles)
    const uint4* pow_hi;            // w_N^(e << 13)
    const uint4* boundary_tw;       // optional: w_M^(m*k) for this pass at index k*S + m (M entries); null -> two-level
    const uint4* coset_lo;          // g^e (forward) or g^-e (inverse), e < 2^13
    const uint4* coset_hi;          // g^(e << 13) / g^-(e << 13)
    fr_t size_inv;                  // n^-1 (Montgomery)
    unsigned long long batch_stride;   // elements between consecutive polynomials
    uint32_t log_n;
    uint32_t pass, npasses;
    uint32_t log_len[NTT_MAX_PASSES];  // l_i of every pass (needed for the final digit reversal)
    uint32_t log_cw;                // tile width of THIS pass
    uint32_t coset_pre;             // multiply input j by g^j     (forward coset, first pass)
    uint32_t scale_post;            // multiply output by n^-1     (inverse, last pass)
    uint32_t coset_post;            // multiply output j by g^-j   (inverse coset, last pass)
    uint32_t radix4;                // 1: process stage pairs on register quads (default), 0: radix-2 stages only
    // Fused exchange of the multi-GPU four-step transform (xchg != 0, LAST pass of a batch of row transforms only): the
    // pass stores its outputs straight into the slabs of the ranks that own them -- transposed, optionally times the
    // twiddle between the two transform axes -- instead of writing the local slab for a separate exchange kernel.  The
    // tile then holds ONE sub-problem of 2^log_cw ADJACENT ROWS (polynomials), so that for every output column the rows
    // of the tile form one contiguous run of 2^log_cw * 32 B in the destination slab.
    uint32_t xchg, x_world, x_rank, x_twiddle, x_log_n;     // x_log_n: log2 of the distributed transform (twiddle domain)
    unsigned long long x_r_total, x_row_base;               // rows per rank, global index of this rank's row 0
    const uint4* x_pow_lo;          // w_N^e for the distributed size N = 2^x_log_n (two-level tables)
    const uint4* x_pow_hi;
    uint4* x_dst[16];               // destination slab of every rank (peer memory for the others)
};
#define NTT_XCHG_MAX_WORLD 16

// ---------------------------------------------------------------------------------------------
// element access helpers (global / shared element = two uint4 halves)
// ---------------------------------------------------------------------------------------------
B200_HD fr_t fr_from_u4(const uint4& lo, const uint4& hi) {
    fr_t r;
    r.v[0] = lo.x; r.v[1] = lo.y; r.v[2] = lo.z; r.v[3] = lo.w;
    r.v[4] = hi.x; r.v[5] = hi.y; r.v[6] = hi.z; r.v[7] = hi.w;
    return r;
}
B200_HD void fr_to_u4(const fr_t& a, uint4& lo, uint4& hi) {
    lo.x = a.v[0]; lo.y = a.v[1]; lo.z = a.v[2]; lo.w = a.v[3];
    hi.x = a.v[4]; hi.y = a.v[5]; hi.z = a.v[6]; hi.w = a.v[7];
}
B200_HD fr_t fr_load(const uint4* base, unsigned long long idx) {
    return fr_from_u4(base[2 * idx], base[2 * idx + 1]);
}
// shared tile: two planes of uint4 so that consecutive elements are consecutive 16-byte words
B200_HD fr_t tile_load(const uint4* sm, uint32_t tile_elems, uint32_t e) {
    return fr_from_u4(sm[e], sm[tile_elems + e]);
}
B200_HD void tile_store(uint4* sm, uint32_t tile_elems, uint32_t e, const fr_t& a) {
    fr_to_u4(a, sm[e], sm[tile_elems + e]);
}
B200_HD uint32_t bitrev32(uint32_t x, uint32_t bits) {
#if defined(__CUDACC__)
    return bits ? (__brev(x) >> (32 - bits)) : 0u;
#else
    uint32_t r = 0;
    for (uint32_t i = 0; i < bits; i++) r |= ((x >> i) & 1u) << (bits - 1 - i);
    return r;
#endif
}
// x^e from the two-level tables (e < 2^26)
B200_HD fr_t pow2level(const uint4* lo, const uint4* hi, unsigned long long e) {
    fr_t a = fr_load(lo, e & ((1u << NTT_POW_LO_LOG) - 1));
    fr_t b = fr_load(hi, e >> NTT_POW_LO_LOG);
    return fp_mul(a, b);
}

// ---------------------------------------------------------------------------------------------
// geometry of one pass
// ---------------------------------------------------------------------------------------------
struct NttGeom {
    uint32_t log_len, log_cw, log_sub, log_stride;   // this pass: L, CW, M_i, S_i
    uint32_t last;                                   // 1 for the final (transposing) pass
    uint32_t tile_elems;                             // L * CW
    uint32_t log_tiles;                              // tiles per polynomial = N / (L * CW)
};
// Compile-time shape of a pass (length, tile width, threads per CTA): with it the stage loops have constant trip
// counts and every shift / mask of the tile arithmetic folds into the instruction -- the specialised kernels of
// ntt.cu; NttDyn (all -1) is the generic run-time form, also what the host shim exercises by default.
// STORE: what the store phase of the pass has to do -- 0: anything (run-time flags); 1: strided pass with its boundary
// twiddle table (exactly one product per element); 2: last pass of a plain forward transform (no product at all).
// The generic store inlines six product sites per element (table / two-level twiddle, n^-1, coset powers) five times:
// 60 % of the kernel's code for paths a forward transform never takes.
template <int L, int CW, int NT, int UNR = 2, int STORE = 0> struct NttShape {
    static constexpr int log_len = L, log_cw = CW, nthreads = NT, unroll = UNR, store = STORE;
};
typedef NttShape<-1, -1, -1, 1, 0> NttDyn;
template <class SH> B200_HD uint32_t ntt_nthreads(uint32_t nthreads) { return SH::nthreads > 0 ? (uint32_t)SH::nthreads : nthreads; }

template <class SH> B200_HD NttGeom ntt_geom_t(const NttPassParams& p) {
    NttGeom g;
    g.log_len = SH::log_len >= 0 ? (uint32_t)SH::log_len : p.log_len[p.pass];
    g.log_cw = SH::log_cw >= 0 ? (uint32_t)SH::log_cw : p.log_cw;
    uint32_t before = 0;
    for (uint32_t i = 0; i < p.pass; i++) before += p.log_len[i];
    g.log_sub = p.log_n - before;
    g.log_stride = g.log_sub - g.log_len;
    g.last = (p.pass + 1 == p.npasses);
    g.tile_elems = 1u << (g.log_len + g.log_cw);
    g.log_tiles = p.log_n - g.log_len - g.log_cw;
    return g;
}
B200_HD NttGeom ntt_geom(const NttPassParams& p) { return ntt_geom_t<NttDyn>(p); }

// Global element index (within the polynomial) of tile element (t, cw) on the INPUT side.
//   strided pass: tile -> (sub-problem p, column base m0): idx = p*M + t*S + m0 + cw
//   last pass   : tile -> (k0 base, rest): row p = (k0_base + cw) * (P / L_0) + rest, idx = p * L + t
B200_HD unsigned long long ntt_in_index(const NttPassParams& p, const NttGeom& g, uint32_t tile, uint32_t t,
                                        uint32_t cw) {
    if (!g.last) {
        uint32_t log_tiles_per_sub = g.log_stride - g.log_cw;
        unsigned long long sub = tile >> log_tiles_per_sub;
        unsigned long long m0 = (unsigned long long)(tile & ((1u << log_tiles_per_sub) - 1)) << g.log_cw;
        return (sub << g.log_sub) + ((unsigned long long)t << g.log_stride) + m0 + cw;
    }
    if (p.npasses == 1) return t;                       // single pass: the whole polynomial, CW = 1
    uint32_t log_rows = p.log_n - g.log_len;            // P = number of sub-problems
    uint32_t l0 = p.log_len[0];
    uint32_t log_k0_tiles = l0 - g.log_cw;
    unsigned long long k0 = ((unsigned long long)(tile & ((1u << log_k0_tiles) - 1)) << g.log_cw) + cw;
    unsigned long long rest = tile >> log_k0_tiles;
    unsigned long long row = (k0 << (log_rows - l0)) + rest;
    return (row << g.log_len) + t;
}

// Fused-exchange last pass: element offset (from p.src, in elements) of tile element (t, cw): row (batch << log_cw) + cw of
// the batch, sub-problem `tile` of that row (its 2^log_len contiguous elements).
B200_HD unsigned long long ntt_xchg_in_offset(const NttPassParams& p, const NttGeom& g, uint32_t tile, uint32_t batch, uint32_t t,
                                              uint32_t cw) {
    const unsigned long long row = ((unsigned long long)batch << g.log_cw) + cw;
    return row * p.batch_stride + ((unsigned long long)tile << g.log_len) + t;
}

// ---------------------------------------------------------------------------------------------
// phase 1: global -> shared (coalesced), optional coset pre-scaling
// ---------------------------------------------------------------------------------------------
template <class SH = NttDyn>
B200_HD void ntt_phase_load(const NttPassParams& p, uint4* sm, uint32_t tile, uint32_t batch, uint32_t tid,
                            uint32_t nthreads_rt) {
    const NttGeom g = ntt_geom_t<SH>(p);
    const uint32_t nthreads = ntt_nthreads<SH>(nthreads_rt);
    const uint4* src = (g.last && p.xchg) ? p.src : p.src + 2ull * batch * p.batch_stride;       // xchg: the offset carries the row
    uint32_t total = 2u * g.tile_elems;                 // 16-byte words
    // Batches of NTT_LOAD_BATCH independent global loads before their shared-memory stores: the trip count is a run-time
    // value, so the compiler keeps ONE load in flight per thread otherwise and the phase pays the global latency once per
    // word (ncu r02: STS.128 behind the LDG carried 8 % of all stall samples, all of them long scoreboard).
#ifndef NTT_LOAD_BATCH
#define NTT_LOAD_BATCH 8
#endif
    uint32_t u = tid;
    for (; u + (NTT_LOAD_BATCH - 1) * nthreads < total; u += NTT_LOAD_BATCH * nthreads) {
        uint4 v[NTT_LOAD_BATCH];
        uint32_t dst[NTT_LOAD_BATCH];
        B200_UNROLL
        for (int k = 0; k < NTT_LOAD_BATCH; k++) {
            const uint32_t uu = u + k * nthreads;
            uint32_t half = uu & 1u, e = uu >> 1;
            uint32_t t, cw;
            if (!g.last) { cw = e & ((1u << g.log_cw) - 1); t = e >> g.log_cw; }      // cw fastest in memory
            else         { t = e & ((1u << g.log_len) - 1); cw = e >> g.log_len; }    // t fastest in memory
            unsigned long long idx = (g.last && p.xchg) ? ntt_xchg_in_offset(p, g, tile, batch, t, cw) : ntt_in_index(p, g, tile, t, cw);
            v[k] = src[2 * idx + half];
            dst[k] = half * g.tile_elems + ((t << g.log_cw) + cw);
        }
        B200_UNROLL
        for (int k = 0; k < NTT_LOAD_BATCH; k++) sm[dst[k]] = v[k];
    }
    for (; u < total; u += nthreads) {
        uint32_t half = u & 1u, e = u >> 1;
        uint32_t t, cw;
        if (!g.last) { cw = e & ((1u << g.log_cw) - 1); t = e >> g.log_cw; }
        else         { t = e & ((1u << g.log_len) - 1); cw = e >> g.log_len; }
        unsigned long long idx = (g.last && p.xchg) ? ntt_xchg_in_offset(p, g, tile, batch, t, cw) : ntt_in_index(p, g, tile, t, cw);
        sm[half * g.tile_elems + ((t << g.log_cw) + cw)] = src[2 * idx + half];
    }
}
// phase 1b (only when coset_pre): x[j] *= g^j     -- distribute_powers(coeffs, g) of coset_fft_in_place
B200_HD void ntt_phase_coset_pre(const NttPassParams& p, uint4* sm, uint32_t tile, uint32_t tid,
                                 uint32_t nthreads) {
    NttGeom g = ntt_geom(p);
    for (uint32_t e = tid; e < g.tile_elems; e += nthreads) {
        uint32_t cw = e & ((1u << g.log_cw) - 1), t = e >> g.log_cw;
        unsigned long long idx = ntt_in_index(p, g, tile, t, cw);
        fr_t x = tile_load(sm, g.tile_elems, e);
        x = fp_mul(x, pow2level(p.coset_lo, p.coset_hi, idx));
        tile_store(sm, g.tile_elems, e, x);
    }
}

// ---------------------------------------------------------------------------------------------
// phase 2 (x log_len): one DIF stage on the tile.  Stage s pairs t and t + d, d = L >> (s + 1):
//     a' = a + b,  b' = (a - b) * w_L^((t mod d) << s)
// ---------------------------------------------------------------------------------------------
// Twiddle source of the in-tile stages: w_L^e at lo[e << shift], hi[e << shift] (two planes of 16-byte words).
// The kernel stages the L/2 twiddles of its pass in shared memory (shift 0); the host shim reads the global
// w_{2^12} table directly (lo = table, hi = table + 1, stride folded into the shift by the accessor below).
struct NttTwiddles {
    const uint4* lo;
    const uint4* hi;
    uint32_t shift;      // index = e << shift
    uint32_t stride;     // distance between consecutive table entries in uint4 units (1: planar, 2: interleaved)
};
B200_HD fr_t ntt_twiddle(const NttTwiddles& t, uint32_t e) {
    const uint32_t i = (e << t.shift) * t.stride;
    return fr_from_u4(t.lo[i], t.hi[i]);
}
B200_HD NttTwiddles ntt_global_twiddles(const NttPassParams& p, uint32_t log_len) {
    NttTwiddles t;
    t.lo = p.tile_tw;
    t.hi = p.tile_tw + 1;
    t.shift = NTT_TILE_TW_LOG - log_len;
    t.stride = 2;
    return t;
}
// phase 0 (device): copy the L/2 twiddles of this pass into shared memory, planar
template <class SH = NttDyn>
B200_HD void ntt_phase_stage_twiddles(const NttPassParams& p, uint4* sm_tw, uint32_t tid, uint32_t nthreads_rt) {
    const uint32_t nthreads = ntt_nthreads<SH>(nthreads_rt);
    const uint32_t log_len = SH::log_len >= 0 ? (uint32_t)SH::log_len : p.log_len[p.pass];
    const uint32_t half = log_len ? (1u << (log_len - 1)) : 0;
    const uint32_t shift = NTT_TILE_TW_LOG - log_len;
    for (uint32_t u = tid; u < 2 * half; u += nthreads) {
        const uint32_t e = u >> 1, h = u & 1;
        sm_tw[h * half + e] = p.tile_tw[2 * ((unsigned long long)e << shift) + h];
    }
}
B200_HD NttTwiddles ntt_shared_twiddles(const uint4* sm_tw, uint32_t log_len) {
    NttTwiddles t;
    t.lo = sm_tw;
    t.hi = sm_tw + (log_len ? (1u << (log_len - 1)) : 0);
    t.shift = 0;
    t.stride = 1;
    return t;
}

template <class SH = NttDyn>
B200_HD void ntt_phase_stage(const NttPassParams& p, uint4* sm, const NttTwiddles& twd, uint32_t s, uint32_t tid,
                             uint32_t nthreads_rt) {
    const NttGeom g = ntt_geom_t<SH>(p);
    const uint32_t nthreads = ntt_nthreads<SH>(nthreads_rt);
    uint32_t log_d = g.log_len - 1 - s;
    uint32_t nbf = g.tile_elems >> 1;                   // butterflies in the tile
    // Butterfly (block b, offset j) pairs t0 = b*2d + j with t0 + d and uses twiddle w_L^(j << s); j == 0 needs no
    // product.  For the long-distance stages consecutive lanes take consecutive j (conflict-free shared memory);
    // for the last stages (d <= 4) lanes are ordered j-major instead, so that whole warps have j == 0 and really
    // skip the multiplication (1/d of the butterflies of a stage; d = 1: all of them).
    const bool j_major = log_d <= 2;
    const uint32_t log_blocks = g.log_len - 1 - log_d;  // blocks of 2d points
    B200_UNROLL_SHAPED
    for (uint32_t u = tid; u < nbf; u += nthreads) {
        uint32_t cw = u & ((1u << g.log_cw) - 1), q = u >> g.log_cw;
        uint32_t j, b;
        if (j_major) { b = q & ((1u << log_blocks) - 1); j = q >> log_blocks; }
        else         { j = q & ((1u << log_d) - 1); b = q >> log_d; }
        uint32_t t0 = (b << (log_d + 1)) | j;
        uint32_t e0 = (t0 << g.log_cw) + cw, e1 = e0 + (1u << (log_d + g.log_cw));
        fr_t a = tile_load(sm, g.tile_elems, e0);
        fr_t bb = tile_load(sm, g.tile_elems, e1);
        fr_t sum = fp_add(a, bb);
        fr_t dif = fp_sub(a, bb);
        uint32_t tw = j << s;                                        // exponent of w_L
        if (tw) dif = fp_mul(dif, ntt_twiddle(twd, tw));
        tile_store(sm, g.tile_elems, e0, sum);
        tile_store(sm, g.tile_elems, e1, dif);
    }
}

// Two consecutive DIF stages (s, s + 1) on register-resident quads: half the shared-memory traffic, half the
// barriers and half the index arithmetic of two radix-2 stages; same number of products.  With d1 = L >> (s + 1),
// d2 = d1 / 2, a quad is {p, p + d2, p + d1, p + d1 + d2} with p = b * 2 * d1 + j (j < d2):
//   stage s    : (x0, x2) twiddle w_L^(j << s)          (x1, x3) twiddle w_L^((j + d2) << s)
//   stage s + 1: (x0', x1') and (x2', x3'), both twiddle w_L^(j << (s + 1))
// For the short-distance quads (d2 <= 4) lanes are ordered j-major so that the j == 0 quads, whose two twiddles are 1,
// fill whole warps and really skip their products.
template <class SH = NttDyn>
B200_HD void ntt_phase_stage2(const NttPassParams& p, uint4* sm, const NttTwiddles& twd, uint32_t s, uint32_t tid,
                              uint32_t nthreads_rt) {
    const NttGeom g = ntt_geom_t<SH>(p);
    const uint32_t nthreads = ntt_nthreads<SH>(nthreads_rt);
    const uint32_t log_d1 = g.log_len - 1 - s, log_d2 = log_d1 - 1;
    const uint32_t nquads = g.tile_elems >> 2;
    const bool j_major = log_d2 <= 2;
    const uint32_t log_blocks = g.log_len - 1 - log_d1;      // blocks of 2 * d1 points
    B200_UNROLL_SHAPED
    for (uint32_t u = tid; u < nquads; u += nthreads) {
        const uint32_t cw = u & ((1u << g.log_cw) - 1), q = u >> g.log_cw;
        uint32_t j, b;
        if (j_major) { b = q & ((1u << log_blocks) - 1); j = q >> log_blocks; }
        else         { j = q & ((1u << log_d2) - 1); b = q >> log_d2; }
        const uint32_t t0 = (b << (log_d1 + 1)) | j;
        const uint32_t e0 = (t0 << g.log_cw) + cw;
        const uint32_t o2 = 1u << (log_d2 + g.log_cw), o1 = o2 << 1;
        fr_t x0 = tile_load(sm, g.tile_elems, e0);
        fr_t x1 = tile_load(sm, g.tile_elems, e0 + o2);
        fr_t x2 = tile_load(sm, g.tile_elems, e0 + o1);
        fr_t x3 = tile_load(sm, g.tile_elems, e0 + o1 + o2);
        // stage s
        fr_t a0 = fp_add(x0, x2), a2 = fp_sub(x0, x2);
        fr_t a1 = fp_add(x1, x3), a3 = fp_sub(x1, x3);
        const uint32_t twa = j << s;
        const uint32_t twb = (j + (1u << log_d2)) << s;                    // never 0
        if (twa) a2 = fp_mul(a2, ntt_twiddle(twd, twa));
        a3 = fp_mul(a3, ntt_twiddle(twd, twb));
        // stage s + 1
        const uint32_t tw2 = j << (s + 1);
        fr_t y0 = fp_add(a0, a1), y1 = fp_sub(a0, a1);
        fr_t y2 = fp_add(a2, a3), y3 = fp_sub(a2, a3);
        if (tw2) {
            fr_t w2 = ntt_twiddle(twd, tw2);
            y1 = fp_mul(y1, w2);
            y3 = fp_mul(y3, w2);
        }
        tile_store(sm, g.tile_elems, e0, y0);
        tile_store(sm, g.tile_elems, e0 + o2, y1);
        tile_store(sm, g.tile_elems, e0 + o1, y2);
        tile_store(sm, g.tile_elems, e0 + o1 + o2, y3);
    }
}

// ---------------------------------------------------------------------------------------------
// phase 3: shared -> global with the bit reversal undone, inter-pass twiddle / final scaling
// ---------------------------------------------------------------------------------------------
// one element of the store phase; tw = its boundary twiddle when the caller has already fetched it (have_tw)
// tile position (t, cw) that holds output (k, cw) of the pass: DIF leaves output k at the bit-reversed row
B200_HD void ntt_store_source(const NttGeom& g, uint32_t e, uint32_t& t, uint32_t& cw) {
    cw = e & ((1u << g.log_cw) - 1);
    t = bitrev32(e >> g.log_cw, g.log_len);
}
template <int STORE = 0>
B200_HD void ntt_store_element(const NttPassParams& p, const NttGeom& g, fr_t x, uint4* dst, uint32_t tile, uint32_t e,
                               bool have_tw, const fr_t& tw, uint32_t batch = 0) {
    uint32_t cw = e & ((1u << g.log_cw) - 1), k = e >> g.log_cw;          // k = output index of this pass
    unsigned long long out;
    if (STORE == 0 && g.last && p.xchg) {
        // column of this output inside its row: digit reversal of the sub-problem index `tile` = (k0, k1, ..) and k
        const uint32_t log_rows = p.log_n - g.log_len;
        unsigned long long col;
        if (p.npasses == 1) {
            col = k;
        } else {
            const uint32_t l0 = p.log_len[0];
            unsigned long long acc = (unsigned long long)tile >> (log_rows - l0);      // k0
            const unsigned long long rest = tile & ((1ull << (log_rows - l0)) - 1);
            uint32_t shift = l0, rem_bits = log_rows - l0;
            for (uint32_t i = 1; i + 1 < p.npasses; i++) {
                rem_bits -= p.log_len[i];
                acc += ((rest >> rem_bits) & ((1ull << p.log_len[i]) - 1)) << shift;
                shift += p.log_len[i];
            }
            col = acc + ((unsigned long long)k << log_rows);
        }
        const unsigned long long row = ((unsigned long long)batch << g.log_cw) + cw;     // local row of the slab
        if (p.scale_post) x = fp_mul(x, p.size_inv);
        if (p.x_twiddle) {
            const unsigned long long ex = ((p.x_row_base + row) * col) & ((1ull << p.x_log_n) - 1);
            if (ex) x = fp_mul(x, pow2level(p.x_pow_lo, p.x_pow_hi, ex));
        }
        // destination: the rank that owns column `col` of the [r_total * world x 2^log_n] matrix, transposed slab
        const unsigned long long c_local = (1ull << p.log_n) / p.x_world;
        const unsigned long long d = col / c_local, cl = col - d * c_local;
        uint4* o = p.x_dst[d] + 2 * (cl * (p.x_r_total * p.x_world) + (unsigned long long)p.x_rank * p.x_r_total + row);
        uint4 lo, hi;
        fr_to_u4(x, lo, hi);
        o[0] = lo;
        o[1] = hi;
        return;
    }
    if (STORE == 1) {                                                     // strided pass, twiddle already fetched
        uint32_t log_tiles_per_sub = g.log_stride - g.log_cw;
        unsigned long long sub = tile >> log_tiles_per_sub;
        unsigned long long m = ((unsigned long long)(tile & ((1u << log_tiles_per_sub) - 1)) << g.log_cw) + cw;
        out = (sub << g.log_sub) + ((unsigned long long)k << g.log_stride) + m;
        if (m && k) x = fp_mul(x, tw);
    } else if (STORE != 2 && !g.last) {
        uint32_t log_tiles_per_sub = g.log_stride - g.log_cw;
        unsigned long long sub = tile >> log_tiles_per_sub;
        unsigned long long m = ((unsigned long long)(tile & ((1u << log_tiles_per_sub) - 1)) << g.log_cw) + cw;
        out = (sub << g.log_sub) + ((unsigned long long)k << g.log_stride) + m;
        // twiddle w_{M}^(m*k): one product from the per-pass table when it exists (same offset as the output
        // inside its sub-problem), else w_N^((m*k mod M) << (log_n - log_sub)) from the two-level tables
        if (p.boundary_tw) {
            if (m && k) x = fp_mul(x, have_tw ? tw : fr_load(p.boundary_tw, ((unsigned long long)k << g.log_stride) + m));
        } else {
            unsigned long long ex = ((m * k) & ((1ull << g.log_sub) - 1)) << (p.log_n - g.log_sub);
            if (ex) x = fp_mul(x, pow2level(p.pow_lo, p.pow_hi, ex));
        }
    } else if (p.npasses == 1) {
        out = k;
    } else {
        // digit reversal of the row index: row = (k0, k1, .., k_{last-1}) -> k0 + L0*k1 + ...
        uint32_t log_rows = p.log_n - g.log_len;
        uint32_t l0 = p.log_len[0];
        uint32_t log_k0_tiles = l0 - g.log_cw;
        unsigned long long k0 = ((unsigned long long)(tile & ((1u << log_k0_tiles) - 1)) << g.log_cw) + cw;
        unsigned long long rest = tile >> log_k0_tiles;            // digits k1 .. k_{last-1}, k1 most significant
        unsigned long long acc = k0;
        uint32_t shift = l0;
        uint32_t rem_bits = log_rows - l0;
        for (uint32_t i = 1; i + 1 < p.npasses; i++) {
            rem_bits -= p.log_len[i];
            unsigned long long digit = (rest >> rem_bits) & ((1ull << p.log_len[i]) - 1);
            acc += digit << shift;
            shift += p.log_len[i];
        }
        out = acc + ((unsigned long long)k << log_rows);
    }
    if (STORE == 0 && g.last) {
        if (p.scale_post) x = fp_mul(x, p.size_inv);
        if (p.coset_post) x = fp_mul(x, pow2level(p.coset_lo, p.coset_hi, out));
    }
    uint4 lo, hi;
    fr_to_u4(x, lo, hi);
    dst[2 * out] = lo;                             // (one 256-bit store per element: no measurable difference, L2 merges the halves)
    dst[2 * out + 1] = hi;
}

// how a tile is laid out in shared memory: planar (default kernel) or warp-column (below)
struct PlanarTile {
    const uint4* sm;
    uint32_t tile_elems, log_cw;
    B200_HDM fr_t get(uint32_t t, uint32_t cw) const { return tile_load(sm, tile_elems, (t << log_cw) + cw); }
};
template <class Tile, class SH = NttDyn>
B200_HD void ntt_phase_store_t(const NttPassParams& p, const Tile& T, uint32_t tile, uint32_t batch, uint32_t tid,
                               uint32_t nthreads_rt) {
    const NttGeom g = ntt_geom_t<SH>(p);
    const uint32_t nthreads = ntt_nthreads<SH>(nthreads_rt);
    uint4* dst = p.dst + 2ull * batch * p.batch_stride;
    uint32_t e = tid;
    if (SH::store == 2) {                                     // last pass, nothing to multiply: digit-reversed store only
        const fr_t none2 = fp_zero<FrP>();
        B200_UNROLL_SHAPED
        for (; e < g.tile_elems; e += nthreads) {
            uint32_t t, cw;
            ntt_store_source(g, e, t, cw);
            ntt_store_element<2>(p, g, T.get(t, cw), dst, tile, e, false, none2);
        }
        return;
    }
    if (SH::store == 1 || (!g.last && p.boundary_tw)) {
        // the boundary twiddles of NTT_STORE_BATCH elements are fetched together, ahead of the products that use them
        // (ncu r02: the first IMAD behind the one-at-a-time twiddle load carried 7 % of all stall samples)
#ifndef NTT_STORE_BATCH
#define NTT_STORE_BATCH 4
#endif
        const uint32_t log_tiles_per_sub = g.log_stride - g.log_cw;
        const unsigned long long m0 = (unsigned long long)(tile & ((1u << log_tiles_per_sub) - 1)) << g.log_cw;
        for (; e + (NTT_STORE_BATCH - 1) * nthreads < g.tile_elems; e += NTT_STORE_BATCH * nthreads) {
            fr_t tw[NTT_STORE_BATCH];
            B200_UNROLL
            for (int q = 0; q < NTT_STORE_BATCH; q++) {
                const uint32_t ee = e + q * nthreads;
                const uint32_t cw = ee & ((1u << g.log_cw) - 1), k = ee >> g.log_cw;
                tw[q] = fr_load(p.boundary_tw, ((unsigned long long)k << g.log_stride) + m0 + cw);
            }
            B200_UNROLL
            for (int q = 0; q < NTT_STORE_BATCH; q++) {
                uint32_t t, cw;
                ntt_store_source(g, e + q * nthreads, t, cw);
                ntt_store_element<SH::store>(p, g, T.get(t, cw), dst, tile, e + q * nthreads, true, tw[q], batch);
            }
        }
    }
    if (SH::store == 1) return;                               // compile-time shape: the batches cover the tile exactly
    const fr_t none = fp_zero<FrP>();
    for (; e < g.tile_elems; e += nthreads) {
        uint32_t t, cw;
        ntt_store_source(g, e, t, cw);
        ntt_store_element(p, g, T.get(t, cw), dst, tile, e, false, none, batch);
    }
}
template <class SH = NttDyn>
B200_HD void ntt_phase_store(const NttPassParams& p, const uint4* sm, uint32_t tile, uint32_t batch, uint32_t tid,
                             uint32_t nthreads) {
    const NttGeom g = ntt_geom_t<SH>(p);
    PlanarTile T;
    T.sm = sm;
    T.tile_elems = g.tile_elems;
    T.log_cw = g.log_cw;
    ntt_phase_store_t<PlanarTile, SH>(p, T, tile, batch, tid, nthreads);
}

// =============================================================================================
// Warp-column variant of a pass of length 2^8 over 4 adjacent columns (ntt_pass_wc_kernel in ntt.cu).
// The default kernel above runs every stage pair as a block-wide phase over the whole tile: four round trips through
// shared memory and six block barriers per tile, and all warps of a CTA sit in the same phase at the same time (ncu r02:
// the multiplier pipe is busy 74 % of the cycles, 1226 instructions per element and pass of which 480 are IMAD.WIDE).
// Here ONE WARP OWNS ONE COLUMN (256 elements, 8 per lane = 64 registers of data) and runs its 8 DIF stages as three
// in-register rounds -- radix 8 (distances 128, 64, 32), radix 8 (16, 8, 4), radix 4 (2, 1) -- exchanging through its own
// slice of the tile with nothing but __syncwarp in between; the CTA (4 warps, one 32 KB tile) meets at two block
// barriers only (tile loaded / tile transformed), so warps and CTAs drift apart and load, exchange and multiply
// phases overlap.  Index arithmetic is compile-time; every lane has 4 independent butterflies in flight per stage.
//
// Shared layout: two planes of 16-byte halves; column cw of a plane starts at cw * 257 words and element t sits at
//     wc_idx(t) = t ^ ((t >> 3) & 3) ^ ((((t >> 5) ^ (t >> 7)) & 1) << 2)
// -- a swizzle of the low three word bits by higher bits of t that makes all access patterns of the kernel hit 8
// distinct 16-byte bank groups per quarter-warp: round 1 (t = lane + 32 i), round 2 (t = 32 a + 4 k + b, lane = 4 a + b),
// round 3 (t = 4 (lane + 32 j) + r), the cooperative fill (4 columns x 2 halves of one row; the plane stride
// 4 * 257 = 4 mod 8 separates the halves, the column stride 257 = 1 mod 8 the columns) and the read-out in
// bit-reversed row order (rows k, k + 1 come from t, t + 128).
// =============================================================================================
#define WC_LOG_LEN 8u
#define WC_LOG_CW 2u
#define WC_COL_STRIDE 257u
#define WC_PLANE_STRIDE (4u * WC_COL_STRIDE)
#define WC_TILE_U4 (2u * WC_PLANE_STRIDE)              // 16-byte words of one tile (32.9 KB)

B200_HD uint32_t wc_idx(uint32_t t) { return t ^ ((t >> 3) & 3u) ^ ((((t >> 5) ^ (t >> 7)) & 1u) << 2); }
B200_HD uint32_t wc_word(uint32_t half, uint32_t t, uint32_t cw) { return half * WC_PLANE_STRIDE + cw * WC_COL_STRIDE + wc_idx(t); }
B200_HD fr_t wc_load(const uint4* sm, uint32_t t, uint32_t cw) { return fr_from_u4(sm[wc_word(0, t, cw)], sm[wc_word(1, t, cw)]); }
B200_HD void wc_store(uint4* sm, uint32_t t, uint32_t cw, const fr_t& x) { fr_to_u4(x, sm[wc_word(0, t, cw)], sm[wc_word(1, t, cw)]); }
struct WcTile {
    const uint4* sm;
    B200_HDM fr_t get(uint32_t t, uint32_t cw) const { return wc_load(sm, t, cw); }
};
B200_HD bool ntt_wc_applicable(const NttPassParams& p) { return p.log_len[p.pass] == WC_LOG_LEN && p.log_cw == WC_LOG_CW; }

// global -> shared, the same coalesced runs as ntt_phase_load (whole CTA)
B200_HD void wc_phase_load(const NttPassParams& p, uint4* sm, uint32_t tile, uint32_t batch, uint32_t tid, uint32_t nthreads) {
    const NttGeom g = ntt_geom(p);
    const uint4* src = p.src + 2ull * batch * p.batch_stride;
    const uint32_t total = 2u * g.tile_elems;
    uint32_t u = tid;
    for (; u + (NTT_LOAD_BATCH - 1) * nthreads < total; u += NTT_LOAD_BATCH * nthreads) {
        uint4 v[NTT_LOAD_BATCH];
        uint32_t dst[NTT_LOAD_BATCH];
        B200_UNROLL
        for (int k = 0; k < NTT_LOAD_BATCH; k++) {
            const uint32_t uu = u + k * nthreads;
            const uint32_t half = uu & 1u, e = uu >> 1;
            uint32_t t, cw;
            if (!g.last) { cw = e & 3u; t = e >> 2; }
            else         { t = e & 255u; cw = e >> 8; }
            v[k] = src[2 * ntt_in_index(p, g, tile, t, cw) + half];
            dst[k] = wc_word(half, t, cw);
        }
        B200_UNROLL
        for (int k = 0; k < NTT_LOAD_BATCH; k++) sm[dst[k]] = v[k];
    }
    for (; u < total; u += nthreads) {
        const uint32_t half = u & 1u, e = u >> 1;
        uint32_t t, cw;
        if (!g.last) { cw = e & 3u; t = e >> 2; }
        else         { t = e & 255u; cw = e >> 8; }
        sm[wc_word(half, t, cw)] = src[2 * ntt_in_index(p, g, tile, t, cw) + half];
    }
}

// DIF butterfly: (a, b) <- (a + b, (a - b) * w)
B200_HD void wc_bf(fr_t& a, fr_t& b, const fr_t& w) {
    const fr_t s = fp_add(a, b);
    b = fp_mul(fp_sub(a, b), w);
    a = s;
}
B200_HD void wc_bf1(fr_t& a, fr_t& b) {                 // twiddle 1
    const fr_t s = fp_add(a, b);
    b = fp_sub(a, b);
    a = s;
}
// three DIF stages on 8 register-resident points x[0..7] at distances 4, 2, 1 (in units of the register index);
// e0[i] (i < 4), e1[i] (i < 2), e2: exponents of w_256 for the three stages
B200_HD void wc_radix8(fr_t* x, const NttTwiddles& twd, const uint32_t* e0, const uint32_t* e1, uint32_t e2) {
    B200_UNROLL
    for (int i = 0; i < 4; i++) wc_bf(x[i], x[i + 4], ntt_twiddle(twd, e0[i]));
    B200_UNROLL
    for (int i = 0; i < 2; i++) {
        const fr_t w = ntt_twiddle(twd, e1[i]);
        wc_bf(x[i], x[i + 2], w);
        wc_bf(x[i + 4], x[i + 6], w);
    }
    {
        const fr_t w = ntt_twiddle(twd, e2);
        B200_UNROLL
        for (int i = 0; i < 8; i += 2) wc_bf(x[i], x[i + 1], w);
    }
}

// round 1: stages 0..2 (distances 128, 64, 32); lane holds t = lane + 32 i.  Forward coset scaling folded in.
B200_HD void wc_round1(const NttPassParams& p, uint4* sm, const NttTwiddles& twd, uint32_t tile, uint32_t cw, uint32_t lane) {
    fr_t x[8];
    B200_UNROLL
    for (int i = 0; i < 8; i++) x[i] = wc_load(sm, lane + 32u * i, cw);
    if (p.coset_pre) {
        const NttGeom g = ntt_geom(p);
        B200_UNROLL
        for (int i = 0; i < 8; i++)
            x[i] = fp_mul(x[i], pow2level(p.coset_lo, p.coset_hi, ntt_in_index(p, g, tile, lane + 32u * i, cw)));
    }
    // stage s pairs t, t + d (d = 128 >> s) with twiddle w^((t mod d) << s)
    const uint32_t e0[4] = {lane, lane + 32u, lane + 64u, lane + 96u};
    const uint32_t e1[2] = {lane << 1, (lane + 32u) << 1};
    wc_radix8(x, twd, e0, e1, lane << 2);
    B200_UNROLL
    for (int i = 0; i < 8; i++) wc_store(sm, lane + 32u * i, cw, x[i]);
}
// round 2: stages 3..5 (distances 16, 8, 4); lane = 4 a + b holds t = 32 a + 4 k + b
B200_HD void wc_round2(uint4* sm, const NttTwiddles& twd, uint32_t cw, uint32_t lane) {
    const uint32_t a = lane >> 2, b = lane & 3u, t0 = 32u * a + b;
    fr_t x[8];
    B200_UNROLL
    for (int k = 0; k < 8; k++) x[k] = wc_load(sm, t0 + 4u * k, cw);
    const uint32_t e0[4] = {b << 3, (b + 4u) << 3, (b + 8u) << 3, (b + 12u) << 3};      // (t mod 16) << 3
    const uint32_t e1[2] = {b << 4, (b + 4u) << 4};                                       // (t mod 8) << 4
    wc_radix8(x, twd, e0, e1, b << 5);                                                    // (t mod 4) << 5
    B200_UNROLL
    for (int k = 0; k < 8; k++) wc_store(sm, t0 + 4u * k, cw, x[k]);
}
// round 3: stages 6, 7 (distances 2, 1); lane holds the quads t = 4 (lane + 32 j) + r
B200_HD void wc_round3(uint4* sm, const NttTwiddles& twd, uint32_t cw, uint32_t lane) {
    const fr_t w64 = ntt_twiddle(twd, 64u);                                               // (t mod 2) << 6 for odd t
    B200_UNROLL
    for (int j = 0; j < 2; j++) {
        const uint32_t t0 = 4u * (lane + 32u * j);
        fr_t x0 = wc_load(sm, t0, cw), x1 = wc_load(sm, t0 + 1, cw), x2 = wc_load(sm, t0 + 2, cw), x3 = wc_load(sm, t0 + 3, cw);
        wc_bf1(x0, x2);
        wc_bf(x1, x3, w64);
        wc_bf1(x0, x1);
        wc_bf1(x2, x3);
        wc_store(sm, t0, cw, x0);
        wc_store(sm, t0 + 1, cw, x1);
        wc_store(sm, t0 + 2, cw, x2);
        wc_store(sm, t0 + 3, cw, x3);
    }
}
B200_HD void wc_phase_store(const NttPassParams& p, const uint4* sm, uint32_t tile, uint32_t batch, uint32_t tid, uint32_t nthreads) {
    WcTile T;
    T.sm = sm;
    ntt_phase_store_t(p, T, tile, batch, tid, nthreads);
}

// =============================================================================================
// Bulk-copy (TMA) variant of a pass: the tile arrives in shared memory through cp.async.bulk (SASS UBLKCP) -- one
// asynchronous copy per contiguous global run, completion counted on an mbarrier -- so the load of tile i + 1 runs under
// the butterflies of tile i (ntt_pass_bulk_kernel in ntt.cu: persistent CTAs, two tile buffers).  A bulk copy lands
// bytes exactly as they lie in global memory, so the tile is INTERLEAVED here (element i = chunks 2i, 2i + 1 of 16 B)
// instead of the two planes of the default kernel:
//     strided pass: run t = the CW adjacent columns of row t          -> element (t, cw) at  t * CW + cw
//     last pass   : run cw = the whole contiguous sub-problem of k0 row cw (2^L elements) -> element (t, cw) at
//                   cw * (2^L + 1) + t   (one element of padding per run: lanes that differ in cw then fall on
//                   different banks exactly like consecutive elements do)
// Bank conflicts of the interleaved layout (a quarter-warp reading the low halves of 8 consecutive elements spans
// 256 B: 2-way) are avoided by letting lanes 4..7 of every quarter-warp touch the HIGH half first (h = (lane >> 2) & 1)
// and swapping in registers: each 128-bit access of a quarter-warp then covers 8 distinct 16-byte bank groups -- the
// software form of the 32-byte TMA swizzle.
// The arithmetic (stages, twiddles, boundary products, digit reversal) is the same as above, element for element.
// =============================================================================================
struct NttBulkTile {
    uint4* sm;
    uint32_t last, log_len, log_cw;
    uint32_t h;                    // 0 / 1: which half this lane loads / stores first (any value gives the same result)
};
B200_HOSTDEV uint32_t ntt_bulk_tile_elems(uint32_t log_len, uint32_t log_cw, uint32_t last) {
    return last ? (((1u << log_len) + 1u) << log_cw) : (1u << (log_len + log_cw));
}
B200_HD uint32_t bulk_idx(const NttBulkTile& T, uint32_t t, uint32_t cw) {
    return T.last ? cw * ((1u << T.log_len) + 1u) + t : (t << T.log_cw) + cw;
}
B200_HD fr_t bulk_load(const NttBulkTile& T, uint32_t i) {
    const uint4 a = T.sm[2 * i + T.h], b = T.sm[2 * i + (T.h ^ 1u)];
    return T.h ? fr_from_u4(b, a) : fr_from_u4(a, b);
}
B200_HD void bulk_store(const NttBulkTile& T, uint32_t i, const fr_t& x) {
    uint4 lo, hi;
    fr_to_u4(x, lo, hi);
    T.sm[2 * i + T.h] = T.h ? hi : lo;
    T.sm[2 * i + (T.h ^ 1u)] = T.h ? lo : hi;
}

// The contiguous global runs of one tile: run r (r < ntt_bulk_runs) starts at element `src_elem` of the polynomial, is
// `elems` long and lands at element `dst_elem` of the tile.
B200_HD uint32_t ntt_bulk_runs(const NttPassParams& p) {
    const NttGeom g = ntt_geom(p);
    return g.last ? (1u << g.log_cw) : (1u << g.log_len);
}
B200_HD void ntt_bulk_run(const NttPassParams& p, uint32_t tile, uint32_t r, unsigned long long& src_elem,
                          uint32_t& dst_elem, uint32_t& elems) {
    const NttGeom g = ntt_geom(p);
    if (!g.last) {
        src_elem = ntt_in_index(p, g, tile, r, 0);
        dst_elem = r << g.log_cw;
        elems = 1u << g.log_cw;
    } else {
        src_elem = ntt_in_index(p, g, tile, 0, r);
        dst_elem = r * ((1u << g.log_len) + 1u);
        elems = 1u << g.log_len;
    }
}

B200_HD void ntt_bulk_coset_pre(const NttPassParams& p, const NttBulkTile& T, uint32_t tile, uint32_t tid, uint32_t nthreads) {
    const NttGeom g = ntt_geom(p);
    for (uint32_t e = tid; e < g.tile_elems; e += nthreads) {
        const uint32_t cw = e & ((1u << g.log_cw) - 1), t = e >> g.log_cw;
        const unsigned long long idx = ntt_in_index(p, g, tile, t, cw);
        const uint32_t i = bulk_idx(T, t, cw);
        fr_t x = bulk_load(T, i);
        x = fp_mul(x, pow2level(p.coset_lo, p.coset_hi, idx));
        bulk_store(T, i, x);
    }
}

B200_HD void ntt_bulk_stage(const NttPassParams& p, const NttBulkTile& T, const NttTwiddles& twd, uint32_t s, uint32_t tid,
                            uint32_t nthreads) {
    const NttGeom g = ntt_geom(p);
    const uint32_t log_d = g.log_len - 1 - s;
    const uint32_t nbf = g.tile_elems >> 1;
    const bool j_major = log_d <= 2;
    const uint32_t log_blocks = g.log_len - 1 - log_d;
    for (uint32_t u = tid; u < nbf; u += nthreads) {
        const uint32_t cw = u & ((1u << g.log_cw) - 1), q = u >> g.log_cw;
        uint32_t j, b;
        if (j_major) { b = q & ((1u << log_blocks) - 1); j = q >> log_blocks; }
        else         { j = q & ((1u << log_d) - 1); b = q >> log_d; }
        const uint32_t t0 = (b << (log_d + 1)) | j;
        const uint32_t i0 = bulk_idx(T, t0, cw), i1 = bulk_idx(T, t0 + (1u << log_d), cw);
        const fr_t a = bulk_load(T, i0), bb = bulk_load(T, i1);
        const fr_t sum = fp_add(a, bb);
        fr_t dif = fp_sub(a, bb);
        const uint32_t tw = j << s;
        if (tw) dif = fp_mul(dif, ntt_twiddle(twd, tw));
        bulk_store(T, i0, sum);
        bulk_store(T, i1, dif);
    }
}

B200_HD void ntt_bulk_stage2(const NttPassParams& p, const NttBulkTile& T, const NttTwiddles& twd, uint32_t s, uint32_t tid,
                             uint32_t nthreads) {
    const NttGeom g = ntt_geom(p);
    const uint32_t log_d1 = g.log_len - 1 - s, log_d2 = log_d1 - 1;
    const uint32_t nquads = g.tile_elems >> 2;
    const bool j_major = log_d2 <= 2;
    const uint32_t log_blocks = g.log_len - 1 - log_d1;
    for (uint32_t u = tid; u < nquads; u += nthreads) {
        const uint32_t cw = u & ((1u << g.log_cw) - 1), q = u >> g.log_cw;
        uint32_t j, b;
        if (j_major) { b = q & ((1u << log_blocks) - 1); j = q >> log_blocks; }
        else         { j = q & ((1u << log_d2) - 1); b = q >> log_d2; }
        const uint32_t t0 = (b << (log_d1 + 1)) | j;
        const uint32_t d2 = 1u << log_d2, d1 = d2 << 1;
        const uint32_t i0 = bulk_idx(T, t0, cw), i1 = bulk_idx(T, t0 + d2, cw), i2 = bulk_idx(T, t0 + d1, cw),
                       i3 = bulk_idx(T, t0 + d1 + d2, cw);
        const fr_t x0 = bulk_load(T, i0), x1 = bulk_load(T, i1), x2 = bulk_load(T, i2), x3 = bulk_load(T, i3);
        fr_t a0 = fp_add(x0, x2), a2 = fp_sub(x0, x2);
        fr_t a1 = fp_add(x1, x3), a3 = fp_sub(x1, x3);
        const uint32_t twa = j << s;
        const uint32_t twb = (j + d2) << s;
        if (twa) a2 = fp_mul(a2, ntt_twiddle(twd, twa));
        a3 = fp_mul(a3, ntt_twiddle(twd, twb));
        const uint32_t tw2 = j << (s + 1);
        fr_t y0 = fp_add(a0, a1), y1 = fp_sub(a0, a1);
        fr_t y2 = fp_add(a2, a3), y3 = fp_sub(a2, a3);
        if (tw2) {
            const fr_t w2 = ntt_twiddle(twd, tw2);
            y1 = fp_mul(y1, w2);
            y3 = fp_mul(y3, w2);
        }
        bulk_store(T, i0, y0);
        bulk_store(T, i1, y1);
        bulk_store(T, i2, y2);
        bulk_store(T, i3, y3);
    }
}

// shared -> global: same addresses, twiddles and scalings as ntt_phase_store
B200_HD void ntt_bulk_store_out(const NttPassParams& p, const NttBulkTile& T, uint32_t tile, uint32_t batch, uint32_t tid,
                                uint32_t nthreads) {
    const NttGeom g = ntt_geom(p);
    uint4* dst = p.dst + 2ull * batch * p.batch_stride;
    uint32_t e = tid;
    if (!g.last && p.boundary_tw) {
        const uint32_t log_tiles_per_sub = g.log_stride - g.log_cw;
        const unsigned long long m0 = (unsigned long long)(tile & ((1u << log_tiles_per_sub) - 1)) << g.log_cw;
        for (; e + (NTT_STORE_BATCH - 1) * nthreads < g.tile_elems; e += NTT_STORE_BATCH * nthreads) {
            fr_t tw[NTT_STORE_BATCH];
            B200_UNROLL
            for (int q = 0; q < NTT_STORE_BATCH; q++) {
                const uint32_t ee = e + q * nthreads;
                const uint32_t cw = ee & ((1u << g.log_cw) - 1), k = ee >> g.log_cw;
                tw[q] = fr_load(p.boundary_tw, ((unsigned long long)k << g.log_stride) + m0 + cw);
            }
            B200_UNROLL
            for (int q = 0; q < NTT_STORE_BATCH; q++) {
                uint32_t t, cw;
                ntt_store_source(g, e + q * nthreads, t, cw);
                ntt_store_element(p, g, bulk_load(T, bulk_idx(T, t, cw)), dst, tile, e + q * nthreads, true, tw[q]);
            }
        }
    }
    const fr_t none = fp_zero<FrP>();
    for (; e < g.tile_elems; e += nthreads) {
        uint32_t t, cw;
        ntt_store_source(g, e, t, cw);
        ntt_store_element(p, g, bulk_load(T, bulk_idx(T, t, cw)), dst, tile, e, false, none);
    }
}
