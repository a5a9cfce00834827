// CPU-side emulation shim (TEST ONLY) for the NTT pass code in snarkos_b200/csrc/ntt_core.cuh.
// Runs exactly the per-phase functions the CUDA kernel runs, for every (tile, batch, tid) in turn, with
// tables built by the same field code.  Verifies the pass decomposition / index arithmetic against the
// oracle without a GPU.  Not part of the shipped library; not a CPU fallback.
#include <cstdio>
#include <cstring>
#include <type_traits>
#include <vector>

#include "../../snarkos_b200/csrc/ntt_core.cuh"
#include "../../snarkos_b200/csrc/ntt_plan.h"

static fr_t fr_const(const uint32_t* limbs) { fr_t r; memcpy(r.v, limbs, sizeof(r.v)); return r; }

static std::vector<uint4> pow_table(const uint32_t* base_limbs, uint32_t nsq, uint32_t count, uint32_t shift) {
    std::vector<uint4> out(2 * (size_t)count);
    fr_t base = fr_const(base_limbs);
    for (uint32_t k = 0; k < nsq; k++) base = fp_sqr(base);
    // incremental: step = base^(1 << shift)
    fr_t step = base;
    for (uint32_t k = 0; k < shift; k++) step = fp_sqr(step);
    fr_t cur = fp_one<FrP>();
    for (uint32_t i = 0; i < count; i++) {
        fr_to_u4(cur, out[2 * i], out[2 * i + 1]);
        cur = fp_mul(cur, step);
    }
    return out;
}

extern "C" int host_ntt_plan(uint32_t log_n, uint32_t* npasses, uint32_t* log_len, uint32_t* log_cw) {
    NttPlan plan;
    if (!ntt_make_plan(log_n, &plan)) return -1;
    *npasses = plan.npasses;
    for (int i = 0; i < NTT_MAX_PASSES; i++) { log_len[i] = plan.log_len[i]; log_cw[i] = plan.log_cw[i]; }
    return 0;
}

// data: batch polynomials of 2^log_n Montgomery Fr (8 x u32 each), stride in elements.  `nthreads` emulated
// threads per CTA.  plan_len/plan_cw: explicit plan (npasses entries) or npasses = 0 to use ntt_make_plan.
// variant 0: the default kernel's phases (planar tile, per-thread loads); 1: the bulk-copy (TMA) variant's phases
// (interleaved tile filled run by run, as cp.async.bulk does; lanes 4..7 of a quarter-warp touch the high half first).
extern "C" int host_ntt_variant(uint32_t* data, uint32_t log_n, uint32_t batch, uint64_t stride, int direction, int coset,
                                uint32_t npasses, const uint32_t* plan_len, const uint32_t* plan_cw, uint32_t nthreads,
                                int use_tables, int variant);
extern "C" int host_ntt(uint32_t* data, uint32_t log_n, uint32_t batch, uint64_t stride, int direction, int coset,
                        uint32_t npasses, const uint32_t* plan_len, const uint32_t* plan_cw, uint32_t nthreads,
                        int use_tables) {
    return host_ntt_variant(data, log_n, batch, stride, direction, coset, npasses, plan_len, plan_cw, nthreads, use_tables, 0);
}
extern "C" int host_ntt_variant(uint32_t* data, uint32_t log_n, uint32_t batch, uint64_t stride, int direction, int coset,
                                uint32_t npasses, const uint32_t* plan_len, const uint32_t* plan_cw, uint32_t nthreads,
                                int use_tables, int variant) {
    if (log_n == 0) return 0;
    NttPlan plan;
    if (npasses == 0) {
        if (!ntt_make_plan(log_n, &plan)) return -1;
    } else {
        memset(&plan, 0, sizeof(plan));
        plan.npasses = npasses;
        for (uint32_t i = 0; i < npasses; i++) { plan.log_len[i] = plan_len[i]; plan.log_cw[i] = plan_cw[i]; }
    }
    const uint32_t* root = direction ? FR_TWO_ADIC_ROOT_INV : FR_TWO_ADIC_ROOT;
    const uint32_t* g = direction ? FR_GENERATOR_INV : FR_GENERATOR;
    uint32_t lo_count = 1u << NTT_POW_LO_LOG;
    uint32_t hi_count = log_n > NTT_POW_LO_LOG ? (1u << (log_n - NTT_POW_LO_LOG)) : 1u;
    auto tile_tw = pow_table(root, FR_TWO_ADICITY - NTT_TILE_TW_LOG, 1u << (NTT_TILE_TW_LOG - 1), 0);
    auto pow_lo = pow_table(root, FR_TWO_ADICITY - log_n, lo_count, 0);
    auto pow_hi = pow_table(root, FR_TWO_ADICITY - log_n, hi_count, NTT_POW_LO_LOG);
    auto coset_lo = pow_table(g, 0, lo_count, 0);
    auto coset_hi = pow_table(g, 0, hi_count, NTT_POW_LO_LOG);
    fr_t nn = fp_zero<FrP>();
    nn.v[log_n >> 5] = 1u << (log_n & 31);
    fr_t size_inv = fp_inv(fp_to_mont(nn));

    // per-pass boundary tables, exactly as ntt_boundary_table_kernel fills them (use_tables = 0 exercises the
    // two-level fallback)
    std::vector<std::vector<uint4>> boundary(NTT_MAX_PASSES);
    if (use_tables) {
        uint32_t before = 0;
        for (uint32_t i = 0; i + 1 < plan.npasses; i++) {
            uint32_t log_sub = log_n - before, log_stride = log_sub - plan.log_len[i];
            before += plan.log_len[i];
            size_t count = (size_t)1 << log_sub;
            boundary[i].resize(2 * count);
            for (size_t q = 0; q < count; q++) {
                unsigned long long k = q >> log_stride, m = q & ((1ull << log_stride) - 1);
                unsigned long long ex = ((m * k) & ((1ull << log_sub) - 1)) << (log_n - log_sub);
                fr_t w = pow2level(pow_lo.data(), pow_hi.data(), ex);
                fr_to_u4(w, boundary[i][2 * q], boundary[i][2 * q + 1]);
            }
        }
    }
    size_t total = ((size_t)(batch - 1) * stride + ((size_t)1 << log_n));
    std::vector<uint4> scratch(2 * total);
    uint4* d = reinterpret_cast<uint4*>(data);
    for (uint32_t i = 0; i < plan.npasses; i++) {
        NttPassParams p;
        memset(&p, 0, sizeof(p));
        bool first = (i == 0), last = (i + 1 == plan.npasses);
        std::vector<uint4> dst_copy;
        p.src = first ? d : scratch.data();
        p.dst = last ? d : scratch.data();
        p.tile_tw = tile_tw.data();
        p.pow_lo = pow_lo.data(); p.pow_hi = pow_hi.data();
        p.boundary_tw = boundary[i].empty() ? nullptr : boundary[i].data();
        p.coset_lo = coset_lo.data(); p.coset_hi = coset_hi.data();
        p.size_inv = size_inv;
        p.batch_stride = stride;
        p.log_n = log_n; p.pass = i; p.npasses = plan.npasses;
        for (int k = 0; k < NTT_MAX_PASSES; k++) p.log_len[k] = plan.log_len[k];
        p.log_cw = plan.log_cw[i];
        p.coset_pre = (first && coset && direction == 0);
        p.scale_post = (last && direction == 1);
        p.coset_post = (last && coset && direction == 1);
        p.radix4 = use_tables ? 1 : 0;           // the table-less run also exercises the radix-2-only path
        uint32_t tile_log = plan.log_len[i] + plan.log_cw[i];
        uint32_t tile_elems = 1u << tile_log;
        uint32_t ntiles = 1u << (log_n - tile_log);
        // a single-pass transform reads and writes the same buffer: like on the GPU, every tile is fully
        // loaded into "shared memory" before it is stored, and tiles of one pass touch disjoint outputs
        // only when src != dst or npasses == 1 (one tile).
        std::vector<uint4> sm(2 * (size_t)ntt_bulk_tile_elems(plan.log_len[i], plan.log_cw[i], last ? 1u : 0u));
        for (uint32_t b = 0; b < batch; b++)
            for (uint32_t tile = 0; tile < ntiles; tile++) {
                if (variant == 1) {
                    const uint4* srcp = p.src + 2ull * b * p.batch_stride;
                    for (uint32_t r = 0; r < ntt_bulk_runs(p); r++) {              // what the bulk copies do
                        unsigned long long se; uint32_t de, cnt;
                        ntt_bulk_run(p, tile, r, se, de, cnt);
                        memcpy(sm.data() + 2 * (size_t)de, srcp + 2 * se, (size_t)cnt * 32);
                    }
                    std::vector<uint4> smtw((size_t)1 << plan.log_len[i]);
                    for (uint32_t tid = 0; tid < nthreads; tid++) ntt_phase_stage_twiddles(p, smtw.data(), tid, nthreads);
                    const NttTwiddles twd = ntt_shared_twiddles(smtw.data(), plan.log_len[i]);
                    auto view = [&](uint32_t tid) {
                        NttBulkTile T;
                        T.sm = sm.data(); T.last = last ? 1u : 0u; T.log_len = plan.log_len[i]; T.log_cw = plan.log_cw[i];
                        T.h = (tid >> 2) & 1u;
                        return T;
                    };
                    if (p.coset_pre)
                        for (uint32_t tid = 0; tid < nthreads; tid++) ntt_bulk_coset_pre(p, view(tid), tile, tid, nthreads);
                    uint32_t s = 0;
                    for (; p.radix4 && s + 1 < plan.log_len[i]; s += 2)
                        for (uint32_t tid = 0; tid < nthreads; tid++) ntt_bulk_stage2(p, view(tid), twd, s, tid, nthreads);
                    for (; s < plan.log_len[i]; s++)
                        for (uint32_t tid = 0; tid < nthreads; tid++) ntt_bulk_stage(p, view(tid), twd, s, tid, nthreads);
                    for (uint32_t tid = 0; tid < nthreads; tid++) ntt_bulk_store_out(p, view(tid), tile, b, tid, nthreads);
                    continue;
                }
                if (variant == 2 && ntt_wc_applicable(p)) {
                    // what ntt_pass_wc_kernel does: cooperative fill (128 threads), per warp (= column) three rounds
                    // with all 32 lanes of a round before the next one, cooperative read-out
                    std::vector<uint4> wsm(WC_TILE_U4), smtw((size_t)1 << WC_LOG_LEN);
                    for (uint32_t tid = 0; tid < 128; tid++) ntt_phase_stage_twiddles(p, smtw.data(), tid, 128);
                    const NttTwiddles twd = ntt_shared_twiddles(smtw.data(), WC_LOG_LEN);
                    for (uint32_t tid = 0; tid < 128; tid++) wc_phase_load(p, wsm.data(), tile, b, tid, 128);
                    for (uint32_t cw = 0; cw < 4; cw++) {
                        for (uint32_t lane = 0; lane < 32; lane++) wc_round1(p, wsm.data(), twd, tile, cw, lane);
                        for (uint32_t lane = 0; lane < 32; lane++) wc_round2(wsm.data(), twd, cw, lane);
                        for (uint32_t lane = 0; lane < 32; lane++) wc_round3(wsm.data(), twd, cw, lane);
                    }
                    for (uint32_t tid = 0; tid < 128; tid++) wc_phase_store(p, wsm.data(), tile, b, tid, 128);
                    continue;
                }
                const bool s82 = plan.log_len[i] == 8 && plan.log_cw[i] == 2, s73 = plan.log_len[i] == 7 && plan.log_cw[i] == 3,
                           s64 = plan.log_len[i] == 6 && plan.log_cw[i] == 4, s92 = plan.log_len[i] == 9 && plan.log_cw[i] == 2,
                           s83 = plan.log_len[i] == 8 && plan.log_cw[i] == 3;
                if (variant == 3 && p.radix4 && (s82 || s73 || s64 || s92 || s83)) {
                    // what ntt_pass_shaped_kernel<L, CW, STORE> does (compile-time shape, 128 threads, the launcher's choice
                    // of the store mode)
                    const int store = (p.coset_pre || p.scale_post || p.coset_post) ? 0 : last ? (plan.npasses > 1 ? 2 : 0) : (p.boundary_tw ? 1 : 0);
                    std::vector<uint4> smtw((size_t)1 << plan.log_len[i]);
                    auto run_shaped = [&](auto shape) {
                        typedef decltype(shape) SH;
                        const uint32_t NT = SH::nthreads;
                        for (uint32_t tid = 0; tid < NT; tid++) ntt_phase_stage_twiddles<SH>(p, smtw.data(), tid, 0);
                        const NttTwiddles twd = ntt_shared_twiddles(smtw.data(), plan.log_len[i]);
                        for (uint32_t tid = 0; tid < NT; tid++) ntt_phase_load<SH>(p, sm.data(), tile, b, tid, 0);
                        if (SH::store == 0 && p.coset_pre)
                            for (uint32_t tid = 0; tid < NT; tid++) ntt_phase_coset_pre(p, sm.data(), tile, tid, NT);
                        uint32_t s = 0;
                        for (; s + 1 < plan.log_len[i]; s += 2)
                            for (uint32_t tid = 0; tid < NT; tid++) ntt_phase_stage2<SH>(p, sm.data(), twd, s, tid, 0);
                        for (; s < plan.log_len[i]; s++)
                            for (uint32_t tid = 0; tid < NT; tid++) ntt_phase_stage<SH>(p, sm.data(), twd, s, tid, 0);
                        for (uint32_t tid = 0; tid < NT; tid++) ntt_phase_store<SH>(p, sm.data(), tile, b, tid, 0);
                    };
                    auto by_store = [&](auto l, auto c) {
                        constexpr int LL = decltype(l)::value, CC = decltype(c)::value, NT = (1 << (LL + CC)) / 8;
                        if (store == 1) run_shaped(NttShape<LL, CC, NT, 1, 1>());
                        else if (store == 2) run_shaped(NttShape<LL, CC, NT, 1, 2>());
                        else run_shaped(NttShape<LL, CC, NT, 1, 0>());
                    };
                    if (s82) by_store(std::integral_constant<int, 8>(), std::integral_constant<int, 2>());
                    else if (s73) by_store(std::integral_constant<int, 7>(), std::integral_constant<int, 3>());
                    else if (s92) by_store(std::integral_constant<int, 9>(), std::integral_constant<int, 2>());
                    else if (s83) by_store(std::integral_constant<int, 8>(), std::integral_constant<int, 3>());
                    else by_store(std::integral_constant<int, 6>(), std::integral_constant<int, 4>());
                    continue;
                }
                for (uint32_t tid = 0; tid < nthreads; tid++) ntt_phase_load(p, sm.data(), tile, b, tid, nthreads);
                if (p.coset_pre)
                    for (uint32_t tid = 0; tid < nthreads; tid++) ntt_phase_coset_pre(p, sm.data(), tile, tid, nthreads);
                // odd tiles read the twiddles straight from the global table, even tiles through the staged copy, so both
                // accessors are exercised
                std::vector<uint4> smtw((size_t)1 << plan.log_len[i]);
                for (uint32_t tid = 0; tid < nthreads; tid++) ntt_phase_stage_twiddles(p, smtw.data(), tid, nthreads);
                NttTwiddles twd = (tile & 1) ? ntt_global_twiddles(p, plan.log_len[i]) : ntt_shared_twiddles(smtw.data(), plan.log_len[i]);
                uint32_t s = 0;
                for (; p.radix4 && s + 1 < plan.log_len[i]; s += 2)
                    for (uint32_t tid = 0; tid < nthreads; tid++) ntt_phase_stage2(p, sm.data(), twd, s, tid, nthreads);
                for (; s < plan.log_len[i]; s++)
                    for (uint32_t tid = 0; tid < nthreads; tid++) ntt_phase_stage(p, sm.data(), twd, s, tid, nthreads);
                for (uint32_t tid = 0; tid < nthreads; tid++) ntt_phase_store(p, sm.data(), tile, b, tid, nthreads);
            }
    }
    return 0;
}

// Host emulation of b200_ntt_rows_exchange_device (ntt_run_device_x with an exchange descriptor): `rows` rows of
// 2^log_len elements are transformed with the generic phases and the LAST pass stores into `world` destination slabs
// (dsts[d], each c_local x (rows * world) elements) exactly as the kernel does.
extern "C" int host_ntt_rows_exchange(const uint32_t* data, uint64_t rows, uint32_t log_len, uint32_t world, uint32_t rank,
                                      uint32_t log_n_total, int direction, int twiddle, uint64_t row_base, uint32_t** dsts,
                                      uint32_t nthreads) {
    NttPlan plan;
    const uint32_t tile_log = (log_len + 8) / 9 < (log_len + 7) / 8 ? 11 : 10;
    if (!ntt_make_plan(log_len, &plan, tile_log)) return -1;
    {
        const uint32_t ll = plan.log_len[plan.npasses - 1];
        uint32_t cw = ll >= 10 ? 0 : 10 - ll;
        if (cw > 4) cw = 4;
        while (cw && (rows & ((1u << cw) - 1))) cw--;
        plan.log_cw[plan.npasses - 1] = cw;
    }
    const uint32_t* root = direction ? FR_TWO_ADIC_ROOT_INV : FR_TWO_ADIC_ROOT;
    const uint32_t lo_count = 1u << NTT_POW_LO_LOG;
    auto hi_count = [&](uint32_t lg) { return lg > NTT_POW_LO_LOG ? (1u << (lg - NTT_POW_LO_LOG)) : 1u; };
    auto tile_tw = pow_table(root, FR_TWO_ADICITY - NTT_TILE_TW_LOG, 1u << (NTT_TILE_TW_LOG - 1), 0);
    auto pow_lo = pow_table(root, FR_TWO_ADICITY - log_len, lo_count, 0);
    auto pow_hi = pow_table(root, FR_TWO_ADICITY - log_len, hi_count(log_len), NTT_POW_LO_LOG);
    auto xpow_lo = pow_table(root, FR_TWO_ADICITY - log_n_total, lo_count, 0);
    auto xpow_hi = pow_table(root, FR_TWO_ADICITY - log_n_total, hi_count(log_n_total), NTT_POW_LO_LOG);
    fr_t nn = fp_zero<FrP>();
    nn.v[log_len >> 5] = 1u << (log_len & 31);
    const fr_t size_inv = fp_inv(fp_to_mont(nn));
    const size_t n = (size_t)1 << log_len;
    std::vector<uint4> scratch(2 * rows * n);
    const uint4* d = reinterpret_cast<const uint4*>(data);
    for (uint32_t i = 0; i < plan.npasses; i++) {
        NttPassParams p;
        memset(&p, 0, sizeof(p));
        const bool first = (i == 0), last = (i + 1 == plan.npasses);
        p.src = first ? d : scratch.data();
        p.dst = scratch.data();
        p.tile_tw = tile_tw.data();
        p.pow_lo = pow_lo.data(); p.pow_hi = pow_hi.data();
        p.size_inv = size_inv;
        p.batch_stride = n;
        p.log_n = log_len; p.pass = i; p.npasses = plan.npasses;
        for (int k = 0; k < NTT_MAX_PASSES; k++) p.log_len[k] = plan.log_len[k];
        p.log_cw = plan.log_cw[i];
        p.scale_post = (last && direction == 1);
        p.radix4 = 1;
        uint32_t ntiles = 1u << (log_len - plan.log_len[i] - plan.log_cw[i]);
        uint32_t nbatch = (uint32_t)rows;
        if (last) {
            p.xchg = 1; p.x_world = world; p.x_rank = rank; p.x_twiddle = twiddle ? 1 : 0; p.x_log_n = log_n_total;
            p.x_r_total = rows; p.x_row_base = row_base;
            p.x_pow_lo = xpow_lo.data(); p.x_pow_hi = xpow_hi.data();
            for (uint32_t w = 0; w < world; w++) p.x_dst[w] = reinterpret_cast<uint4*>(dsts[w]);
            ntiles = 1u << (log_len - plan.log_len[i]);
            nbatch = (uint32_t)(rows >> plan.log_cw[i]);
        }
        std::vector<uint4> sm(2 * ((size_t)1 << (plan.log_len[i] + plan.log_cw[i])));
        std::vector<uint4> smtw((size_t)1 << plan.log_len[i]);
        for (uint32_t b = 0; b < nbatch; b++)
            for (uint32_t tile = 0; tile < ntiles; tile++) {
                for (uint32_t tid = 0; tid < nthreads; tid++) ntt_phase_load(p, sm.data(), tile, b, tid, nthreads);
                for (uint32_t tid = 0; tid < nthreads; tid++) ntt_phase_stage_twiddles(p, smtw.data(), tid, nthreads);
                const NttTwiddles twd = ntt_shared_twiddles(smtw.data(), plan.log_len[i]);
                uint32_t s = 0;
                for (; s + 1 < plan.log_len[i]; s += 2)
                    for (uint32_t tid = 0; tid < nthreads; tid++) ntt_phase_stage2(p, sm.data(), twd, s, tid, nthreads);
                for (; s < plan.log_len[i]; s++)
                    for (uint32_t tid = 0; tid < nthreads; tid++) ntt_phase_stage(p, sm.data(), twd, s, tid, nthreads);
                for (uint32_t tid = 0; tid < nthreads; tid++) ntt_phase_store(p, sm.data(), tile, b, tid, nthreads);
            }
    }
    return 0;
}
