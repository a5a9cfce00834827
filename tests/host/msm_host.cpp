// CPU-side emulation shim (TEST ONLY) for the batched-affine bucket rounds in snarkos_b200/csrc/msm_affine.cuh.
// Runs the per-thread functions the CUDA kernels run (denominators -> batch inversion -> additions), thread by
// thread, around a plain counting sort, then finishes with the XYZZ code of msm_core.cuh.  Verifies the round
// schedule, the exceptional cases and the index arithmetic against the oracle without a GPU.  Not part of the
// shipped library; not a CPU fallback.
#include <cstdio>
#include <cstring>
#include <vector>

#include "../../snarkos_b200/csrc/msm_affine.cuh"
#include "../../snarkos_b200/csrc/msm_glv.cuh"

static void invert_all(std::vector<uint4>& data, size_t n) {
    if (n <= 64) {
        const uint32_t nthreads = 8;
        for (uint32_t t = 0; t < nthreads; t++) fq_inv_small_thread(data.data(), n, t, nthreads);
        return;
    }
    const size_t nt = (n + MSM_INV_CHUNK - 1) / MSM_INV_CHUNK;
    std::vector<uint4> partial(3 * nt);
    for (size_t t = 0; t < nt + 1; t++) fq_inv_up_thread(partial.data(), data.data(), n, t);
    invert_all(partial, nt);
    for (size_t t = 0; t < nt + 1; t++) fq_inv_down_thread(data.data(), partial.data(), n, t);
}

// out_jac: 144 B Jacobian.  pts: n affine images (x@0, y@48, infinity@96).  scalars: n x 8 u32 canonical.
// rounds: number of batched-affine rounds before the XYZZ finish.  stats[0..rounds): list length after each round.
extern "C" int host_msm_affine(uint8_t* out_jac, const uint8_t* pts, size_t n, size_t stride, const uint32_t* scalars,
                               uint32_t c, uint32_t rounds, uint32_t* stats) {
    const MsmShape sh = msm_shape(c);
    const uint32_t K = sh.nwin * sh.nbuckets;
    std::vector<uint4> packed((n ? n : 1) * G1_BASE_U4);        // 128-byte base records, pad filled with junk
    for (size_t i = 0; i < n; i++) {
        const uint8_t* s = pts + i * stride;
        memset(&packed[i * G1_BASE_U4], 0, 96);
        memset(&packed[i * G1_BASE_U4 + 6], 0xa5, 32);
        if (!s[96]) memcpy(&packed[i * G1_BASE_U4], s, 96);
    }
    std::vector<std::vector<uint32_t>> lists(K);
    for (size_t i = 0; i < n; i++) {
        uint32_t carry = 0;
        for (uint32_t w = 0; w < sh.nwin; w++) {
            uint32_t neg;
            uint32_t d = msm_signed_digit(scalars + 8 * i, w, c, carry, neg);
            if (d) lists[(size_t)w * sh.nbuckets + d - 1].push_back((uint32_t)i | (neg << 31));
        }
    }
    std::vector<uint32_t> entries, off(K + 1);
    for (uint32_t k = 0; k < K; k++) {
        off[k] = (uint32_t)entries.size();
        entries.insert(entries.end(), lists[k].begin(), lists[k].end());
    }
    off[K] = (uint32_t)entries.size();

    std::vector<uint4> cur_x, cur_y;        // planes of the current round's list (empty: still the gather form)
    bool gathered = true;
    for (uint32_t r = 0; r < rounds; r++) {
        std::vector<uint32_t> noff(K + 1);
        uint32_t tot = 0;
        for (uint32_t k = 0; k < K; k++) {
            noff[k] = tot;
            tot += (off[k + 1] - off[k] + 1) >> 1;
        }
        noff[K] = tot;
        PairRound rd;
        rd.src = gathered ? packed.data() : cur_x.data();
        rd.src_y = gathered ? nullptr : cur_y.data();
        rd.entries = gathered ? entries.data() : nullptr;
        rd.off = off.data();
        rd.noff = noff.data();
        rd.K = K;
        const uint32_t T = (tot + MSM_PAIRS_PER_THREAD - 1) / MSM_PAIRS_PER_THREAD + 2;     // + idle threads past the end
        std::vector<uint4> pre(3 * (size_t)(tot ? tot : 1)), partial(3 * (size_t)T);
        std::vector<uint4> next_x(3 * (size_t)(tot ? tot : 1)), next_y(3 * (size_t)(tot ? tot : 1));
        for (uint32_t t = 0; t < T; t++) pair_denoms_thread(rd, t, pre.data(), partial.data());
        invert_all(partial, T);
        for (uint32_t t = 0; t < T; t++) pair_add_thread(rd, t, pre.data(), partial.data(), next_x.data(), next_y.data());
        cur_x.swap(next_x);
        cur_y.swap(next_y);
        off.swap(noff);
        gathered = false;
        if (stats) stats[r] = tot;
    }
    std::vector<g1_xyzz_mem_t> buckets(K);
    for (uint32_t k = 0; k < K; k++) {
        g1_xyzz_t acc = g1_xyzz_infinity();
        for (uint32_t e = off[k]; e < off[k + 1]; e++) {
            g1_affine_t a;
            if (gathered) {
                a = g1_unpack(g1_load_packed(&packed[(size_t)(entries[e] & 0x7fffffffu) * G1_BASE_U4]));
                if (entries[e] >> 31) a.y = fp_neg(a.y);
            } else {
                a = g1_unpack(g1_load_planes(&cur_x[3 * (size_t)e], &cur_y[3 * (size_t)e]));
            }
            g1_madd(acc, a);
        }
        g1_xyzz_store(&buckets[k], acc);
    }
    std::vector<g1_xyzz_mem_t> wsum(sh.nwin);
    for (uint32_t w = 0; w < sh.nwin; w++)
        g1_xyzz_store(&wsum[w], msm_reduce_segment(buckets.data() + (size_t)w * sh.nbuckets, 0, sh.nbuckets));
    g1_xyzz_t total = msm_fold_windows(wsum.data(), sh.nwin, c);
    fq_t X, Y, Z;
    g1_xyzz_to_jacobian(total, X, Y, Z);
    memcpy(out_jac, X.v, 48);
    memcpy(out_jac + 48, Y.v, 48);
    memcpy(out_jac + 96, Z.v, 48);
    return 0;
}

// GLV split of n scalars (8 x u32 each) -> k1, k2 (4 x u32 each), and phi(P) = (beta x, y) of n affine points
extern "C" void host_glv_split(uint32_t* k1, uint32_t* k2, const uint32_t* k, size_t n) {
    for (size_t i = 0; i < n; i++) msm_glv_split(k + 8 * i, k1 + 4 * i, k2 + 4 * i);
}
extern "C" void host_glv_endo_x(uint32_t* out_x, const uint32_t* x, size_t n) {
    for (size_t i = 0; i < n; i++) {
        fq_t a;
        memcpy(a.v, x + 12 * i, 48);
        a = fp_mul(a, msm_glv_beta());
        memcpy(out_x + 12 * i, a.v, 48);
    }
}
