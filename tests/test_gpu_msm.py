"""GPU parity of VariableBase::msm through the C ABI against the oracle (bit-exact after affine normalisation,
SURVEY.md fact 5), golden vector, edge cases, resident bases, and an exact size-independent check at the
BASELINE sizes: with P_i = k_i * G the MSM must equal (sum_i s_i k_i mod r) * G."""
import numpy as np
import pytest

from oracle import bls12_377 as O
from oracle import c_oracle as C
from tests import helpers as H

pytestmark = pytest.mark.gpu
KAT = H.load_kat()


def gpu_msm(bases, scalars):
    import snarkos_b200 as S
    return H.jac_bytes_to_affine(S.VariableBase.msm(bases, scalars))


def oracle_msm(bases, scalars):
    return H.jac_bytes_to_affine(C.msm(bases, scalars))


def test_golden_vector():
    G = tuple(KAT["G"])
    got = gpu_msm(H.bases_array([G, tuple(KAT["2G"]), tuple(KAT["3G"])]), H.scalars_array([1, 2, 3]))
    assert got == tuple(KAT["MSM_1_2_3__G_2G_3G"])


@pytest.mark.parametrize("n", [0, 1, 2, 5, 14, 15, 33, 100, 1000])
def test_small_sizes_vs_naive(n):
    """mirrors snarkVM variable_base::tests::test_msm (naive vs msm for sizes 1..1000)"""
    rng = O.SplitMix64(500 + n)
    pts = O.random_points(rng, n) if n else []
    sc = O.random_fr(rng, n)
    want = O.msm_naive(pts, sc) if n <= 100 else oracle_msm(H.bases_array(pts), H.scalars_array(sc))
    assert gpu_msm(H.bases_array(pts), H.scalars_array(sc)) == want


def test_edge_cases():
    rng = O.SplitMix64(8)
    n = 64
    pts = O.random_points(rng, n)
    sc = O.random_fr(rng, n)
    sc[0], sc[1], sc[2], sc[3] = 0, 1, O.R_MOD - 1, O.R_MOD - 2
    pts[4] = None                                   # point at infinity inside bases
    pts[6], sc[6] = pts[5], sc[5]                   # duplicate (point, scalar): P + P in a bucket
    pts[8], sc[8] = O.g1_neg(pts[7]), sc[7]         # P + (-P) in a bucket
    assert gpu_msm(H.bases_array(pts), H.scalars_array(sc)) == O.msm_naive(pts, sc)
    # all scalars equal / all points equal / all zero
    assert gpu_msm(H.bases_array(pts), H.scalars_array([sc[9]] * n)) == O.msm_naive(pts, [sc[9]] * n)
    assert gpu_msm(H.bases_array([pts[0]] * n), H.scalars_array(sc)) == O.g1_mul(pts[0], sum(sc) % O.R_MOD)
    assert gpu_msm(H.bases_array(pts), H.scalars_array([0] * n)) is None
    assert gpu_msm(H.bases_array([None] * n), H.scalars_array(sc)) is None
    # result exactly infinity from non-trivial terms
    assert gpu_msm(H.bases_array([pts[0], pts[0]]), H.scalars_array([5, O.R_MOD - 5])) is None
    # fewer scalars than bases: min(len) like snarkVM
    assert gpu_msm(H.bases_array(pts), H.scalars_array(sc[:10])) == O.msm_naive(pts[:10], sc[:10])


def _synthetic(n, seed):
    import torch
    import snarkos_b200 as S
    dev = S.synthetic_bases(n, seed=seed)
    torch.cuda.synchronize()
    return dev


@pytest.mark.parametrize("log_n", [10, 12, 16])
def test_vs_c_oracle_pippenger(log_n):
    """BASELINE config 1 (2^16 random bases / scalars): GPU vs the CPU restatement, bit-exact in affine"""
    n = 1 << log_n
    bases = _synthetic(n, 42).cpu().numpy()
    sc = H.random_scalars_np(np.random.default_rng(log_n), n)
    assert gpu_msm(bases, sc) == oracle_msm(bases, sc)


def test_resident_bases_and_device_path():
    import torch
    import snarkos_b200 as S
    n = 1 << 12
    dbases = _synthetic(n, 7)
    sc = H.random_scalars_np(np.random.default_rng(3), n)
    want = oracle_msm(dbases.cpu().numpy(), sc)
    rb = S.ResidentBases(dbases)
    assert H.jac_bytes_to_affine(rb.msm(sc)) == want
    dsc = torch.from_numpy(sc.view(np.int64)).cuda()
    out = rb.msm(dsc)
    torch.cuda.synchronize()
    assert H.jac_bytes_to_affine(out.cpu().numpy()) == want
    # prefix of the registered set (KZG commit of a lower-degree polynomial)
    assert H.jac_bytes_to_affine(rb.msm(sc[:1000])) == oracle_msm(dbases.cpu().numpy()[:1000 * 104], sc[:1000])
    out = S.VariableBase.msm(dbases, dsc)
    torch.cuda.synchronize()
    assert H.jac_bytes_to_affine(out.cpu().numpy()) == want
    rb.release()
    with pytest.raises(ValueError):
        rb.msm(sc)
    # partial-sum combine used by the multi-GPU path
    halves = torch.stack([S.VariableBase.msm(dbases[:n // 2 * 104], dsc[:n // 2]),
                          S.VariableBase.msm(dbases[n // 2 * 104:], dsc[n // 2:])])
    tot = S.sum_projective(halves)
    torch.cuda.synchronize()
    assert H.jac_bytes_to_affine(tot.cpu().numpy()) == want


@pytest.mark.parametrize("kind", ["all_equal", "two_values", "tiny", "top_heavy"])
def test_skewed_scalar_distributions(kind):
    """Heavily loaded buckets (one bucket receiving most points) must go through the task split + combine rounds."""
    import torch
    import snarkos_b200 as S
    n, seed = 1 << 14, 21
    dbases = _synthetic(n, seed)
    rng = np.random.default_rng(4)
    if kind == "all_equal":
        vals = [O.R_MOD - 12345] * n
    elif kind == "two_values":
        vals = [3 if i % 3 else (1 << 200) + 7 for i in range(n)]
    elif kind == "tiny":
        vals = [int(x) for x in rng.integers(0, 4, size=n)]
    else:
        vals = [(1 << 252) + int(x) for x in rng.integers(0, 1 << 20, size=n)]
    sc = H.scalars_array(vals)
    out = S.VariableBase.msm(dbases, torch.from_numpy(sc.view(np.int64)).cuda())
    torch.cuda.synchronize()
    k = H.splitmix64_at(seed, np.arange(n))
    assert H.jac_bytes_to_affine(out.cpu().numpy()) == O.g1_mul(O.G1_GEN, H.dot_mod_r(sc, k))


@pytest.mark.parametrize("log_n", [20, 24])
def test_full_size_exact_identity(log_n):
    """size-independent exact check at BASELINE scale: P_i = k_i G  =>  MSM = (sum s_i k_i mod r) G"""
    import torch
    import snarkos_b200 as S
    n, seed = 1 << log_n, 99
    dbases = _synthetic(n, seed)
    sc = H.random_scalars_np(np.random.default_rng(log_n), n)
    sc[:4] = H.scalars_array([0, 1, O.R_MOD - 1, O.R_MOD - 2])
    dsc = torch.from_numpy(sc.view(np.int64)).cuda()
    out = S.VariableBase.msm(dbases, dsc)
    torch.cuda.synchronize()
    k = H.splitmix64_at(seed, np.arange(n))
    assert H.jac_bytes_to_affine(out.cpu().numpy()) == O.g1_mul(O.G1_GEN, H.dot_mod_r(sc, k))


def test_kzg_commit_resident_powers():
    """KZG10::commit over resident powers: Montgomery coefficients in, conversion on the device, zero coefficients
    (leading and interior) contribute nothing, hiding = second MSM on the gamma powers added."""
    import torch
    import snarkos_b200 as S
    n = 1 << 12
    g_powers = _synthetic(n, 31)
    gamma_powers = _synthetic(64, 32)
    powers = S.Powers(g_powers, gamma_powers)
    rng = np.random.default_rng(9)
    coeffs_mont = H.random_fr_mont_np(rng, (3000,))
    coeffs_mont[:17] = 0                      # leading zeros (skip_leading_zeros_and_convert_to_bigints)
    coeffs_mont[100:140] = 0
    blind_mont = H.random_fr_mont_np(rng, (40,))
    plain = C.fr_from_mont(coeffs_mont)
    plain_b = C.fr_from_mont(blind_mont)
    hb, hg = g_powers.cpu().numpy(), gamma_powers.cpu().numpy()
    want = oracle_msm(hb[:3000 * 104], plain)
    assert H.jac_bytes_to_affine(S.KZG10.commit(powers, coeffs_mont)) == want
    dev = S.KZG10.commit(powers, torch.from_numpy(coeffs_mont.view(np.int64)).cuda())
    torch.cuda.synchronize()
    assert H.jac_bytes_to_affine(dev.cpu().numpy()) == want
    want_h = O.g1_add(want, oracle_msm(hg[:40 * 104], plain_b))
    assert H.jac_bytes_to_affine(S.KZG10.commit(powers, coeffs_mont, blind_mont)) == want_h
    with pytest.raises(S.B200Error):
        S.KZG10.commit(powers, H.random_fr_mont_np(rng, (n + 1,)))
    powers.release()


@pytest.mark.parametrize("n,auto_table", [(1 << 12, 1), (1 << 12, 0), (1 << 9, 1)])
def test_kzg_commit_batch_shared_powers(n, auto_table):
    """k polynomials of different lengths against ONE resident set in one launch set == each commit on its own (and the
    oracle): the tabulated path (auto table at 2^10..2^20), the plain packed path and a set below the table threshold;
    zero polynomials, empty polynomials, repeated-digit coefficients and a length equal to the whole set included."""
    import torch
    import snarkos_b200 as S
    S.set_option("msm_auto_table", auto_table)
    try:
        g_powers = _synthetic(n, 41)
        powers = S.Powers(g_powers)
    finally:
        S.set_option("msm_auto_table", 1)
    rng = np.random.default_rng(19)
    lens = [n, 700 if n > 700 else n // 2, 0, 1, 333, n - 1, 64]
    polys = [H.random_fr_mont_np(rng, (L,)) for L in lens]
    polys[4][:] = 0                                                   # the zero polynomial
    polys[1][:50] = 0                                                 # leading zeros
    one = H.random_fr_mont_np(rng, (1,))
    polys[6][:] = one                                                 # all coefficients equal: one bucket per window
    hb = g_powers.cpu().numpy()
    want = [oracle_msm(hb[:L * 104], C.fr_from_mont(p)) if L else None for L, p in zip(lens, polys)]
    got = S.KZG10.commit_batch(powers, polys)
    assert got.shape == (len(lens), 144)
    for m in range(len(lens)):
        assert H.jac_bytes_to_affine(got[m]) == want[m], f"polynomial {m} (len {lens[m]})"
        assert H.jac_bytes_to_affine(S.KZG10.commit(powers, polys[m])) == want[m]
    dev = S.KZG10.commit_batch(powers, [torch.from_numpy(p.view(np.int64)).cuda() for p in polys])
    torch.cuda.synchronize()
    dev = dev.cpu().numpy()
    for m in range(len(lens)):
        assert H.jac_bytes_to_affine(dev[m]) == want[m]
    with pytest.raises(S.B200Error):
        S.KZG10.commit_batch(powers, [polys[0], H.random_fr_mont_np(rng, (n + 1,))])
    powers.release()


def test_batched_small_msms():
    """config 4 shape: many independent small MSMs (tens of points each) in one call == each one on its own"""
    import snarkos_b200 as S
    rng = O.SplitMix64(77)
    sizes = [0, 1, 7, 33, 64, 2, 0, 51, 100, 18] * 3
    pts, sc, off = [], [], [0]
    for k in sizes:
        pts += O.random_points(rng, k) if k else []
        sc += O.random_fr(rng, k)
        off.append(off[-1] + k)
    if len(pts) > 5:
        pts[3] = None
        sc[4] = 0
    bases, scal = H.bases_array(pts), H.scalars_array(sc)
    out = S.msm_batch(bases, scal, off)
    assert out.shape == (len(sizes), 144)
    for m, k in enumerate(sizes):
        lo, hi = off[m], off[m + 1]
        assert H.jac_bytes_to_affine(out[m]) == O.msm_naive(pts[lo:hi], sc[lo:hi]), m
    # 256 transactions x 40 points against per-call results
    n_tx, per = 256, 40
    dev = _synthetic(n_tx * per, 55).cpu().numpy()
    sc2 = H.random_scalars_np(np.random.default_rng(8), n_tx * per)
    off2 = np.arange(n_tx + 1, dtype=np.uint64) * per
    out2 = S.msm_batch(dev, sc2, off2)
    for m in (0, 1, 100, 255):
        want = oracle_msm(dev[m * per * 104:(m + 1) * per * 104], sc2[m * per:(m + 1) * per])
        assert H.jac_bytes_to_affine(out2[m]) == want


@pytest.mark.parametrize("window_bits", [0, 16, 19])
def test_tabulated_resident_bases(window_bits):
    """resident bases with their window table 2^(c*w) * P_i: same group element as the plain path, for full and
    prefix-length scalar vectors, including points at infinity and extreme scalars"""
    import torch
    import snarkos_b200 as S
    n = 1 << 11
    dbases = _synthetic(n, 91)
    hb = dbases.cpu().numpy().copy()
    hb[5 * 104:6 * 104] = 0
    hb[5 * 104 + 96] = 1                                  # a point at infinity inside the set
    sc = H.random_scalars_np(np.random.default_rng(12), n)
    sc[:4] = H.scalars_array([0, 1, O.R_MOD - 1, O.R_MOD - 2])
    rb = S.ResidentBases(hb, tabulate=True, window_bits=window_bits)
    assert H.jac_bytes_to_affine(rb.msm(sc)) == oracle_msm(hb, sc)
    assert H.jac_bytes_to_affine(rb.msm(sc[:700])) == oracle_msm(hb[:700 * 104], sc[:700])
    out = rb.msm(torch.from_numpy(sc.view(np.int64)).cuda())
    torch.cuda.synchronize()
    assert H.jac_bytes_to_affine(out.cpu().numpy()) == oracle_msm(hb, sc)
    rb.release()


def test_host_buffer_streaming_path():
    """host-buffer calls with >= 2^23 points stream geometrically growing point ranges (2^20, 2^20, 2^21, ...) into one bucket array (copy of range i+1
    overlapped with the accumulation of range i); exact identity check incl. a ragged last range"""
    import torch
    import snarkos_b200 as S
    n, seed = (1 << 23) + 12345, 5
    hb = _synthetic(n, seed).cpu().numpy()
    sc = H.random_scalars_np(np.random.default_rng(23), n)
    sc[:3] = H.scalars_array([0, O.R_MOD - 1, 1])
    got = S.VariableBase.msm(hb, sc)
    k = H.splitmix64_at(seed, np.arange(n))
    assert H.jac_bytes_to_affine(got) == O.g1_mul(O.G1_GEN, H.dot_mod_r(sc, k))


@pytest.mark.parametrize("two,rounds", [(1, -1), (1, 3), (0, -1)])
@pytest.mark.parametrize("n,c,first_log,chunk_log", [(5000, 6, 10, 11), (3000, 4, 10, 10), (70000, 9, 12, 14)])
def test_streamed_ranges_small(b200_opt, two, rounds, n, c, first_log, chunk_log):
    """the streamed host path at sizes the oracle finishes: ranges 2^first, 2^first, ... <= 2^chunk through three staging
    buffers, alternating between the thread's two compute streams (two = 1); the first range of a stream fills a bucket
    array, later ranges are accumulated into it (msm_accumulate_kernel<.., ADD>), the two streams' arrays are added at the finish"""
    b200_opt("msm_host_chunk_log", chunk_log)
    b200_opt("msm_host_first_log", first_log)
    b200_opt("msm_window_bits", c)
    b200_opt("msm_stream_two", two)
    b200_opt("msm_affine_rounds", rounds)
    rng = O.SplitMix64(7700 + n + c)
    pts = O.random_points(rng, 48)
    pts[5] = None
    pts = [pts[(i * 7) % len(pts)] for i in range(n)]           # repeated points: doublings inside and across ranges
    sc = O.random_fr(rng, n)
    sc[0], sc[1], sc[2] = 0, O.R_MOD - 1, 1
    sc[100:140] = [sc[99]] * 40                                   # a run of equal scalars across pair boundaries
    bases, scal = H.bases_array(pts), H.scalars_array(sc)
    assert gpu_msm(bases, scal) == oracle_msm(bases, scal)


@pytest.mark.parametrize("seg_len,quad_max", [(1, 0), (3, 0), (7, 1 << 30), (49, 0), (49, 1 << 30), (100000, 0)])
def test_bucket_reduction_segment_lengths(b200_opt, seg_len, quad_max):
    """the running-sum reduction for segment lengths that do not divide the bucket count (the one-wave rule picks
    ceil(K / resident threads), 49 at 2^24), with one lane and with a quad of lanes per segment"""
    b200_opt("msm_seg_len", seg_len)
    b200_opt("msm_reduce_quad_max", quad_max)
    b200_opt("msm_window_bits", 9)
    rng = O.SplitMix64(4242 + seg_len)
    n = 3000
    pts = O.random_points(rng, 40)
    pts = [pts[(i * 11) % len(pts)] for i in range(n)]
    sc = O.random_fr(rng, n)
    sc[0], sc[1] = 0, O.R_MOD - 1
    bases, scal = H.bases_array(pts), H.scalars_array(sc)
    assert gpu_msm(bases, scal) == oracle_msm(bases, scal)


def test_streamed_ranges_degenerate(b200_opt):
    """all scalars equal: ONE bucket per window holds every point of every range; all points equal: every pair is a
    doubling"""
    b200_opt("msm_host_chunk_log", 11)
    b200_opt("msm_host_first_log", 10)
    b200_opt("msm_window_bits", 5)
    rng = O.SplitMix64(99)
    n = 6000
    pts = O.random_points(rng, 32)
    pts = [pts[i % 32] for i in range(n)]
    s0 = O.random_fr(rng, 1)[0]
    bases = H.bases_array(pts)
    assert gpu_msm(bases, H.scalars_array([s0] * n)) == oracle_msm(bases, H.scalars_array([s0] * n))
    sc = O.random_fr(rng, n)
    assert gpu_msm(H.bases_array([pts[0]] * n), H.scalars_array(sc)) == O.g1_mul(pts[0], sum(sc) % O.R_MOD)
    assert gpu_msm(bases, H.scalars_array([0] * n)) is None


# ---------------------------------------------------------------------------------------------
# batched-affine pair rounds (msm_affine.cuh), forced on at small sizes through the msm_affine_rounds option
# ---------------------------------------------------------------------------------------------
@pytest.mark.parametrize("rounds", [1, 3, 8])
@pytest.mark.parametrize("n,c", [(1, 4), (33, 4), (1000, 6), (5000, 8)])
def test_affine_rounds_vs_oracle(b200_opt, rounds, n, c):
    """snarkVM batched::batch_add counterpart: pairwise affine additions with one batch inversion per round;
    narrow windows make buckets of tens to hundreds of points, `rounds` beyond log2(bucket size) must be no-ops."""
    b200_opt("msm_affine_rounds", rounds)
    b200_opt("msm_window_bits", c)
    rng = O.SplitMix64(1500 + n + rounds)
    pts = O.random_points(rng, min(n, 64))
    pts = [pts[i % len(pts)] for i in range(n)]               # repeated points: doublings appear in later rounds
    sc = O.random_fr(rng, n)
    bases, scal = H.bases_array(pts), H.scalars_array(sc)
    want = O.msm_naive(pts, sc) if n <= 33 else oracle_msm(bases, scal)
    assert gpu_msm(bases, scal) == want


@pytest.mark.parametrize("rounds", [2, 7])
def test_affine_rounds_edge_cases(b200_opt, rounds):
    b200_opt("msm_affine_rounds", rounds)
    b200_opt("msm_window_bits", 4)
    rng = O.SplitMix64(81)
    n = 64
    pts = O.random_points(rng, n)
    sc = O.random_fr(rng, n)
    sc[0], sc[1], sc[2], sc[3] = 0, 1, O.R_MOD - 1, O.R_MOD - 2
    pts[4] = None
    pts[6], sc[6] = pts[5], sc[5]                   # P + P as one pair
    pts[8], sc[8] = O.g1_neg(pts[7]), sc[7]         # P + (-P) as one pair
    pts[10], sc[10] = pts[9], O.R_MOD - sc[9]       # cancels through the sign bit
    assert gpu_msm(H.bases_array(pts), H.scalars_array(sc)) == O.msm_naive(pts, sc)
    assert gpu_msm(H.bases_array(pts), H.scalars_array([sc[11]] * n)) == O.msm_naive(pts, [sc[11]] * n)
    assert gpu_msm(H.bases_array([pts[0]] * n), H.scalars_array(sc)) == O.g1_mul(pts[0], sum(sc) % O.R_MOD)
    assert gpu_msm(H.bases_array([pts[0]] * n), H.scalars_array([sc[12]] * n)) == O.g1_mul(pts[0], n * sc[12] % O.R_MOD)
    assert gpu_msm(H.bases_array(pts), H.scalars_array([0] * n)) is None
    assert gpu_msm(H.bases_array([None] * n), H.scalars_array(sc)) is None
    assert gpu_msm(H.bases_array([pts[0], pts[0]]), H.scalars_array([5, O.R_MOD - 5])) is None


@pytest.mark.parametrize("kind", ["all_equal", "two_values", "tiny", "top_heavy", "uniform"])
def test_affine_rounds_skewed_distributions(b200_opt, kind):
    """2^16 points, 4 rounds: a bucket holding every point is halved four times and finished by the XYZZ chunks."""
    import torch
    import snarkos_b200 as S
    b200_opt("msm_affine_rounds", 4)
    n, seed = 1 << 16, 23
    dbases = _synthetic(n, seed)
    rng = np.random.default_rng(5)
    if kind == "all_equal":
        vals = [O.R_MOD - 12345] * n
    elif kind == "two_values":
        vals = [3 if i % 3 else (1 << 200) + 7 for i in range(n)]
    elif kind == "tiny":
        vals = [int(x) for x in rng.integers(0, 4, size=n)]
    elif kind == "top_heavy":
        vals = [(1 << 252) + int(x) for x in rng.integers(0, 1 << 20, size=n)]
    else:
        vals = None
    sc = H.scalars_array(vals) if vals is not None else H.random_scalars_np(rng, n)
    out = S.VariableBase.msm(dbases, torch.from_numpy(sc.view(np.int64)).cuda())
    torch.cuda.synchronize()
    k = H.splitmix64_at(seed, np.arange(n))
    assert H.jac_bytes_to_affine(out.cpu().numpy()) == O.g1_mul(O.G1_GEN, H.dot_mod_r(sc, k))


def test_affine_rounds_sliced_on_helper_stream(b200_opt):
    """msm_slices > 1: denominators + inversion of slice i + 1 on the thread's high-priority helper stream under
    the additions of slice i (kept as a measured alternative, default off) -- same result"""
    import torch
    import snarkos_b200 as S
    b200_opt("msm_affine_rounds", 3)
    b200_opt("msm_slices", 3)
    n, seed = 1 << 20, 31
    dbases = _synthetic(n, seed)
    sc = H.random_scalars_np(np.random.default_rng(6), n)
    out = S.VariableBase.msm(dbases, torch.from_numpy(sc.view(np.int64)).cuda())
    torch.cuda.synchronize()
    k = H.splitmix64_at(seed, np.arange(n))
    assert H.jac_bytes_to_affine(out.cpu().numpy()) == O.g1_mul(O.G1_GEN, H.dot_mod_r(sc, k))


@pytest.mark.parametrize("n", [1, 3, 100, 4097])
def test_glv_split_path_equals_plain_path(b200_opt, n):
    """VariableBase::msm splits every scalar by the endomorphism (k = k1 + k2 * lambda, phi(P) = (beta x, y): 2n points,
    127-bit scalars, half the windows); with msm_glv = 0 the same call runs on the 253-bit scalars -- both must give
    the oracle's point, including infinity among the bases and the extreme scalars"""
    rng = O.SplitMix64(3100 + n)
    pts = O.random_points(rng, min(n, 32))
    pts = [pts[i % len(pts)] for i in range(n)]
    sc = O.random_fr(rng, n)
    sc[0] = O.R_MOD - 1
    if n > 2:
        pts[1], sc[2] = None, 0
    bases, scal = H.bases_array(pts), H.scalars_array(sc)
    want = O.msm_naive(pts, sc) if n <= 100 else oracle_msm(bases, scal)
    assert gpu_msm(bases, scal) == want
    b200_opt("msm_glv", 0)
    assert gpu_msm(bases, scal) == want


def test_resident_set_in_glv_form(b200_opt):
    """resident sets above 2^20 points (and smaller ones with the window table switched off) are stored as
    (P_i, phi(P_i)) pairs and every MSM against them runs on the GLV halves: exact identity at 2^20 + 77 points, a
    prefix call against the oracle, KZG commit on a small GLV-resident set"""
    import torch
    import snarkos_b200 as S
    n, seed = (1 << 20) + 77, 41
    dbases = _synthetic(n, seed)
    sc = H.random_scalars_np(np.random.default_rng(9), n)
    sc[:3] = H.scalars_array([O.R_MOD - 1, 0, 1])
    rb = S.ResidentBases(dbases)
    out = rb.msm(torch.from_numpy(sc.view(np.int64)).cuda())
    torch.cuda.synchronize()
    k = H.splitmix64_at(seed, np.arange(n))
    assert H.jac_bytes_to_affine(out.cpu().numpy()) == O.g1_mul(O.G1_GEN, H.dot_mod_r(sc, k))
    hb = dbases[:600 * 104].cpu().numpy()
    assert H.jac_bytes_to_affine(rb.msm(sc[:600])) == oracle_msm(hb, sc[:600])
    rb.release()
    b200_opt("msm_auto_table", 0)
    small = S.ResidentBases(dbases[:4096 * 104])
    assert H.jac_bytes_to_affine(small.msm(sc[:4096])) == oracle_msm(dbases[:4096 * 104].cpu().numpy(), sc[:4096])
    small.release()


# ---------------------------------------------------------------------------------------------
# round-2 regressions: head-combine bound with tabulated bases, non-canonical scalars, forced XYZZ fallback, 2^26
# ---------------------------------------------------------------------------------------------
@pytest.mark.parametrize("kind", ["all_equal", "repeated_digit", "three_digits"])
@pytest.mark.parametrize("api", ["registered", "kzg_commit"])
def test_tabulated_structured_scalars(kind, api):
    """With tabulated bases every window of every point lands in ONE bucket set, so a bucket can hold n * nwin entries:
    structured scalars (every 16-bit digit equal) pile 3n .. 16n entries on one bucket and need head-combine strides far
    above n / chunk (the loop bound assumed at most n entries per bucket and dropped partial sums)."""
    import snarkos_b200 as S
    n = 1 << 10                                            # auto-tabulated, c = 16, chunk = 16
    dbases = _synthetic(n, 123)
    hb = dbases.cpu().numpy()
    if kind == "all_equal":
        vals = [0x0001000100010001000100010001000100010001000100010001000100010001 % O.R_MOD] * n
    elif kind == "repeated_digit":
        vals = [0x000100010001] * n                        # the advisor's example: 3n entries in bucket 0
    else:
        vals = [(1 << 32) + (1 << 16) + 1 + ((i % 2) << 48) for i in range(n)]
    sc = H.scalars_array(vals)
    want = O.g1_mul(O.G1_GEN, H.dot_mod_r(sc, H.splitmix64_at(123, np.arange(n))))
    if api == "registered":
        rb = S.ResidentBases(hb)                           # 2^10 points: tabulated automatically
        assert H.jac_bytes_to_affine(rb.msm(sc)) == want
        rb.release()
    else:
        powers = S.Powers(dbases, None)
        assert H.jac_bytes_to_affine(S.KZG10.commit(powers, C.fr_to_mont(sc))) == want
        powers.release()


def test_non_canonical_scalars_are_reduced(b200_opt):
    """scalars cross the ABI canonical (< r); a 256-bit k >= r still means (k mod r) * P and must not be truncated by
    the 127-bit GLV halves or the 253-bit window count"""
    rng = O.SplitMix64(4242)
    n = 40
    pts = O.random_points(rng, n)
    vals = O.random_fr(rng, n)
    vals[0], vals[1], vals[2], vals[3] = O.R_MOD, O.R_MOD + 1, (1 << 256) - 1, 13 * O.R_MOD + 5
    vals[4] = (1 << 255) + 12345
    want = O.msm_naive(pts, [v % O.R_MOD for v in vals])
    bases, scal = H.bases_array(pts), H.scalars_array(vals)
    assert gpu_msm(bases, scal) == want
    b200_opt("msm_glv", 0)
    assert gpu_msm(bases, scal) == want


def test_forced_xyzz_fallback_when_lists_do_not_fit(b200_opt):
    """the pair-round lists of a large call are ~25 GiB at 2^24; when they cannot be allocated the call must run the
    XYZZ-only accumulation, say so through a counter, and return the same point.  The allocation failure is forced
    through the msm_list_budget_bytes option."""
    import torch
    import snarkos_b200 as S
    n, seed = 1 << 20, 77
    dbases = _synthetic(n, seed)
    sc = H.random_scalars_np(np.random.default_rng(31), n)
    dsc = torch.from_numpy(sc.view(np.int64)).cuda()
    want = O.g1_mul(O.G1_GEN, H.dot_mod_r(sc, H.splitmix64_at(seed, np.arange(n))))
    b200_opt("msm_affine_rounds", 3)
    before = S.counter("msm_xyzz_fallbacks")
    out = S.VariableBase.msm(dbases, dsc)
    torch.cuda.synchronize()
    assert S.counter("msm_xyzz_fallbacks") == before
    assert H.jac_bytes_to_affine(out.cpu().numpy()) == want
    b200_opt("msm_list_budget_bytes", 1 << 20)
    out = S.VariableBase.msm(dbases, dsc)
    torch.cuda.synchronize()
    assert S.counter("msm_xyzz_fallbacks") == before + 1
    assert H.jac_bytes_to_affine(out.cpu().numpy()) == want


def test_single_gpu_2_26_exact_identity():
    """north_star upper size on ONE GPU: 2^26 points (6.5 GiB of bases, 2 GiB of scalars, ~60 GiB of pair-round lists)"""
    import torch
    import snarkos_b200 as S
    n, seed = 1 << 26, 2026
    dbases = _synthetic(n, seed)
    gen = torch.Generator(device="cuda")
    gen.manual_seed(5)
    dsc = torch.randint(-(1 << 63), (1 << 63) - 1, (n, 4), dtype=torch.int64, device="cuda", generator=gen)
    dsc[:, 3] &= (1 << 60) - 1
    out = S.VariableBase.msm(dbases, dsc)
    torch.cuda.synchronize()
    sc = dsc.cpu().numpy().view(np.uint64)
    del dbases, dsc
    torch.cuda.empty_cache()
    k = H.splitmix64_at(seed, np.arange(n))
    assert H.jac_bytes_to_affine(out.cpu().numpy()) == O.g1_mul(O.G1_GEN, H.dot_mod_r(sc, k))


# ---------------------------------------------------------------------------------------------------------------------
# output forms of a commitment (a8 / f4) and KZG10::open (a9)
# ---------------------------------------------------------------------------------------------------------------------
def test_batch_normalize_and_compress():
    """Projective::batch_normalization / to_affine and the compressed encoding: device == C oracle == Python oracle"""
    import torch
    import snarkos_b200 as S
    n = 100
    bases = _synthetic(n, 5).cpu().numpy().reshape(n, 104)
    k = np.arange(3, 3 + n, dtype=np.uint64)
    # non-trivial Jacobian representatives: k_i * P_i through the device's XYZZ ladder (Z != 1)
    L = S.lib()
    jac = np.zeros((n, 144), dtype=np.uint8)
    from snarkos_b200 import _lib
    import ctypes
    _lib.check(L.b200_debug_g1_op(2, jac.ctypes.data_as(ctypes.c_void_p), bases.ctypes.data_as(ctypes.c_void_p),
                                  k.ctypes.data_as(ctypes.c_void_p), n, 104))
    jac[17] = 0                                                        # Z = 0: infinity
    jac[17, 0] = 1
    jac[40, 96:] = 0                                                   # another infinity, inside a chunk of 16
    for stride in (104, 112):
        got = S.g1_batch_normalize(jac, stride=stride)
        want = C.g1_to_affine(jac, stride)
        for i in range(n):
            if want[i, 96]:
                assert got[i, 96] == 1
            else:
                assert bytes(got[i]) == bytes(want[i]), i
    comp = S.g1_compress(jac)
    assert np.array_equal(comp, C.g1_compress(jac))
    for i in (0, 1, 17, 40, 99):
        assert bytes(comp[i]) == O.g1_compress(H.jac_bytes_to_affine(jac[i]))
        assert O.g1_decompress(bytes(comp[i])) == H.jac_bytes_to_affine(jac[i])
    dcomp = S.g1_compress(torch.from_numpy(jac).cuda())
    torch.cuda.synchronize()
    assert np.array_equal(dcomp.cpu().numpy(), comp)
    assert S.g1_compress(np.zeros((0, 144), dtype=np.uint8)).shape == (0, 48)


@pytest.mark.parametrize("n", [1, 2, 100, 3000])
def test_kzg_open_matches_witness_commitment(n):
    """KZG10::open: w = commit((p - p(z)) / (X - z)), evaluation p(z); with a blinding polynomial the gamma-power part
    is added.  Also the KZG identity p(beta) - p(z) = w(beta) * (beta - z) in the exponent (powers = beta^i * G)."""
    import snarkos_b200 as S
    rng = O.SplitMix64(1000 + n)
    beta = O.random_fr(rng, 1)[0]
    npow = max(n, 64)
    # powers beta^i * G through the oracle (scalar multiples of the generator)
    pw, acc = [], 1
    for _ in range(npow):
        pw.append(acc)
        acc = acc * beta % O.R_MOD
    g_pts = [O.g1_mul(O.G1_GEN, e) for e in pw] if npow <= 200 else None     # long sets: synthetic bases + C oracle
    p = O.random_fr(rng, n)
    z = O.random_fr(rng, 1)[0]
    R = O.R_MOD
    ev = sum(c * pow(z, i, R) for i, c in enumerate(p)) % R
    q, carry = [0] * max(n - 1, 0), 0
    for j in range(n - 1, -1, -1):
        carry = (p[j] + z * carry) % R
        if j >= 1:
            q[j - 1] = carry
    if g_pts is None:
        bases = _synthetic(npow, 77)
        hb = bases.cpu().numpy()
        powers = S.Powers(bases)
        want = oracle_msm(hb[:(n - 1) * 104], H.scalars_array(q)) if n > 1 else None
    else:
        hb = H.bases_array(g_pts)
        powers = S.Powers(hb)
        want = O.g1_mul(O.G1_GEN, sum(c * pw[i] for i, c in enumerate(q)) % R) if n > 1 else None
        # the pairing-free form of the KZG check: (p(beta) - p(z)) * G == (beta - z) * W
        pb = sum(c * pw[i] for i, c in enumerate(p)) % R
        assert O.g1_mul(O.G1_GEN, (pb - ev) % R) == O.g1_mul(want, (beta - z) % R)
    w, v, rv = S.KZG10.open(powers, H.fr_mont_array(p), H.fr_mont_array([z])[0])
    assert rv is None
    assert H.fr_from_mont_array(v.reshape(1, 4)) == [ev]
    assert H.jac_bytes_to_affine(w) == want
    powers.release()
