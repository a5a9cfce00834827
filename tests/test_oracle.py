"""Pins the oracle (CPU, test infrastructure) before anything trusts it:
   Python big-int oracle  <- known-answer vectors tests/golden/kat.json
   C oracle (liboracle.so) <- Python oracle on seeded random inputs.
Mirrors the style of snarkVM's own self-consistency tests (variable_base::tests::test_msm: naive vs
standard; fft domain tests: roots of unity, fft o ifft = id, coset round-trip) -- SURVEY.md 8c."""
import numpy as np
import pytest

from oracle import bls12_377 as O
from oracle import c_oracle as C
from tests import helpers as H

KAT = H.load_kat()


def test_moduli_and_montgomery_constants():
    assert O.R_MOD == KAT["r_hex"] and O.P_MOD == KAT["p_hex"]
    assert H.limbs_to_ints(np.array([KAT["fr_R_limbs"]], dtype=np.uint64))[0] == O.FR_R
    assert H.limbs_to_ints(np.array([KAT["fr_R2_limbs"]], dtype=np.uint64))[0] == O.FR_R * O.FR_R % O.R_MOD
    assert H.limbs_to_ints(np.array([KAT["fq_R_limbs"]], dtype=np.uint64))[0] == O.FQ_R
    assert H.limbs_to_ints(np.array([KAT["fq_R2_limbs_hex"]], dtype=np.uint64))[0] == O.FQ_R * O.FQ_R % O.P_MOD
    assert H.limbs_to_ints(np.array([KAT["fr_gen22_mont_limbs"]], dtype=np.uint64))[0] == O.fr_to_mont(22)
    assert O.FR_TWO_ADIC_ROOT == KAT["fr_two_adic_root"]
    assert H.limbs_to_ints(np.array([KAT["fr_two_adic_root_mont_limbs"]], dtype=np.uint64))[0] == O.fr_to_mont(O.FR_TWO_ADIC_ROOT)
    assert (-pow(O.R_MOD, -1, 1 << 32)) % (1 << 32) == 0xFFFFFFFF
    assert (-pow(O.P_MOD, -1, 1 << 32)) % (1 << 32) == 0xFFFFFFFF


def test_roots_of_unity_kat():
    assert O.EvaluationDomain(4).group_gen == KAT["omega_4"]
    assert O.EvaluationDomain(8).group_gen == KAT["omega_8"]
    assert O.EvaluationDomain(1 << 20).group_gen == KAT["omega_2^20"]
    assert O.EvaluationDomain(1 << 20).size_inv == KAT["(2^20)^-1"]
    assert O.EvaluationDomain(4).generator_inv == KAT["22^-1"]
    w = O.FR_TWO_ADIC_ROOT
    assert pow(w, 1 << 47, O.R_MOD) == 1 and pow(w, 1 << 46, O.R_MOD) != 1


def test_g1_kat():
    G = tuple(KAT["G"])
    assert G == O.G1_GEN and O.is_on_curve(G)
    assert O.g1_add(G, G) == tuple(KAT["2G"])
    assert O.g1_add(O.g1_add(G, G), G) == tuple(KAT["3G"])
    assert O.g1_mul(G, 3) == tuple(KAT["3G"])
    assert O.g1_mul(G, O.R_MOD) is None
    assert O.g1_mul(G, O.R_MOD - 1) == (G[0], O.P_MOD - G[1])
    got = O.msm_naive([G, tuple(KAT["2G"]), tuple(KAT["3G"])], [1, 2, 3])
    assert got == tuple(KAT["MSM_1_2_3__G_2G_3G"]) == O.g1_mul(G, 14)


def test_ntt_kat():
    assert O.EvaluationDomain(4).fft([1, 2, 3, 4]) == KAT["NTT_4_1234"]
    assert O.EvaluationDomain(4).coset_fft([1, 2, 3, 4]) == KAT["cosetNTT_4_1234"]
    assert O.EvaluationDomain(8).fft(list(range(1, 9))) == KAT["NTT_8_1to8"]


@pytest.mark.parametrize("n", [1, 2, 4, 16, 64, 256])
def test_ntt_vs_naive_dft_and_roundtrips(n):
    rng = O.SplitMix64(1000 + n)
    x = O.random_fr(rng, n)
    d = O.EvaluationDomain(n)
    y = d.fft(x)
    assert y == d.dft_naive(x)
    assert d.ifft(y) == x
    assert d.coset_ifft(d.coset_fft(x)) == x
    # delta -> all ones, ones -> n * delta, zero padding
    assert d.fft([1]) == [1] * n
    assert d.fft([1] * n) == [n % O.R_MOD] + [0] * (n - 1)


def test_pippenger_vs_naive_python():
    rng = O.SplitMix64(7)
    for n in (1, 5, 33, 100):
        pts = O.random_points(rng, n)
        sc = O.random_fr(rng, n)
        assert O.msm_pippenger(pts, sc) == O.msm_naive(pts, sc)


# ---------------------------------------------------------------- C oracle vs Python oracle
def test_c_field_ops_vs_python():
    rng = O.SplitMix64(42)
    n = 200
    a = [rng.below(O.R_MOD, 253) for _ in range(n)] + [0, 1, O.R_MOD - 1]
    b = [rng.below(O.R_MOD, 253) for _ in range(n)] + [O.R_MOD - 1, O.R_MOD - 1, O.R_MOD - 1]
    A, B = H.ints_to_limbs(a, 4), H.ints_to_limbs(b, 4)
    assert H.limbs_to_ints(C.fr_mul(A, B)) == [O.mont_mul(x, y, O.R_MOD, O.FR_R_INV) for x, y in zip(a, b)]
    assert H.limbs_to_ints(C.fr_add(A, B)) == [(x + y) % O.R_MOD for x, y in zip(a, b)]
    assert H.limbs_to_ints(C.fr_sub(A, B)) == [(x - y) % O.R_MOD for x, y in zip(a, b)]
    assert H.limbs_to_ints(C.fr_to_mont(A)) == [O.fr_to_mont(x) for x in a]
    assert H.limbs_to_ints(C.fr_from_mont(A)) == [O.fr_from_mont(x) for x in a]
    a = [rng.below(O.P_MOD, 377) for _ in range(n)] + [0, 1, O.P_MOD - 1]
    b = [rng.below(O.P_MOD, 377) for _ in range(n)] + [O.P_MOD - 1, O.P_MOD - 1, O.P_MOD - 1]
    A, B = H.ints_to_limbs(a, 6), H.ints_to_limbs(b, 6)
    assert H.limbs_to_ints(C.fq_mul(A, B)) == [O.mont_mul(x, y, O.P_MOD, O.FQ_R_INV) for x, y in zip(a, b)]
    assert H.limbs_to_ints(C.fq_add(A, B)) == [(x + y) % O.P_MOD for x, y in zip(a, b)]
    assert H.limbs_to_ints(C.fq_sub(A, B)) == [(x - y) % O.P_MOD for x, y in zip(a, b)]
    inv = H.limbs_to_ints(C.fq_inv(A[:20]))
    for x, xi in zip(a[:20], inv):  # Montgomery inverse: (x/R)^-1 * R
        assert O.fq_from_mont(xi) * O.fq_from_mont(x) % O.P_MOD == 1


@pytest.mark.parametrize("n", [0, 1, 3, 14, 15, 33, 100, 1000])
def test_c_msm_vs_python(n):
    rng = O.SplitMix64(99 + n)
    pts = O.random_points(rng, n) if n else []
    sc = O.random_fr(rng, n)
    if n >= 3:   # edge cases of SURVEY 7: zero scalar, r-1, point at infinity, duplicates
        sc[0] = 0
        sc[1] = O.R_MOD - 1
        pts[2] = None
    if n >= 33:
        pts[5] = pts[4]
        sc[5] = sc[4]
        pts[7] = O.g1_neg(pts[6])
        sc[7] = sc[6]
    want = O.msm_pippenger(pts, sc) if n > 100 else O.msm_naive(pts, sc)
    got = C.msm(H.bases_array(pts), H.scalars_array(sc))
    assert H.jac_bytes_to_affine(got) == want


def test_c_msm_kat():
    G = tuple(KAT["G"])
    got = C.msm(H.bases_array([G, tuple(KAT["2G"]), tuple(KAT["3G"])]), H.scalars_array([1, 2, 3]))
    assert H.jac_bytes_to_affine(got) == tuple(KAT["MSM_1_2_3__G_2G_3G"])


def test_c_g1_mul_u64():
    rng = O.SplitMix64(5)
    ks = [0, 1, 2, (1 << 64) - 1] + [rng.next() for _ in range(6)]
    out = C.g1_mul_u64(np.frombuffer(O.affine_bytes(O.G1_GEN), dtype=np.uint8), np.array(ks, dtype=np.uint64))
    for k, row in zip(ks, out):
        assert O.affine_from_bytes(bytes(row)) == O.g1_mul(O.G1_GEN, k)
        assert C.g1_is_on_curve(row)


@pytest.mark.parametrize("log_n", [0, 1, 2, 3, 5, 8, 10])
@pytest.mark.parametrize("direction,coset", [(0, 0), (1, 0), (0, 1), (1, 1)])
def test_c_ntt_vs_python(log_n, direction, coset):
    n = 1 << log_n
    rng = O.SplitMix64(31 * log_n + 2 * direction + coset)
    batch = 3
    polys = [O.random_fr(rng, n) for _ in range(batch)]
    d = O.EvaluationDomain(n)
    f = {(0, 0): d.fft, (1, 0): d.ifft, (0, 1): d.coset_fft, (1, 1): d.coset_ifft}[(direction, coset)]
    want = [f(p) for p in polys]
    data = H.fr_mont_array([v for p in polys for v in p])
    got = C.ntt(data, log_n, batch=batch, direction=direction, coset=coset)
    assert H.fr_from_mont_array(got) == [v for p in want for v in p]


def test_c_ntt_kat_and_inner_parallel():
    got = C.ntt(H.fr_mont_array([1, 2, 3, 4]), 2)
    assert H.fr_from_mont_array(got) == KAT["NTT_4_1234"]
    got = C.ntt(H.fr_mont_array([1, 2, 3, 4]), 2, coset=1)
    assert H.fr_from_mont_array(got) == KAT["cosetNTT_4_1234"]
    got = C.ntt(H.fr_mont_array(list(range(1, 9))), 3)
    assert H.fr_from_mont_array(got) == KAT["NTT_8_1to8"]
    # batch = 1 takes the inside-one-polynomial parallel path; must agree with the serial one
    rng = O.SplitMix64(77)
    x = H.fr_mont_array(O.random_fr(rng, 1 << 12))
    a = C.ntt(x, 12, nthreads=1)
    b = C.ntt(x, 12, nthreads=4)
    assert np.array_equal(a, b)
    assert np.array_equal(C.ntt(a, 12, direction=1), x)


# ---------------------------------------------------------------------------------------------
# batched::msm restated (the variant snarkVM dispatches BLS12-377 G1 to) against the other two formulations
# ---------------------------------------------------------------------------------------------
@pytest.mark.parametrize("n", [1, 14, 15, 31, 32, 100, 1000, 5000])
def test_batched_msm_equals_standard_and_naive(n):
    """mirrors snarkVM variable_base::tests::test_msm (naive vs standard vs batched, sizes 1..1000)"""
    rng = np.random.default_rng(n)
    pts = C.g1_sequence(777 + n, n)
    sc = H.random_scalars_np(rng, n)
    if n > 40:
        sc[0] = 0
        sc[1] = H.scalars_array([O.R_MOD - 1])[0]
        pts[3] = 0
        pts[3, 96] = 1                                   # infinity inside the bases
        pts[7], sc[7] = pts[6], sc[6]                    # equal points in one bucket: the pair is a doubling
    std = C.g1_to_affine(C.msm(pts.reshape(-1), sc))
    bat = C.g1_to_affine(C.msm_batched(pts.reshape(-1), sc))
    assert np.array_equal(std, bat)
    if n <= 100:
        ks = [777 + n + i for i in range(n)]
        svals = H.limbs_to_ints(sc)
        total = 0
        for i in range(n):
            if pts[i, 96] == 0:
                k = ks[i] if not (n > 40 and i == 7) else ks[6]
                total += svals[i] * k
        want = O.g1_mul(O.G1_GEN, total % O.R_MOD)
        assert H.jac_bytes_to_affine(C.msm_batched(pts.reshape(-1), sc)) == want
    # degenerate distributions: every point in one bucket / every pair a doubling
    same_s = np.tile(sc[n // 2:n // 2 + 1], (n, 1))
    assert np.array_equal(C.g1_to_affine(C.msm(pts.reshape(-1), same_s)), C.g1_to_affine(C.msm_batched(pts.reshape(-1), same_s)))
    same_p = np.tile(pts[n // 2:n // 2 + 1], (n, 1))
    assert np.array_equal(C.g1_to_affine(C.msm(same_p.reshape(-1), sc)), C.g1_to_affine(C.msm_batched(same_p.reshape(-1), sc)))


def test_point_sequence_and_msm_many():
    g = np.frombuffer(O.affine_bytes(O.G1_GEN), dtype=np.uint8)
    k0 = 99990
    assert np.array_equal(C.g1_sequence(k0, 5000), C.g1_mul_u64(g, np.arange(k0, k0 + 5000, dtype=np.uint64)))
    assert C.g1_sequence(0, 3)[0, 96] == 1               # 0 * G = infinity
    rng = np.random.default_rng(3)
    pts = C.g1_sequence(5, 300).reshape(-1)
    sc = H.random_scalars_np(rng, 300)
    off = np.array([0, 40, 40, 100, 300], dtype=np.uint64)
    out = C.msm_many(pts, sc, off)
    for m in range(4):
        lo, hi = int(off[m]), int(off[m + 1])
        assert np.array_equal(C.g1_to_affine(out[m]), C.g1_to_affine(C.msm(pts[lo * 104:hi * 104], sc[lo:hi])))


def test_ntt_precomputation_is_cached_and_consistent():
    """the oracle keeps snarkVM's FFTPrecomputation per domain; repeated calls, both directions and the coset variants
    still agree with the big-int DFT"""
    rng = O.SplitMix64(17)
    x = O.random_fr(rng, 64)
    d = O.EvaluationDomain(64)
    a = H.fr_mont_array(x)
    for _ in range(2):
        assert H.fr_from_mont_array(C.ntt(a, 6)) == d.fft(x)
        assert H.fr_from_mont_array(C.ntt(a, 6, direction=1)) == d.ifft(x)
        assert H.fr_from_mont_array(C.ntt(a, 6, coset=1)) == d.coset_fft(x)
        assert H.fr_from_mont_array(C.ntt(a, 6, direction=1, coset=1)) == d.coset_ifft(x)
