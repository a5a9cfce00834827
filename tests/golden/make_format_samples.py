"""Writes the oraclefmt_*.bin samples: the snarkVM fixture format (FIXTURES.md) filled by the PYTHON big-int oracle
(affine chord-and-tangent MSM, textbook NTT) -- an implementation independent of both the C oracle and the CUDA path.
They are NOT reference outputs; they exist so that the fixture reader and its two consumers run in every test session.

    python tests/golden/make_format_samples.py
"""
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))

from oracle import bls12_377 as O                      # noqa: E402
from tests.golden import fixture_format as F          # noqa: E402


def jacobian_bytes(pt):
    """a (non-normalised) Jacobian image of an affine point: (x z^2, y z^3, z) with z = 3, or (1, 1, 0) for infinity"""
    if pt is None:
        one = O.fq_to_mont(1).to_bytes(48, "little")
        return one + one + bytes(48)
    z = 3
    X, Y = pt[0] * z * z % O.P_MOD, pt[1] * z * z * z % O.P_MOD
    return b"".join(O.fq_to_mont(v).to_bytes(48, "little") for v in (X, Y, z))


def msm_sample(tag, pts, scalars):
    res = O.msm_naive(pts, scalars)
    F.write_msm(os.path.join(HERE, f"oraclefmt_msm_{tag}.bin"),
                b"".join(O.affine_bytes(p) for p in pts), b"".join(O.scalar_bytes(s) for s in scalars), len(pts),
                jacobian_bytes(res), O.affine_bytes(res), O.g1_compress(res))


def ntt_sample(tag, log_n, vals):
    d = O.EvaluationDomain(1 << log_n)
    enc = lambda xs: b"".join(O.fr_bytes_mont(v) for v in xs)
    F.write_ntt(os.path.join(HERE, f"oraclefmt_ntt_{tag}.bin"), log_n, enc(vals),
                [enc(d.fft(vals)), enc(d.ifft(vals)), enc(d.coset_fft(vals)), enc(d.coset_ifft(vals))])


def main():
    rng = O.SplitMix64(20261019)
    for n in (1, 14, 15, 33, 200):
        msm_sample(f"n{n}", O.random_points(rng, n), O.random_fr(rng, n))
    pts, sc = O.random_points(rng, 24), O.random_fr(rng, 24)
    G = O.G1_GEN
    pts[3] = None
    pts[10], pts[11], pts[12] = G, G, O.g1_neg(G)
    sc[5], sc[6], sc[7] = 0, 1, O.R_MOD - 1
    sc[10] = sc[11] = sc[12] = sc[13]
    msm_sample("edge", pts, sc)
    msm_sample("cancel", [G, O.g1_neg(G)], [7, 7])                 # result = infinity
    for log_n in (0, 1, 3, 8, 11):
        ntt_sample(f"log{log_n}", log_n, O.random_fr(rng, 1 << log_n))
    ntt_sample("log9_padded", 9, O.random_fr(rng, 300))


if __name__ == "__main__":
    main()
