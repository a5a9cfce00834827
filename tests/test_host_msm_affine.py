"""CPU-side verification of the DEVICE batched-affine bucket rounds (snarkos_b200/csrc/msm_affine.cuh): the
per-thread functions the CUDA kernels execute (denominators -> batch inversion -> pairwise affine additions) are
run thread by thread on the host (tests/host/msm_host.cpp, PTX emulated bit-exactly) and the resulting MSM is
compared with the oracle -- the counterpart of snarkVM's batched::batch_add [UPSTREAM
algorithms/src/msm/variable_base/batched.rs] which the reference tests against the naive sum
(variable_base::tests::test_msm)."""
import ctypes
import os
import subprocess

import numpy as np
import pytest

from oracle import bls12_377 as O
from oracle import c_oracle as C
from tests import helpers as H

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)


@pytest.fixture(scope="module")
def hostlib():
    src = os.path.join(HERE, "host", "msm_host.cpp")
    so = os.path.join(HERE, "host", "libmsm_host.so")
    deps = [src] + [os.path.join(ROOT, "snarkos_b200", "csrc", f)
                    for f in ("field.cuh", "ec.cuh", "msm_core.cuh", "msm_affine.cuh", "msm_glv.cuh", "ptx_ops.cuh")]
    if not os.path.exists(so) or any(os.path.getmtime(d) > os.path.getmtime(so) for d in deps):
        subprocess.run(["g++", "-O2", "-std=c++17", "-shared", "-fPIC", "-x", "c++", src, "-o", so], check=True)
    return ctypes.CDLL(so)


def run(lib, pts, scalars, c, rounds):
    bases = H.bases_array(pts)
    sc = H.scalars_array(scalars)
    out = np.zeros(144, dtype=np.uint8)
    stats = np.zeros(max(rounds, 1), dtype=np.uint32)
    rc = lib.host_msm_affine(out.ctypes.data_as(ctypes.c_void_p), bases.ctypes.data_as(ctypes.c_void_p),
                             ctypes.c_size_t(len(pts)), ctypes.c_size_t(104), sc.ctypes.data_as(ctypes.c_void_p),
                             ctypes.c_uint32(c), ctypes.c_uint32(rounds), stats.ctypes.data_as(ctypes.c_void_p))
    assert rc == 0
    return H.jac_bytes_to_affine(out), stats


@pytest.mark.parametrize("n,c,rounds", [(1, 4, 2), (7, 3, 1), (40, 4, 3), (200, 4, 4), (150, 5, 8), (64, 2, 10)])
def test_rounds_match_oracle(hostlib, n, c, rounds):
    rng = O.SplitMix64(900 + n + rounds)
    pts = O.random_points(rng, n)
    sc = O.random_fr(rng, n)
    want = O.msm_naive(pts, sc) if n <= 40 else H.jac_bytes_to_affine(C.msm(H.bases_array(pts), H.scalars_array(sc)))
    got, stats = run(hostlib, pts, sc, c, rounds)
    assert got == want
    # the lists shrink round by round down to one point per non-empty bucket
    assert all(stats[i + 1] <= stats[i] for i in range(rounds - 1))
    assert run(hostlib, pts, sc, c, 0)[0] == want          # rounds = 0: the plain XYZZ path of the same harness


def test_exceptional_pairs(hostlib):
    """P + P (doubling), P + (-P) (infinity), infinity operands, repeated points: every branch of pair_classify,
    in pairs (adjacent entries of one bucket) and across rounds."""
    rng = O.SplitMix64(77)
    pts = O.random_points(rng, 48)
    sc = O.random_fr(rng, 48)
    # identical (point, scalar) runs: whole buckets of equal points -> doublings in every round
    for i in range(1, 8):
        pts[i], sc[i] = pts[0], sc[0]
    # opposite points with the same scalar: P + (-P) inside a bucket
    pts[9], sc[9] = O.g1_neg(pts[8]), sc[8]
    # same point, opposite scalars: the sign bit makes the pair cancel
    pts[11], sc[11] = pts[10], O.R_MOD - sc[10]
    # points at infinity in the bases
    pts[12] = None
    pts[13] = None
    sc[13] = sc[12]
    sc[14], sc[15], sc[16] = 0, 1, O.R_MOD - 1
    want = O.msm_naive(pts, sc)
    for c, rounds in ((3, 1), (3, 4), (4, 7), (6, 2)):
        assert run(hostlib, pts, sc, c, rounds)[0] == want
    # everything equal: one bucket per window holding all n points
    same = [pts[20]] * 33
    s = [sc[20]] * 33
    assert run(hostlib, same, s, 4, 6)[0] == O.g1_mul(pts[20], 33 * sc[20] % O.R_MOD)
    # total exactly infinity
    assert run(hostlib, [pts[0], pts[0]], [5, O.R_MOD - 5], 4, 2)[0] is None
    assert run(hostlib, [None] * 5, sc[:5], 4, 2)[0] is None


def test_glv_split_and_endomorphism(hostlib):
    """k = k1 + k2 * lambda with k1, k2 < 2^127 (device code of msm_glv.cuh against Python divmod), and
    (beta x, y) = lambda * (x, y) on the oracle's curve arithmetic"""
    u = 0x8508C00000000001
    lam = u * u - 1
    assert (lam * lam + lam + 1) % O.R_MOD == 0
    rng = O.SplitMix64(4242)
    ks = [0, 1, lam - 1, lam, lam + 1, 2 * lam - 1, 2 * lam, O.R_MOD - 1, O.R_MOD - 2, (O.R_MOD // lam) * lam, (O.R_MOD // lam) * lam - 1]
    ks += [rng.below(O.R_MOD, 253) for _ in range(20000)]
    K = H.ints_to_limbs(ks, 4)
    k1 = np.zeros((len(ks), 2), dtype=np.uint64)
    k2 = np.zeros((len(ks), 2), dtype=np.uint64)
    hostlib.host_glv_split(k1.ctypes.data_as(ctypes.c_void_p), k2.ctypes.data_as(ctypes.c_void_p),
                           K.ctypes.data_as(ctypes.c_void_p), ctypes.c_size_t(len(ks)))
    a, b = H.limbs_to_ints(k1), H.limbs_to_ints(k2)
    for k, x, y in zip(ks, a, b):
        assert (x, y) == (k % lam, k // lam) and x < (1 << 127) and y < (1 << 127)
    pts = O.random_points(rng, 4)
    xs = H.ints_to_limbs([p[0] * (1 << 384) % O.P_MOD for p in pts], 6)
    out = np.zeros_like(xs)
    hostlib.host_glv_endo_x(out.ctypes.data_as(ctypes.c_void_p), xs.ctypes.data_as(ctypes.c_void_p), ctypes.c_size_t(len(pts)))
    rinv = pow(1 << 384, -1, O.P_MOD)
    for p, bx in zip(pts, H.limbs_to_ints(out)):
        assert (bx * rinv % O.P_MOD, p[1]) == O.g1_mul(p, lam)
