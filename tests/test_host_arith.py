"""CPU-side verification of the DEVICE arithmetic sources (snarkos_b200/csrc/{field,ec}.cuh).

The headers are compiled with plain g++ through tests/host/field_host.cpp, where every PTX carry-chain
instruction is emulated bit-exactly (csrc/ptx_ops.cuh), so the exact instruction sequences that run on
the B200 are checked against the oracle here, without a GPU.  This is a test shim, not a CPU fallback:
the shipped library contains device code only."""
import ctypes
import os
import subprocess

import numpy as np
import pytest

from oracle import bls12_377 as O
from tests import helpers as H

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)


@pytest.fixture(scope="module")
def hostlib():
    src = os.path.join(HERE, "host", "field_host.cpp")
    so = os.path.join(HERE, "host", "libfield_host.so")
    deps = [src] + [os.path.join(ROOT, "snarkos_b200", "csrc", f) for f in ("field.cuh", "ec.cuh", "ptx_ops.cuh")]
    if not os.path.exists(so) or any(os.path.getmtime(d) > os.path.getmtime(so) for d in deps):
        subprocess.run(["g++", "-O2", "-std=c++17", "-shared", "-fPIC", "-x", "c++", src, "-o", so], check=True)
    return ctypes.CDLL(so)


def _p(a):
    return a.ctypes.data_as(ctypes.c_void_p)


def _bin(lib, name, a, b, nl):
    A, B = H.ints_to_limbs(a, nl), H.ints_to_limbs(b, nl)
    out = np.empty_like(A)
    getattr(lib, name)(_p(out), _p(A), _p(B), ctypes.c_size_t(len(a)))
    return H.limbs_to_ints(out)


def _un(lib, name, a, nl):
    A = H.ints_to_limbs(a, nl)
    out = np.empty_like(A)
    getattr(lib, name)(_p(out), _p(A), ctypes.c_size_t(len(a)))
    return H.limbs_to_ints(out)


def _edge(mod):
    return [0, 1, 2, mod - 1, mod - 2, (1 << 32) - 1, 1 << 32, (1 << 64) - 1, mod >> 1, (mod >> 1) + 1]


@pytest.mark.parametrize("name,mod,bits,nl,rinv", [("fr", O.R_MOD, 253, 4, O.FR_R_INV), ("fq", O.P_MOD, 377, 6, O.FQ_R_INV)])
def test_field_ops(hostlib, name, mod, bits, nl, rinv):
    rng = O.SplitMix64(2024)
    e = _edge(mod)
    a = [rng.below(mod, bits) for _ in range(3000)] + [x for x in e for _ in e]
    b = [rng.below(mod, bits) for _ in range(3000)] + [y for _ in e for y in e]
    want = [x * y * rinv % mod for x, y in zip(a, b)]
    assert _bin(hostlib, f"host_{name}_mul", a, b, nl) == want
    assert _bin(hostlib, f"host_{name}_mul_cc", a, b, nl) == want          # carry-chain variant
    assert _un(hostlib, f"host_{name}_sqr", a, nl) == [x * x * rinv % mod for x in a]         # dedicated squaring
    assert _bin(hostlib, f"host_{name}_mul_wide", a, b, nl) == want                           # wide product + stand-alone REDC
    assert _bin(hostlib, f"host_{name}_mms0", a, b, nl) == [0] * len(a)                       # a*b - b*a, one reduction
    assert _bin(hostlib, f"host_{name}_mms2", a, b, nl) == [(x * x - y * y) * rinv % mod for x, y in zip(a, b)]
    assert _bin(hostlib, f"host_{name}_add", a, b, nl) == [(x + y) % mod for x, y in zip(a, b)]
    assert _bin(hostlib, f"host_{name}_sub", a, b, nl) == [(x - y) % mod for x, y in zip(a, b)]
    assert _un(hostlib, f"host_{name}_neg", a, nl) == [(-x) % mod for x in a]
    inv = _un(hostlib, f"host_{name}_inv", a[:40] + e, nl)
    R = pow(rinv, -1, mod)
    for x, xi in zip(a[:40] + e, inv):
        want = 0 if x == 0 else pow(x * rinv % mod, -1, mod) * R % mod
        assert xi == want
    # binary extended Euclid (the batch-inversion tail of the MSM pair rounds): same Montgomery inverse, 0 -> 0
    xs = a[:600] + e
    inv = _un(hostlib, f"host_{name}_inv_gcd", xs, nl)
    assert inv == [0 if x == 0 else pow(x * rinv % mod, -1, mod) * R % mod for x in xs]


def test_fr_mont_conversions(hostlib):
    rng = O.SplitMix64(11)
    a = [rng.below(O.R_MOD, 253) for _ in range(100)] + _edge(O.R_MOD)
    assert _un(hostlib, "host_fr_to_mont", a, 4) == [O.fr_to_mont(x) for x in a]
    assert _un(hostlib, "host_fr_from_mont", a, 4) == [O.fr_from_mont(x) for x in a]


def _jac(lib_call):
    out = np.zeros(144, dtype=np.uint8)
    lib_call(out)
    return H.jac_bytes_to_affine(out)


def test_g1_madd_chain_with_edge_cases(hostlib):
    rng = O.SplitMix64(5)
    pts = O.random_points(rng, 40)
    # exceptional cases in sequence: infinity operands, P + P, P + (-P), start from infinity
    seq = [None, pts[0], pts[0], pts[1], O.g1_neg(pts[1]), None, pts[2], pts[2], pts[2]] + pts[3:]
    neg = [0] * len(seq)
    neg[10] = 1
    neg[12] = 1
    want = None
    for p, s in zip(seq, neg):
        want = O.g1_add(want, O.g1_neg(p) if s else p)
    bases = H.bases_array(seq)
    negs = np.array(neg, dtype=np.uint8)
    got = _jac(lambda out: hostlib.host_g1_madd_chain(_p(out), _p(bases), _p(negs), ctypes.c_size_t(len(seq)), ctypes.c_size_t(104)))
    assert got == want
    # sum to infinity exactly
    seq = [pts[0], pts[1], O.g1_neg(O.g1_add(pts[0], pts[1]))]
    bases = H.bases_array(seq)
    got = _jac(lambda out: hostlib.host_g1_madd_chain(_p(out), _p(bases), None, ctypes.c_size_t(3), ctypes.c_size_t(104)))
    assert got is None


def test_g1_add_tree_dbl_mul(hostlib):
    rng = O.SplitMix64(6)
    pts = O.random_points(rng, 33)
    seq = pts + [pts[0], None, O.g1_neg(pts[5])]
    want = None
    for p in seq:
        want = O.g1_add(want, p)
    bases = H.bases_array(seq)
    got = _jac(lambda out: hostlib.host_g1_add_tree(_p(out), _p(bases), ctypes.c_size_t(len(seq)), ctypes.c_size_t(104)))
    assert got == want
    # equal leaves force the add -> double branch; opposite leaves force infinity
    for seq, want in (([pts[0], pts[0]], O.g1_add(pts[0], pts[0])), ([pts[0], O.g1_neg(pts[0])], None)):
        bases = H.bases_array(seq)
        got = _jac(lambda out: hostlib.host_g1_add_tree(_p(out), _p(bases), ctypes.c_size_t(2), ctypes.c_size_t(104)))
        assert got == want
    g = H.bases_array([O.G1_GEN])
    assert _jac(lambda out: hostlib.host_g1_dbl(_p(out), _p(g))) == tuple(H.load_kat()["2G"])
    for k in (0, 1, 2, 3, 14, (1 << 64) - 1, rng.next()):
        got = _jac(lambda out: hostlib.host_g1_mul_u64(_p(out), _p(g), ctypes.c_uint64(k)))
        assert got == O.g1_mul(O.G1_GEN, k)
        aff = np.zeros(104, dtype=np.uint8)
        hostlib.host_g1_mul_u64_affine(_p(aff), _p(g), ctypes.c_uint64(k), ctypes.c_size_t(104))
        assert O.affine_from_bytes(bytes(aff)) == O.g1_mul(O.G1_GEN, k)


def test_g1_jacobian_doubling_run(hostlib):
    """g1_dbl_k (dbl-2009-l in Jacobian, a = 0) against the oracle: 2^(k+1) * P, and infinity stays infinity"""
    rng = O.SplitMix64(9)
    pts = O.random_points(rng, 3)
    for p in pts:
        b = H.bases_array([p])
        for k in (0, 1, 2, 11, 18):
            got = _jac(lambda out: hostlib.host_g1_dbl_k(_p(out), _p(b), ctypes.c_uint32(k)))
            assert got == O.g1_mul(p, 1 << (k + 1))
    inf = H.bases_array([None])
    assert _jac(lambda out: hostlib.host_g1_dbl_k(_p(out), _p(inf), ctypes.c_uint32(7))) is None
