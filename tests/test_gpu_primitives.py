"""GPU parity (bit-exact) of the device field and curve primitives against the oracle, through the C ABI's
diagnostic entry points (b200_debug_field_op / b200_debug_g1_op)."""
import ctypes

import numpy as np
import pytest

from oracle import bls12_377 as O
from oracle import c_oracle as C
from tests import helpers as H

pytestmark = pytest.mark.gpu


def _p(a):
    return a.ctypes.data_as(ctypes.c_void_p)


def field_op(op, a, b, nl):
    import snarkos_b200 as S
    A = H.ints_to_limbs(a, nl)
    B = H.ints_to_limbs(b, nl) if b is not None else None
    out = np.empty_like(A)
    S._lib.check(S.lib().b200_debug_field_op(op, _p(out), _p(A), _p(B) if B is not None else None, len(a)))
    return H.limbs_to_ints(out)


def _edge(mod):
    return [0, 1, 2, mod - 1, mod - 2, (1 << 32) - 1, 1 << 32, (1 << 64) - 1, mod >> 1, (mod >> 1) + 1]


@pytest.mark.parametrize("name,base,mod,bits,nl,rinv", [("fr", 0, O.R_MOD, 253, 4, O.FR_R_INV), ("fq", 3, O.P_MOD, 377, 6, O.FQ_R_INV)])
def test_field_ops_bit_exact(name, base, mod, bits, nl, rinv):
    rng = O.SplitMix64(77)
    e = _edge(mod)
    a = [rng.below(mod, bits) for _ in range(5000)] + [x for x in e for _ in e]
    b = [rng.below(mod, bits) for _ in range(5000)] + [y for _ in e for y in e]
    assert field_op(base + 0, a, b, nl) == [x * y * rinv % mod for x, y in zip(a, b)]
    assert field_op(base + 1, a, b, nl) == [(x + y) % mod for x, y in zip(a, b)]
    assert field_op(base + 2, a, b, nl) == [(x - y) % mod for x, y in zip(a, b)]
    R = pow(rinv, -1, mod)
    inv = field_op(7 if name == "fr" else 6, a[:64] + e, None, nl)
    for x, xi in zip(a[:64] + e, inv):
        assert xi == (0 if x == 0 else pow(x * rinv % mod, -1, mod) * R % mod)


def test_field_mul_bulk_vs_c_oracle():
    rng = np.random.default_rng(1)
    import snarkos_b200 as S
    for nl, op, f in ((4, 0, C.fr_mul), (6, 3, C.fq_mul)):
        n = 1 << 16
        a = rng.integers(0, 1 << 63, size=(n, nl), dtype=np.uint64)
        b = rng.integers(0, 1 << 63, size=(n, nl), dtype=np.uint64)
        a[:, -1] &= np.uint64((1 << 55) - 1)      # keep below both moduli
        b[:, -1] &= np.uint64((1 << 55) - 1)
        out = np.empty_like(a)
        S._lib.check(S.lib().b200_debug_field_op(op, _p(out), _p(a), _p(b), n))
        assert np.array_equal(out, f(a, b))


def test_g1_ops_bit_exact():
    import snarkos_b200 as S
    rng = O.SplitMix64(9)
    pts = O.random_points(rng, 24)
    a = pts + [pts[0], pts[1], None, pts[2], None]
    b = pts[1:] + [pts[0]] + [pts[0], O.g1_neg(pts[1]), pts[3], None, None]     # P+P, P+(-P), inf+P, P+inf, inf+inf
    A, B = H.bases_array(a), H.bases_array(b)
    out = np.zeros((len(a), 144), dtype=np.uint8)
    S._lib.check(S.lib().b200_debug_g1_op(0, _p(out), _p(A), _p(B), len(a), 104))
    for i in range(len(a)):
        assert H.jac_bytes_to_affine(out[i]) == O.g1_add(a[i], b[i]), i
    S._lib.check(S.lib().b200_debug_g1_op(1, _p(out), _p(A), None, len(a), 104))
    for i in range(len(a)):
        assert H.jac_bytes_to_affine(out[i]) == O.g1_add(a[i], a[i]), i
    ks = [0, 1, 2, 3, (1 << 64) - 1] + [rng.next() for _ in range(len(a) - 5)]
    K = np.array(ks, dtype=np.uint64)
    S._lib.check(S.lib().b200_debug_g1_op(2, _p(out), _p(A), _p(K), len(a), 104))
    for i in range(len(a)):
        assert H.jac_bytes_to_affine(out[i]) == O.g1_mul(a[i], ks[i]), i


def test_synthetic_bases_match_oracle():
    import torch
    import snarkos_b200 as S
    n, seed = 300, 1234567890
    dev = S.synthetic_bases(n, seed=seed)
    torch.cuda.synchronize()
    got = dev.cpu().numpy().reshape(n, 104)
    k = H.splitmix64_at(seed, np.arange(n))
    want = C.g1_mul_u64(np.frombuffer(O.affine_bytes(O.G1_GEN), dtype=np.uint8), k)
    assert np.array_equal(got, want)
    assert O.affine_from_bytes(bytes(got[7])) == O.g1_mul(O.G1_GEN, int(k[7]))
