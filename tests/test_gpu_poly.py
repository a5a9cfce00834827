"""GPU parity of the device-resident polynomial glue (snarkos_b200/poly.py) against Python big-int arithmetic, and an
end-to-end quotient computation in the style of a Varuna sumcheck round: coset FFT -> pointwise -> divide by the
vanishing polynomial -> coset iFFT, all on the device, checked against the oracle's polynomial arithmetic."""
import numpy as np
import pytest

from oracle import bls12_377 as O
from tests import helpers as H

pytestmark = pytest.mark.gpu


def dev(vals):
    import torch
    return torch.from_numpy(H.fr_mont_array(vals).view(np.int64)).cuda()


def host(t):
    import torch
    torch.cuda.synchronize()
    return H.fr_from_mont_array(t.cpu().numpy().view(np.uint64).reshape(-1, 4))


def test_elementwise_ops():
    from snarkos_b200 import poly as P
    rng = O.SplitMix64(1)
    n = 1000
    a, b, c = (O.random_fr(rng, n) for _ in range(3))
    a[0], b[1], a[2], b[2] = 0, 0, O.R_MOD - 1, O.R_MOD - 1
    R = O.R_MOD
    assert host(P.mul(dev(a), dev(b))) == [x * y % R for x, y in zip(a, b)]
    assert host(P.add(dev(a), dev(b))) == [(x + y) % R for x, y in zip(a, b)]
    assert host(P.sub(dev(a), dev(b))) == [(x - y) % R for x, y in zip(a, b)]
    assert host(P.mul_add(dev(a), dev(b), dev(c))) == [(x * y + z) % R for x, y, z in zip(a, b, c)]
    assert host(P.mul_sub(dev(a), dev(b), dev(c))) == [(x * y - z) % R for x, y, z in zip(a, b, c)]
    assert host(P.mul(dev(a), dev([12345]))) == [x * 12345 % R for x in a]                # broadcast scalar
    out = dev([0] * n)
    da = dev(a)
    P.mul(da, dev(b), out=out)
    assert host(out) == [x * y % R for x, y in zip(a, b)] and host(da) == a               # out-of-place leaves a


@pytest.mark.parametrize("n", [1, 5, 127, 2048, 2049, 70000, (1 << 20) + 3])
def test_batch_inversion(n):
    from snarkos_b200 import poly as P
    rng = np.random.default_rng(n)
    x = H.random_fr_mont_np(rng, (n,))
    x[::97] = 0                                     # zeros stay zero
    import torch
    t = torch.from_numpy(x.view(np.int64)).cuda()
    P.batch_inversion(t)
    from snarkos_b200 import poly
    prod = poly.mul(t.clone(), torch.from_numpy(x.view(np.int64)).cuda())           # x * x^-1 = 1 (Montgomery one) or 0
    torch.cuda.synchronize()
    got = prod.cpu().numpy().view(np.uint64)
    one = H.fr_mont_array([1])[0]
    zero_rows = np.all(x == 0, axis=1)
    assert np.all(got[~zero_rows] == one) and np.all(got[zero_rows] == 0)
    assert np.all(t.cpu().numpy().view(np.uint64)[zero_rows] == 0)
    # spot check against Python
    idx = [i for i in (1, n // 2, n - 1) if 0 <= i < n and not zero_rows[i]]
    for i in idx:
        v = H.fr_from_mont_array(x[i:i + 1])[0]
        assert H.fr_from_mont_array(t[i:i + 1].cpu().numpy().view(np.uint64))[0] == pow(v, -1, O.R_MOD)


@pytest.mark.parametrize("log_h,log_k", [(3, 3), (4, 6), (10, 12)])
def test_divide_by_vanishing_and_quotient_pipeline(log_h, log_k):
    """h(X) = (a(X) * b(X) - c(X)) / v_H(X) where c = a*b mod v_H is chosen so that the division is exact:
    computed on the device via coset FFTs over K and compared with big-int polynomial arithmetic."""
    import snarkos_b200 as S
    from snarkos_b200 import poly as P
    nh, nk = 1 << log_h, 1 << log_k
    rng = O.SplitMix64(log_k)
    # pick the quotient q and a remainder-free product: t(X) = q(X) * (X^nh - 1)
    deg_q = nk - nh - 1
    q = O.random_fr(rng, deg_q + 1) if deg_q >= 0 else []
    t = [0] * nk
    for i, v in enumerate(q):
        t[i + nh] = (t[i + nh] + v) % O.R_MOD
        t[i] = (t[i] - v) % O.R_MOD
    dk = S.EvaluationDomain(nk)
    ev = dev(t)
    dk.coset_fft_in_place(ev)
    # sanity of the periodic table: compare a few quotient evaluations with Python
    P.divide_by_vanishing_poly_on_coset_in_place(ev, log_k, log_h)
    w = O.EvaluationDomain(nk).group_gen
    got = host(ev)
    for i in (0, 1, nk // 2 + 1, nk - 1):
        x = 22 * pow(w, i, O.R_MOD) % O.R_MOD
        want = sum(c * pow(x, e, O.R_MOD) for e, c in enumerate(q)) % O.R_MOD
        assert got[i] == want
    dk.coset_ifft_in_place(ev)
    assert host(ev) == (q + [0] * (nk - len(q)))


@pytest.mark.parametrize("n", [1, 2, 63, 64, 65, 1000, 4096, 4097, 70000])
def test_divide_by_linear(n):
    """quotient and remainder of p by (X - z) == synthetic division with Python big-ints (chunk boundaries at 64 and
    64^2 = 4096 are the recursion levels of the device algorithm)"""
    from snarkos_b200 import poly as P
    rng = O.SplitMix64(n)
    p = O.random_fr(rng, n)
    z = O.random_fr(rng, 1)[0]
    R = O.R_MOD
    q, carry = [0] * max(n - 1, 0), 0
    for j in range(n - 1, -1, -1):
        carry = (p[j] + z * carry) % R
        if j >= 1:
            q[j - 1] = carry
    got_q, got_r = P.divide_by_linear(dev(p), dev([z]))
    assert host(got_r) == [carry]
    assert host(got_q) == q if n > 1 else got_q.numel() == 0
    # degenerate points
    for zz in (0, 1, R - 1):
        gq, gr = P.divide_by_linear(dev(p), dev([zz]))
        assert host(gr) == [sum(c * pow(zz, i, R) for i, c in enumerate(p)) % R]


def test_linear_combination():
    from snarkos_b200 import poly as P
    rng = O.SplitMix64(5)
    lens = [100, 1, 0, 257, 64]
    polys = [O.random_fr(rng, L) for L in lens]
    cs = O.random_fr(rng, len(lens))
    import torch
    dp = [dev(p) if p else torch.empty((0, 4), dtype=torch.int64, device="cuda") for p in polys]
    got = host(P.linear_combination(dp, dev(cs)))
    want = [sum(c * p[i] for c, p in zip(cs, polys) if i < len(p)) % O.R_MOD for i in range(max(lens))]
    assert got == want
