"""Device-resident prover slices (snarkos_b200/varuna.py) against the oracle: round-1-style commitment of witness
evaluations (iFFT + KZG commit, nothing leaves HBM in between) and a sumcheck-style quotient through coset FFTs."""
import numpy as np
import pytest

from oracle import bls12_377 as O
from oracle import c_oracle as C
from tests import helpers as H

pytestmark = pytest.mark.gpu


def test_commit_evaluations_matches_oracle():
    import torch
    import snarkos_b200 as S
    log_h, batch = 10, 3
    n = 1 << log_h
    powers_dev = S.synthetic_bases(n, seed=3)
    powers = S.Powers(powers_dev)
    rng = np.random.default_rng(2)
    evals = H.random_fr_mont_np(rng, (batch, n))
    t = torch.from_numpy(evals.view(np.int64)).cuda()
    coeffs, commitments = S.varuna.commit_evaluations(powers, t)
    torch.cuda.synchronize()
    hb = powers_dev.cpu().numpy()
    for b in range(batch):
        want_coeffs = C.ntt(evals[b], log_h, direction=1)
        assert np.array_equal(coeffs[b].cpu().numpy().view(np.uint64), want_coeffs)
        want = H.jac_bytes_to_affine(C.msm(hb, C.fr_from_mont(want_coeffs)))
        assert H.jac_bytes_to_affine(commitments[b].cpu().numpy()) == want
    powers.release()


def test_quotient_on_coset_matches_bigint_polynomials():
    import torch
    import snarkos_b200 as S
    log_h, log_k = 5, 7
    nh, nk = 1 << log_h, 1 << log_k
    rng = O.SplitMix64(9)
    R = O.R_MOD
    a = O.random_fr(rng, nh + 3)
    b = O.random_fr(rng, nh + 5)
    prod = [0] * (len(a) + len(b) - 1)
    for i, x in enumerate(a):
        for j, y in enumerate(b):
            prod[i + j] = (prod[i + j] + x * y) % R
    # c = (a*b) mod (X^nh - 1), q = (a*b - c) / (X^nh - 1)
    c = [0] * nh
    for i, v in enumerate(prod):
        c[i % nh] = (c[i % nh] + v) % R
    diff = [(v - (c[i] if i < nh else 0)) % R for i, v in enumerate(prod)]
    q = [0] * (len(diff) - nh)
    work = diff[:]
    for i in range(len(work) - 1, nh - 1, -1):          # divide by X^nh - 1
        q[i - nh] = work[i]
        work[i - nh] = (work[i - nh] + work[i]) % R
        work[i] = 0
    assert all(v == 0 for v in work)

    def dev(v):
        return torch.from_numpy(H.fr_mont_array(v + [0] * (nk - len(v))).view(np.int64)).cuda()
    h = S.varuna.quotient_on_coset(dev(a), dev(b), dev(c), log_h, log_k)
    torch.cuda.synchronize()
    got = H.fr_from_mont_array(h.cpu().numpy().view(np.uint64).reshape(-1, 4))
    assert got == q + [0] * (nk - len(q))
