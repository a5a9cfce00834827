"""The CPU twin of the stand-in row-check prover (oracle/rowcheck_prover.py) on its own: a proof verifies against the
SRS trapdoor, a tampered proof or a witness that violates z_a * z_b = z_c does not."""
import numpy as np
import pytest

from oracle import bls12_377 as O
from oracle import rowcheck_prover as RP


@pytest.mark.parametrize("log_h", [1, 4, 6])
def test_twin_proves_and_verifies(log_h):
    beta = O.random_fr(O.SplitMix64(log_h), 1)[0]
    bases = RP.powers_of_beta(beta, 1 << log_h)
    w = RP.random_witness(np.random.default_rng(log_h), log_h)
    proof = RP.prove(bases, log_h, w)
    assert len(proof) == 368
    assert RP.verify_with_trapdoor(proof, log_h, beta)
    for pos in (3, 100, 200, 330):
        bad = bytearray(proof)
        bad[pos] ^= 1
        try:
            ok = RP.verify_with_trapdoor(bytes(bad), log_h, beta)
        except (ValueError, AssertionError):
            ok = False                      # a flipped bit can make a commitment undecodable
        assert not ok
    if log_h >= 2:
        w[2, 0, 0] ^= np.uint64(1)
        with pytest.raises(AssertionError):
            RP.prove(bases, log_h, w)
