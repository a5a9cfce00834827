"""End-to-end stand-in prover with Varuna's shape (snarkos_b200/varuna.py RowCheckProver): NTT -> polynomial glue ->
batched KZG commits -> opening -> compressed wire encoding, all on the device, against its CPU twin
(oracle/rowcheck_prover.py) byte for byte, and verified with the SRS trapdoor."""
import numpy as np
import pytest

from oracle import rowcheck_prover as RP
from oracle import bls12_377 as O

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("log_h", [4, 7])
def test_proof_bytes_identical_and_verify(log_h):
    import torch
    import snarkos_b200 as S
    n = 1 << log_h
    beta = O.random_fr(O.SplitMix64(31 + log_h), 1)[0]
    bases = RP.powers_of_beta(beta, n)
    witness = RP.random_witness(np.random.default_rng(log_h), log_h)
    want = RP.prove(bases, log_h, witness)
    assert len(want) == 4 * 48 + 4 * 32 + 48
    assert RP.verify_with_trapdoor(want, log_h, beta)
    powers = S.Powers(bases)
    prover = S.varuna.RowCheckProver(powers, log_h)
    got = prover.prove(torch.from_numpy(witness.view(np.int64)).cuda())
    assert got == want
    # a corrupted witness must not verify (the quotient is then not a polynomial of degree < |H|)
    bad = bytearray(want)
    bad[200] ^= 1
    assert not RP.verify_with_trapdoor(bytes(bad), log_h, beta)
    powers.release()


@pytest.mark.parametrize("log_h", [12, 16])
def test_proof_bytes_identical_at_varuna_sizes(log_h):
    """the domain sizes a credits.aleo transfer proves at (2^14 .. 2^17): synthetic SRS, CPU twin on the
    host cores, byte-identical proof"""
    import torch
    import snarkos_b200 as S
    n = 1 << log_h
    bases = S.synthetic_bases(n, seed=77)
    witness = RP.random_witness(np.random.default_rng(100 + log_h), log_h)
    powers = S.Powers(bases)
    got = S.varuna.RowCheckProver(powers, log_h).prove(torch.from_numpy(witness.view(np.int64)).cuda())
    want = RP.prove(bases.cpu().numpy(), log_h, witness)
    assert got == want
    powers.release()
