"""EvaluationDomain::new and its fields [UPSTREAM algorithms/src/fft/domain.rs; SURVEY.md 8a row a3] against the
known-answer constants of SURVEY appendix A (tests/golden/kat.json) and against the oracle's own domain."""
import numpy as np
import pytest

from oracle import bls12_377 as O
from tests import helpers as H

KAT = H.load_kat()


def _domain(n):
    from snarkos_b200.fft import EvaluationDomain
    return EvaluationDomain.new(n)


def test_roots_of_unity_match_the_known_answers():
    assert _domain(4).group_gen == int(KAT["omega_4"])
    assert _domain(8).group_gen == int(KAT["omega_8"])
    d = _domain(1 << 20)
    assert d.group_gen == int(KAT["omega_2^20"])
    assert d.size_inv == int(KAT["(2^20)^-1"])
    assert d.generator_inv == int(KAT["22^-1"])


@pytest.mark.parametrize("n", [1, 2, 3, 5, 1000, 1 << 16, (1 << 20) + 1])
def test_fields_like_snarkvm(n):
    d = _domain(n)
    assert d.size >= n and d.size & (d.size - 1) == 0 and (d.size == 1 or d.size // 2 < n)
    assert d.size == 1 << d.log_size_of_group
    r = O.R_MOD
    assert pow(d.group_gen, d.size, r) == 1 and (d.size == 1 or pow(d.group_gen, d.size // 2, r) == r - 1)
    assert d.group_gen * d.group_gen_inv % r == 1
    assert d.size_as_field_element * d.size_inv % r == 1
    assert d.generator_inv * 22 % r == 1
    od = O.EvaluationDomain(n)
    assert (d.size, d.group_gen, d.group_gen_inv, d.size_inv, d.generator_inv) == (od.size, od.group_gen, od.group_gen_inv, od.size_inv, od.generator_inv)


def test_elements_vanishing_and_montgomery_image():
    from snarkos_b200.fft import fr_to_mont_limbs
    d = _domain(8)
    el = list(d.elements())
    assert el[0] == 1 and el[1] == d.group_gen and len(set(el)) == 8
    assert all(d.evaluate_vanishing_polynomial(x) == 0 for x in el)
    assert d.evaluate_vanishing_polynomial(5) == (pow(5, 8, O.R_MOD) - 1) % O.R_MOD
    # the in-memory form of the generator 22 (SURVEY 8c golden limbs)
    assert [int(x) for x in fr_to_mont_limbs(22)] == [int(x) for x in KAT["fr_gen22_mont_limbs"]]


def test_domain_larger_than_the_two_adicity():
    from snarkos_b200.fft import EvaluationDomain
    with pytest.raises(ValueError):
        EvaluationDomain.new((1 << 47) + 1)
