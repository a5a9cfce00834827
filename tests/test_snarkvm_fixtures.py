"""Consumers of the fixture files of tests/golden/FIXTURES.md.

`snarkvm_*.bin` are written by snarkVM itself (rust/snarkvm-algorithms-b200/src/bin/make_fixtures.rs, rev dea322b): when
present they pin the C oracle (CPU test) and the CUDA path through the C ABI (GPU test) to the reference.  They cannot
be produced in this image (no Rust toolchain), so those tests SKIP with a loud "parity unpinned" reason when none
exist.  `oraclefmt_*.bin` (Python big-int oracle, same format) go through the same code in every run."""
import numpy as np
import pytest

from oracle import bls12_377 as O
from oracle import c_oracle as C
from tests.golden import fixture_format as F

UNPINNED = ("parity unpinned: no snarkVM-generated fixtures under tests/golden/ (needs cargo + network: "
            "`cargo run --release --features harness --bin make_fixtures` in rust/snarkvm-algorithms-b200)")


def _check_layout(fx):
    assert fx["stride"] >= 97 and fx["stride"] % 8 == 0, f"{fx['path']}: G1Affine stride {fx['stride']}"
    assert fx["offsets"] == (0, 48, 96), f"{fx['path']}: G1Affine field offsets {fx['offsets']} != (0, 48, 96): the C ABI's layout assumption is wrong"
    assert fx["proj_bytes"] == 144


def _affine_of_jac(jac: bytes):
    return O.jacobian_from_bytes(jac)


def _check_msm_fixture(fx, msm_fn, compress_fn=None, normalize_fn=None):
    """msm_fn(bases_u8, scalars_u64x4, stride) -> 144 B Jacobian image"""
    _check_layout(fx)
    n, stride = fx["n"], fx["stride"]
    bases = np.frombuffer(fx["bases"], dtype=np.uint8)
    scalars = np.frombuffer(fx["scalars"], dtype=np.uint64).reshape(n, 4)
    want = _affine_of_jac(fx["result_jac"])
    # the fixture is self-consistent: to_affine() and the compressed form describe the same point
    assert O.affine_from_bytes(fx["result_affine"][:104] if stride >= 104 else fx["result_affine"]) == want
    assert O.g1_compress(want) == fx["compressed"], f"{fx['path']}: compressed encoding differs from the restated rule"
    assert O.g1_decompress(fx["compressed"]) == want
    got_jac = msm_fn(bases, scalars, stride)
    assert _affine_of_jac(bytes(got_jac)) == want, f"{fx['path']}: MSM result differs"
    if compress_fn is not None:
        assert bytes(compress_fn(got_jac)) == fx["compressed"]
        assert bytes(compress_fn(np.frombuffer(fx["result_jac"], dtype=np.uint8))) == fx["compressed"]
    if normalize_fn is not None:
        aff = bytes(normalize_fn(got_jac, stride))
        assert aff[:97] == fx["result_affine"][:97] if want is not None else aff[96] == 1


def _check_ntt_fixture(fx, ntt_fn):
    """ntt_fn(input_u64[n_in, 4], log_n, direction, coset) -> uint64 [2^log_n, 4]"""
    for key, direction, coset in (("fft", 0, 0), ("ifft", 1, 0), ("coset_fft", 0, 1), ("coset_ifft", 1, 1)):
        got = ntt_fn(fx["input"], fx["log_n"], direction, coset)
        assert np.array_equal(np.asarray(got, dtype=np.uint64).reshape(-1, 4), fx[key]), f"{fx['path']}: {key} differs"


# ---- C oracle (CPU) ---------------------------------------------------------------------------
def _oracle_msm(bases, scalars, stride):
    return C.msm_batched(bases, scalars, stride) if scalars.shape[0] >= 15 else C.msm(bases, scalars, stride)


def _oracle_ntt(inp, log_n, direction, coset):
    buf = np.zeros((1 << log_n, 4), dtype=np.uint64)
    buf[: inp.shape[0]] = inp
    return C.ntt(buf, log_n, 1, direction, coset)


def _oracle_compress(jac):
    return C.g1_compress(np.asarray(jac, dtype=np.uint8).reshape(1, 144))[0]


def _oracle_normalize(jac, stride):
    return C.g1_to_affine(np.asarray(jac, dtype=np.uint8).reshape(1, 144), stride)[0]


def test_format_samples_against_the_c_oracle():
    msm_files, ntt_files = F.find("oraclefmt")
    assert len(msm_files) >= 6 and len(ntt_files) >= 5, "run python tests/golden/make_format_samples.py"
    for p in msm_files:
        _check_msm_fixture(F.read_msm(p), _oracle_msm, _oracle_compress, _oracle_normalize)
    for p in ntt_files:
        _check_ntt_fixture(F.read_ntt(p), _oracle_ntt)


def test_snarkvm_fixtures_against_the_c_oracle():
    msm_files, ntt_files = F.find("snarkvm")
    if not msm_files and not ntt_files:
        pytest.skip(UNPINNED)
    for p in msm_files:
        _check_msm_fixture(F.read_msm(p), _oracle_msm, _oracle_compress, _oracle_normalize)
    for p in ntt_files:
        _check_ntt_fixture(F.read_ntt(p), _oracle_ntt)


# ---- CUDA path through the C ABI (GPU) ----------------------------------------------------------
def _gpu_fns():
    import snarkos_b200 as S

    def msm(bases, scalars, stride):
        return S.VariableBase.msm(bases, scalars, stride=stride) if scalars.shape[0] else S.VariableBase.msm(bases, scalars)

    def ntt(inp, log_n, direction, coset):
        d = S.EvaluationDomain(1 << log_n)
        return [d.fft_in_place, d.ifft_in_place, d.coset_fft_in_place, d.coset_ifft_in_place][direction + 2 * coset](inp.copy())

    return msm, ntt, (lambda jac: S.g1_compress(np.asarray(jac, dtype=np.uint8).reshape(1, 144))[0]), \
        (lambda jac, stride: S.g1_batch_normalize(np.asarray(jac, dtype=np.uint8).reshape(1, 144), stride=stride)[0])


@pytest.mark.gpu
def test_format_samples_against_the_cuda_path():
    msm, ntt, compress, normalize = _gpu_fns()
    msm_files, ntt_files = F.find("oraclefmt")
    for p in msm_files:
        _check_msm_fixture(F.read_msm(p), msm, compress, normalize)
    for p in ntt_files:
        _check_ntt_fixture(F.read_ntt(p), ntt)


@pytest.mark.gpu
def test_snarkvm_fixtures_against_the_cuda_path():
    msm_files, ntt_files = F.find("snarkvm")
    if not msm_files and not ntt_files:
        pytest.skip(UNPINNED)
    msm, ntt, compress, normalize = _gpu_fns()
    for p in msm_files:
        _check_msm_fixture(F.read_msm(p), msm, compress, normalize)
    for p in ntt_files:
        _check_ntt_fixture(F.read_ntt(p), ntt)
