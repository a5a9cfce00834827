"""CPU twin (TEST INFRASTRUCTURE) of the stand-in row-check prover in snarkos_b200/varuna.py: the same protocol, step
for step, on the C oracle (oracle.c: NTT, batched::msm port, compressed encoding) with Python big-ints for the glue.
It is NOT a restatement of Varuna's AHP (not on disk); it exists so that the device prover -- NTT, polynomial glue,
batched KZG commits, opening, wire encoding chained together -- is checked byte for byte against an independent CPU
implementation, and so that bench.py has a CPU leg for the same workload.  Only tests/ and bench.py import it."""
import hashlib

import numpy as np

from . import bls12_377 as O
from . import c_oracle as C

R = O.R_MOD


def challenge(transcript: bytes, label: bytes) -> int:
    h = hashlib.blake2s(label + transcript).digest() + hashlib.blake2s(b"\x01" + label + transcript).digest()
    return int.from_bytes(h, "little") % R


def _to_ints(mont: np.ndarray):
    raw = np.ascontiguousarray(C.fr_from_mont(mont), dtype=np.uint64).tobytes()
    return [int.from_bytes(raw[i:i + 32], "little") for i in range(0, len(raw), 32)]


def _scalars(vals) -> np.ndarray:
    return np.frombuffer(b"".join(int(v).to_bytes(32, "little") for v in vals), dtype=np.uint64).reshape(-1, 4)


def _mont(vals) -> np.ndarray:
    return np.frombuffer(b"".join(O.fr_to_mont(int(v)).to_bytes(32, "little") for v in vals), dtype=np.uint64).reshape(-1, 4).copy()


def _commit(bases: np.ndarray, canonical: np.ndarray) -> bytes:
    n = canonical.shape[0]
    jac = C.msm_batched(bases, canonical) if n >= 15 else C.msm(bases, canonical)
    return C.g1_compress(jac)[0].tobytes()


def prove(bases: np.ndarray, log_h: int, witness_evals: np.ndarray) -> bytes:
    """bases: uint8 G1Affine images of the powers (>= 2^log_h of them); witness_evals: uint64 [3, 2^log_h, 4] Montgomery
    evaluations of z_a, z_b, z_c over H.  Returns the proof bytes (layout: snarkos_b200/varuna.py RowCheckProver.prove)."""
    n = 1 << log_h
    coeffs = [C.ntt(witness_evals[i], log_h, direction=1) for i in range(3)]
    transcript = b"".join(_commit(bases, C.fr_from_mont(c)) for c in coeffs)
    # h = (z_a z_b - z_c) / v_H on the coset g K, |K| = 2 |H|:  v_H(g w^i) = g^|H| (-1)^i - 1
    pad = np.zeros((2 * n, 4), dtype=np.uint64)
    ev = []
    for c in coeffs:
        buf = pad.copy()
        buf[:n] = c
        ev.append(C.ntt(buf, log_h + 1, coset=1))
    q = C.fr_sub(C.fr_mul(ev[0], ev[1]), ev[2])
    gh = pow(O.FR_GENERATOR, n, R)
    inv = _mont([pow((gh - 1) % R, -1, R), pow((-gh - 1) % R, -1, R)])
    q = C.fr_mul(q, np.tile(inv, (n, 1)))
    h = C.ntt(q, log_h + 1, direction=1, coset=1)
    assert not h[n - 1:].any(), "z_a z_b - z_c is not divisible by v_H (bad witness)"
    h = np.ascontiguousarray(h[:n])
    transcript += _commit(bases, C.fr_from_mont(h))
    z = challenge(transcript, b"z")
    polys = [_to_ints(c) for c in coeffs] + [_to_ints(h)]
    evals = []
    for p in polys:
        acc = 0
        for c in reversed(p):
            acc = (acc * z + c) % R
        evals.append(acc)
    transcript += b"".join(v.to_bytes(32, "little") for v in evals)
    xi = challenge(transcript, b"xi")
    xs = [pow(xi, i, R) for i in range(4)]
    comb = [sum(x * p[i] for x, p in zip(xs, polys)) % R for i in range(n)]
    w, carry = [0] * (n - 1), 0
    for j in range(n - 1, 0, -1):
        carry = (comb[j] + z * carry) % R
        w[j - 1] = carry
    transcript += _commit(bases, _scalars(w)) if n > 1 else C.g1_compress(C.msm(bases, _scalars([])))[0].tobytes()
    return transcript


def verify_with_trapdoor(proof: bytes, log_h: int, beta: int) -> bool:
    """Designated-verifier check for an SRS powers[i] = beta^i * G with KNOWN beta (tests only; a real verifier uses two
    pairings): the row-check identity at z, and the KZG opening in the exponent,
        sum_i xi^i C_i - p(z) G == (beta - z) W."""
    n = 1 << log_h
    coms = [O.g1_decompress(proof[48 * i:48 * i + 48]) for i in range(4)]
    evals = [int.from_bytes(proof[192 + 32 * i:224 + 32 * i], "little") for i in range(4)]
    W = O.g1_decompress(proof[320:368])
    z = challenge(proof[:192], b"z")
    xi = challenge(proof[:320], b"xi")
    if (evals[0] * evals[1] - evals[2] - evals[3] * (pow(z, n, R) - 1)) % R != 0:
        return False
    lhs, pz = None, 0
    for i in range(4):
        lhs = O.g1_add(lhs, O.g1_mul(coms[i], pow(xi, i, R)))
        pz = (pz + pow(xi, i, R) * evals[i]) % R
    lhs = O.g1_add(lhs, O.g1_neg(O.g1_mul(O.G1_GEN, pz)))
    return lhs == O.g1_mul(W, (beta - z) % R)


def powers_of_beta(beta: int, n: int) -> np.ndarray:
    """G1Affine images of beta^i * G, i < n (Python big-int scalar multiplications: small n only)"""
    out, acc = [], 1
    for _ in range(n):
        out.append(O.g1_mul(O.G1_GEN, acc))
        acc = acc * beta % R
    return np.frombuffer(b"".join(O.affine_bytes(p) for p in out), dtype=np.uint8).copy()


def random_witness(rng: np.random.Generator, log_h: int) -> np.ndarray:
    """Montgomery evaluations [3, 2^log_h, 4] of z_a, z_b and z_c = z_a * z_b"""
    n = 1 << log_h
    s = rng.integers(0, 1 << 63, size=(2, n, 4), dtype=np.uint64) * np.uint64(2) + rng.integers(0, 2, size=(2, n, 4), dtype=np.uint64)
    s[..., 3] &= np.uint64((1 << 60) - 1)
    out = np.zeros((3, n, 4), dtype=np.uint64)
    out[:2] = s
    out[2] = C.fr_mul(s[0], s[1])
    return out
