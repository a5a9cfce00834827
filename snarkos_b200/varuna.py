"""Device-resident slices of snarkVM's Varuna prover built from the two primitives plus the polynomial glue
[UPSTREAM algorithms/src/snark/varuna/ahp/prover/round_functions/*; SURVEY.md 8a row a10, 8f rank 2].

These are NOT the AHP (no transcript, no constraint system): they are the data-parallel shapes the prover's rounds are
made of, wired so that nothing leaves HBM between a transform, the pointwise stage and the commitment:

  commit_evaluations   round 1:   witness evaluations over H --iFFT--> coefficients --KZG10::commit--> commitments
  quotient_on_coset    rounds 2-4: h(X) = (a(X) * b(X) - c(X)) / v_H(X) via coset FFTs over a larger domain K

All tensors are CUDA tensors of Montgomery Fr limbs ([.., 4] int64) as snarkVM holds them in memory.
"""
from __future__ import annotations

import torch

from . import poly
from .fft import EvaluationDomain
from .kzg10 import KZG10, Powers


def commit_evaluations(powers: Powers, evals: torch.Tensor, blinding=None):
    """evals: [batch, |H|, 4] evaluations over the domain H (natural order).  Interpolates every row in place (one batched
    iFFT) and commits to each coefficient vector against the resident powers.  Returns (coefficients, [commitments])."""
    if evals.dim() != 3:
        raise ValueError("expected [batch, size, 4]")
    dom = EvaluationDomain(evals.shape[1])
    if dom.size != evals.shape[1]:
        raise ValueError("evaluation vectors must have a power-of-two length")
    dom.ifft_in_place(evals)                                   # all rows in one launch set
    commitments = []
    for b in range(evals.shape[0]):
        blind = None if blinding is None else blinding[b]
        commitments.append(KZG10.commit(powers, evals[b], blind))
    return evals, commitments


def quotient_on_coset(a: torch.Tensor, b: torch.Tensor, c: torch.Tensor, log_h: int, log_k: int) -> torch.Tensor:
    """Coefficients (length 2^log_k, zero padded) of h(X) = (a(X) b(X) - c(X)) / v_H(X), for coefficient vectors a, b, c
    already zero-padded to |K| = 2^log_k with deg(a b) < |K|, where v_H(X) = X^|H| - 1 divides a b - c.
    coset FFT x3 -> fused multiply-subtract -> division by the vanishing polynomial on the coset -> coset iFFT."""
    n = 1 << log_k
    for t in (a, b, c):
        if t.numel() * t.element_size() != n * 32:
            raise ValueError("operands must be zero-padded to the coset domain size")
    dom = EvaluationDomain(n)
    ea, eb, ec = a.clone(), b.clone(), c.clone()
    for t in (ea, eb, ec):
        dom.coset_fft_in_place(t)
    poly.mul_sub(ea, eb, ec, out=ea)                           # a*b - c on the coset
    poly.divide_by_vanishing_poly_on_coset_in_place(ea, log_k, log_h)
    dom.coset_ifft_in_place(ea)
    return ea


# ---------------------------------------------------------------------------------------------------------------------
# A complete prover with Varuna's SHAPE, end to end on the device.
#
# This is a STAND-IN protocol, not Varuna's AHP (whose constraint system, matrices and Poseidon transcript live in
# snarkVM sources that are not on disk): a KZG-based argument for the row check  z_a(X) * z_b(X) - z_c(X) = h(X) * v_H(X)
# over a domain H, with exactly the data-parallel steps a Varuna proof is made of, in the same order and on the same
# operand sizes:
#   round 1   witness evaluations over H  --batched iFFT-->  coefficients  --one batched KZG commit-->  3 commitments
#   round 2   coset FFTs over K = 2|H|, pointwise a*b - c, division by v_H on the coset, coset iFFT, commit h
#   round 3   evaluations at the challenge z, linear combination by powers of xi, witness (p - p(z)) / (X - z), commit
#   proof     compressed G1 commitments (48 B each, the encoding a Varuna proof carries) + canonical evaluations
# Challenges come from a hash of the proof bytes so far (Fiat-Shamir; snarkVM uses a Poseidon sponge on the CPU at the
# same points), which forces the same device -> host round trips a real prover has.  The CPU twin used by the tests
# and by bench.py's CPU leg is oracle/rowcheck_prover.py: both must produce byte-identical proofs.
# ---------------------------------------------------------------------------------------------------------------------
import hashlib

import numpy as np

from . import _lib
from .msm import g1_compress

FR_MODULUS = 0x12ab655e9a2ca55660b44d1e5c37b00159aa76fed00000010a11800000000001
_FR_R = (1 << 256) % FR_MODULUS
_FR_RINV = pow(_FR_R, -1, FR_MODULUS)


def challenge(transcript: bytes, label: bytes) -> int:
    """Fiat-Shamir challenge in [0, r): BLAKE2s(label || transcript) widened to 512 bits, reduced mod r."""
    h = hashlib.blake2s(label + transcript).digest() + hashlib.blake2s(b"\x01" + label + transcript).digest()
    return int.from_bytes(h, "little") % FR_MODULUS


def _fr_dev(v: int, device) -> torch.Tensor:
    m = (v * _FR_R) % FR_MODULUS
    return torch.from_numpy(np.frombuffer(m.to_bytes(32, "little"), dtype=np.int64).copy()).reshape(1, 4).to(device)


def _fr_host(t: torch.Tensor) -> int:
    m = int.from_bytes(t.cpu().numpy().tobytes()[:32], "little")
    return (m * _FR_RINV) % FR_MODULUS


class RowCheckProver:
    """Holds the resident SRS (`powers_of_beta_g`, at least |H| powers) for a domain H of size 2^log_h."""

    def __init__(self, powers: Powers, log_h: int):
        self.powers = powers
        self.log_h = log_h
        self.n = 1 << log_h

    def prove(self, witness_evals: torch.Tensor) -> bytes:
        """witness_evals: CUDA [3, |H|, 4] Montgomery evaluations of z_a, z_b, z_c over H with z_a * z_b = z_c pointwise
        (consumed: interpolated in place).  Returns the proof bytes:
            C_a | C_b | C_c | C_h   (4 x 48 B compressed G1)
            z_a(z) | z_b(z) | z_c(z) | h(z)   (4 x 32 B canonical little-endian)
            W   (48 B compressed G1)"""
        n, log_h = self.n, self.log_h
        if tuple(witness_evals.shape) != (3, n, 4):
            raise ValueError("expected [3, |H|, 4]")
        dev = witness_evals.device
        # ---- round 1
        EvaluationDomain(n).ifft_in_place(witness_evals)
        coeffs = witness_evals
        com = KZG10.commit_batch(self.powers, [coeffs[0], coeffs[1], coeffs[2]])
        transcript = g1_compress(com).cpu().numpy().tobytes()                  # device -> host: the first challenge needs it
        # ---- round 2: h = (z_a z_b - z_c) / v_H on the coset of K = 2|H|
        k = torch.zeros((3, 2 * n, 4), dtype=torch.int64, device=dev)
        k[:, :n] = coeffs
        EvaluationDomain(2 * n).coset_fft_in_place(k)
        poly.mul_sub(k[0], k[1], k[2], out=k[0])
        poly.divide_by_vanishing_poly_on_coset_in_place(k[0], log_h + 1, log_h)
        h = k[0]
        EvaluationDomain(2 * n).coset_ifft_in_place(h)
        h = h[:n]                                                              # deg h <= |H| - 2
        com_h = KZG10.commit_batch(self.powers, [h])
        transcript += g1_compress(com_h).cpu().numpy().tobytes()
        z = challenge(transcript, b"z")
        z_dev = _fr_dev(z, dev)
        # ---- round 3: evaluations, combination, opening
        polys = [coeffs[0], coeffs[1], coeffs[2], h]
        evals = [poly.divide_by_linear(p, z_dev)[1] for p in polys]
        ev_bytes = b"".join(_fr_host(e).to_bytes(32, "little") for e in evals)
        transcript += ev_bytes
        xi = challenge(transcript, b"xi")
        xis = torch.cat([_fr_dev(pow(xi, i, FR_MODULUS), dev) for i in range(4)])
        combined = poly.linear_combination(polys, xis)
        w, _, _ = KZG10.open(self.powers, combined, z_dev)
        transcript += g1_compress(w.reshape(1, -1)).cpu().numpy().tobytes()
        return transcript
