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
