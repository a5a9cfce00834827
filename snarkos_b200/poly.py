"""Device-resident polynomial glue between (i)FFTs, mirroring what snarkVM's Varuna prover does on `Evaluations` /
`DensePolynomial` vectors between transforms [UPSTREAM algorithms/src/fft/evaluations.rs, fields batch_inversion,
fft/domain.rs divide_by_vanishing_poly_on_coset_in_place; SURVEY.md 8f rank 2].  All arguments are CUDA tensors of
Montgomery Fr limbs ([n, 4] int64); everything is enqueued on the current stream and nothing touches the host."""
from __future__ import annotations

import ctypes

import torch

from . import _lib


def _p(t):
    return ctypes.c_void_p(t.data_ptr()) if t is not None else None


def _s():
    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)


def _n(t) -> int:
    return t.numel() * t.element_size() // 32


def _op(op, a, b, c=None, out=None):
    out = a if out is None else out
    scalar = 1 if _n(b) == 1 and _n(a) != 1 else 0
    _lib.check(_lib.lib().b200_fr_vec_op_device(op, _p(out), _p(a), _p(b), _p(c), _n(a), scalar, _s()))
    return out


def mul(a, b, out=None):
    """out = a * b element-wise (b may be a single broadcast element); in place on `a` by default."""
    return _op(0, a, b, out=out)


def add(a, b, out=None):
    return _op(1, a, b, out=out)


def sub(a, b, out=None):
    return _op(2, a, b, out=out)


def mul_add(a, b, c, out=None):
    """out = a * b + c"""
    return _op(3, a, b, c, out=out)


def mul_sub(a, b, c, out=None):
    """out = a * b - c"""
    return _op(4, a, b, c, out=out)


def batch_inversion(a):
    """in-place 1/a element-wise, zeros stay zero (snarkVM `batch_inversion`)."""
    _lib.check(_lib.lib().b200_fr_batch_inverse_device(_p(a), _n(a), _s()))
    return a


def divide_by_vanishing_poly_on_coset_in_place(evals, log_k: int, log_h: int):
    """evals over the coset 22*K (|K| = 2^log_k) divided by the vanishing polynomial of H (|H| = 2^log_h)."""
    if _n(evals) != 1 << log_k:
        raise ValueError("evaluation vector must have the size of the coset domain")
    _lib.check(_lib.lib().b200_fr_divide_by_vanishing_on_coset_device(_p(evals), log_k, log_h, _s()))
    return evals


def linear_combination(polys, coeffs, out_len: int | None = None):
    """sum_m coeffs[m] * polys[m] (polynomials of different lengths, zero-extended): the combination by opening
    challenges in `KZG10::open` / `SonicKZG10::open_combinations`.  polys: list of CUDA [n_m, 4] tensors; coeffs: CUDA
    [k, 4] Montgomery Fr."""
    k = len(polys)
    lens = [_n(p) for p in polys]
    out_len = max(lens) if out_len is None else out_len
    flat = torch.cat([p.contiguous().reshape(-1) for p in polys if p.numel()]) if any(lens) else torch.empty(0, dtype=torch.int64, device=coeffs.device)
    off = (ctypes.c_uint64 * (k + 1))(*([0] + [sum(lens[: i + 1]) for i in range(k)]))
    out = torch.empty((out_len, 4), dtype=torch.int64, device=coeffs.device)
    _lib.check(_lib.lib().b200_fr_linear_combination_device(_p(out), _p(flat), off, k, _p(coeffs), out_len, _s()))
    return out


def divide_by_linear(poly, point):
    """(quotient, remainder) of poly by (X - point); remainder = poly(point).  CUDA [n, 4] and [1, 4] Montgomery Fr."""
    n = _n(poly)
    q = torch.empty((max(n - 1, 0), 4), dtype=torch.int64, device=poly.device)
    rem = torch.empty((1, 4), dtype=torch.int64, device=poly.device)
    _lib.check(_lib.lib().b200_fr_divide_by_linear_device(_p(q) if n > 1 else None, _p(poly), n, _p(point), _p(rem), _s()))
    return q, rem
