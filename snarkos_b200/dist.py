"""Multi-GPU partitioning of the hot path (one process per GPU, torch.distributed).

  * MSM shards by point range: every rank runs the full Pippenger on its slice of (bases, scalars); the 144-byte
    partial sums are all-gathered and added.  An elliptic-curve addition is not an NCCL reduction op, so it is
    gather + local add; nothing else crosses NVLink.
  * A single large NTT runs four-step: the polynomial, block-distributed in natural order, is viewed as an
    N1 x N2 row-major matrix (index j = j1*N2 + j2, output k = k1 + N1*k2):
        all-to-all  (row slab -> column slab)            local N1-point NTTs over j1, twiddle w_N^(j2*k1)
        all-to-all  (column slab -> k1 slab)             local N2-point NTTs over j2
        all-to-all  (k1 slab -> natural block order)     only when natural order out is requested
    The local transforms are ordinary batched calls of the single-GPU NTT; the inverse transform is the same
    schedule with the inverse roots, and its two local inverse NTTs scale by N1^-1 and N2^-1 = N^-1 in total.
    Forward coset scaling is applied to the natural-order input block, inverse coset scaling to the natural-order
    output block.

The local operations are injectable (`ops`), so the exchange / index logic is tested on CPU with the gloo backend and
the oracle standing in for the kernels (tests/test_dist_cpu.py); on GPUs the default ops call the CUDA library.
"""
from __future__ import annotations

import ctypes
from typing import Callable, Optional

import torch
import torch.distributed as dist

from . import _lib

FR_LIMBS = 4            # int64 limbs per Fr element in a tensor row
PROJECTIVE_BYTES = 144


def shard_range(n: int, rank: int, world: int) -> tuple[int, int]:
    """Contiguous point range [lo, hi) of `rank`: sizes differ by at most one, ranges tile [0, n)."""
    base, extra = divmod(n, world)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


# ----------------------------------------------------------------------------------------------------------------
# default (CUDA) local operations
# ----------------------------------------------------------------------------------------------------------------
class CudaOps:
    """Local kernels through the C ABI, on the current CUDA stream."""

    @staticmethod
    def _stream():
        return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)

    def msm(self, bases: torch.Tensor, scalars: torch.Tensor, stride: int = 104) -> torch.Tensor:
        from .msm import VariableBase
        return VariableBase.msm(bases, scalars, stride)

    def sum_projective(self, points: torch.Tensor) -> torch.Tensor:
        from .msm import sum_projective
        return sum_projective(points)

    def ntt_rows(self, mat: torch.Tensor, log_len: int, direction: int) -> None:
        """in-place NTT of every row of a contiguous [rows, 2^log_len, 4] tensor"""
        rows = mat.shape[0]
        _lib.check(_lib.lib().b200_ntt_fr_bls12_377_device(ctypes.c_void_p(mat.data_ptr()), log_len, rows, 1 << log_len,
                                                           direction, 0, self._stream()))

    def mul_powers(self, mat: torch.Tensor, log_n: int, direction: int, kind: int, rows: int, cols: int, row_base: int,
                   col_base: int) -> None:
        _lib.check(_lib.lib().b200_fr_mul_powers_device(ctypes.c_void_p(mat.data_ptr()), log_n, direction, kind, rows, cols,
                                                        row_base, col_base, self._stream()))


# ----------------------------------------------------------------------------------------------------------------
# MSM
# ----------------------------------------------------------------------------------------------------------------
def msm_sharded(bases_local, scalars_local, group=None, ops=None, stride: int = 104) -> torch.Tensor:
    """sum over all ranks of the local MSMs.  Every rank passes its own point range and gets the full result
    (uint8[144] Jacobian)."""
    ops = ops or CudaOps()
    part = ops.msm(bases_local, scalars_local, stride).contiguous()
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    if world == 1:
        return part
    gathered = torch.empty((world, PROJECTIVE_BYTES), dtype=torch.uint8, device=part.device)
    dist.all_gather_into_tensor(gathered.view(-1), part.view(-1), group=group)
    return ops.sum_projective(gathered)


# ----------------------------------------------------------------------------------------------------------------
# four-step NTT
# ----------------------------------------------------------------------------------------------------------------
def _all_to_all(chunks: torch.Tensor, group) -> torch.Tensor:
    """chunks[d] goes to rank d; returns out with out[s] = the chunk rank s sent here.  chunks: [world, ...] contiguous."""
    out = torch.empty_like(chunks)
    try:
        dist.all_to_all_single(out, chunks, group=group)
    except (RuntimeError, NotImplementedError):          # backends without alltoall (old gloo): emulate with all_gather
        world = dist.get_world_size(group)
        rank = dist.get_rank(group)
        allc = [torch.empty_like(chunks) for _ in range(world)]
        dist.all_gather(allc, chunks, group=group)
        for s in range(world):
            out[s] = allc[s][rank]
    return out


def _exchange_transpose(mat: torch.Tensor, world: int, group) -> torch.Tensor:
    """mat: local slab [R_local, C, 4] of a row-distributed [R, C] matrix.  Returns this rank's slab of the
    transpose, [C_local, R, 4] (C_local = C / world, R = R_local * world): one all-to-all plus local re-tiling."""
    r_local, c, limbs = mat.shape
    c_local = c // world
    # chunk d = columns of rank d: [world, R_local, C_local, 4]
    send = mat.view(r_local, world, c_local, limbs).permute(1, 0, 2, 3).contiguous()
    recv = _all_to_all(send, group) if world > 1 else send
    # recv[s] = rows of rank s, my columns -> [R, C_local, 4] -> transpose -> [C_local, R, 4]
    return recv.reshape(world * r_local, c_local, limbs).permute(1, 0, 2).contiguous()


def ntt_distributed(block: torch.Tensor, log_n: int, direction: int = 0, coset: int = 0, natural_out: bool = True, group=None,
                    ops=None, log_n1: Optional[int] = None) -> torch.Tensor:
    """Four-step (i)NTT of ONE polynomial of 2^log_n Fr elements distributed over the ranks of `group`.

    block: this rank's natural-order block [2^log_n / world, 4] (int64 Montgomery limbs).  Returns this rank's block of
    the result: natural order if natural_out, else the 'k1 slab' layout [N1/world, N2] holding X[k1 + N1*k2] for the
    rank's k1 range (what a following pointwise stage can consume without the third exchange)."""
    ops = ops or CudaOps()
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    rank = dist.get_rank(group) if dist.is_initialized() else 0
    n = 1 << log_n
    n1_log = log_n1 if log_n1 is not None else log_n // 2
    n2_log = log_n - n1_log
    n1, n2 = 1 << n1_log, 1 << n2_log
    if n1 % world or n2 % world:
        raise ValueError("both matrix sides must be divisible by the world size")
    if block.shape[0] * world != n:
        raise ValueError("block must hold 2^log_n / world elements")
    block = block.contiguous()
    per = n // world
    if coset and direction == 0:                                   # distribute_powers(coeffs, g) on the natural-order input
        ops.mul_powers(block, log_n, 0, 1, per, 1, rank * per, 0)
    a = block.view(n1 // world, n2, FR_LIMBS)                      # rows j1 of my slab
    c = _exchange_transpose(a, world, group)                       # [N2/world, N1]: my columns j2, all j1
    ops.ntt_rows(c, n1_log, direction)                             # over j1 -> k1
    ops.mul_powers(c, log_n, direction, 0, n2 // world, n1, rank * (n2 // world), 0)     # w_N^(j2 * k1)
    e = _exchange_transpose(c, world, group)                       # [N1/world, N2]: my k1, all j2
    ops.ntt_rows(e, n2_log, direction)                             # over j2 -> k2 : e[k1, k2] = X[k1 + N1*k2]
    if not natural_out:
        if coset and direction == 1:
            raise ValueError("inverse coset scaling needs natural order out")
        return e
    f = _exchange_transpose(e, world, group)                       # [N2/world, N1]: my k2 range, all k1 = natural block
    out = f.view(per, FR_LIMBS)
    if coset and direction == 1:                                   # distribute_powers(coeffs, g^-1) on the natural-order output
        ops.mul_powers(out, log_n, 1, 1, per, 1, rank * per, 0)
    return out
