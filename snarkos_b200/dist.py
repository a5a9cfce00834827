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

  * `ntt_distributed_fused` is the same schedule with every exchange done by ONE kernel
    (`b200_fr_exchange_transpose_device`): transpose + all-to-all (+ the twiddle) in a single pass whose stores go
    straight into the destination rank's slab over NVLink (peer memory mapped through CUDA IPC, `PeerExchange`);
    NCCL only provides the barriers between writers and readers of a slab.

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
    if coset and direction == 1 and not natural_out:               # checked before anything is enqueued
        raise ValueError("inverse coset scaling needs natural order out")
    block = block.contiguous()
    per = n // world
    if coset and direction == 0:                                   # distribute_powers(coeffs, g) on the natural-order input
        block = block.clone()                                      # `block` is an input: the caller's tensor is not modified
        ops.mul_powers(block, log_n, 0, 1, per, 1, rank * per, 0)
    a = block.view(n1 // world, n2, FR_LIMBS)                      # rows j1 of my slab
    c = _exchange_transpose(a, world, group)                       # [N2/world, N1]: my columns j2, all j1
    ops.ntt_rows(c, n1_log, direction)                             # over j1 -> k1
    ops.mul_powers(c, log_n, direction, 0, n2 // world, n1, rank * (n2 // world), 0)     # w_N^(j2 * k1)
    e = _exchange_transpose(c, world, group)                       # [N1/world, N2]: my k1, all j2
    ops.ntt_rows(e, n2_log, direction)                             # over j2 -> k2 : e[k1, k2] = X[k1 + N1*k2]
    if not natural_out:
        return e
    f = _exchange_transpose(e, world, group)                       # [N2/world, N1]: my k2 range, all k1 = natural block
    out = f.view(per, FR_LIMBS)
    if coset and direction == 1:                                   # distribute_powers(coeffs, g^-1) on the natural-order output
        ops.mul_powers(out, log_n, 1, 1, per, 1, rank * per, 0)
    return out


# ----------------------------------------------------------------------------------------------------------------
# fused four-step: exchanges as peer-memory stores
# ----------------------------------------------------------------------------------------------------------------
class _DevMem:
    """raw device memory as a torch tensor (zero copy) through __cuda_array_interface__"""

    def __init__(self, ptr: int, nbytes: int):
        self.__cuda_array_interface__ = {"shape": (nbytes // 8,), "typestr": "<i8", "data": (ptr, False), "version": 3}


def _as_tensor(ptr: int, elems: int, device) -> torch.Tensor:
    return torch.as_tensor(_DevMem(ptr, elems * 32), device=device).view(elems, FR_LIMBS)


class PeerExchange:
    """Two exchange slabs (B, C) of `per_elems` Fr elements on every rank, allocated by the library with cudaMalloc and
    mapped into every other rank's address space through CUDA IPC, so that a kernel on one GPU stores into the slab of
    another.  `barrier()` is a one-element NCCL all-reduce on the current stream: it orders the writers of a slab before
    its readers without blocking the host."""

    NBUF = 2

    def __init__(self, per_elems: int, group=None):
        self.group = group
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        self.per = per_elems
        self.device = torch.device("cuda", torch.cuda.current_device())
        L = _lib.lib()
        self._own, handles = [], torch.zeros((self.NBUF, 64), dtype=torch.uint8)
        for b in range(self.NBUF):
            ptr = ctypes.c_void_p()
            h = (ctypes.c_uint8 * 64)()
            _lib.check(L.b200_peer_buffer_alloc(per_elems * 32, ctypes.byref(ptr), h))
            self._own.append(ptr.value)
            handles[b] = torch.frombuffer(bytearray(h), dtype=torch.uint8)
        self.ptrs = [[None] * self.world for _ in range(self.NBUF)]
        self._opened = []
        if self.world > 1:
            allh = torch.empty((self.world, self.NBUF, 64), dtype=torch.uint8, device=self.device)
            dist.all_gather_into_tensor(allh.view(-1), handles.to(self.device).view(-1), group=group)
            allh = allh.cpu()
        else:
            allh = handles.view(1, self.NBUF, 64)
        for d in range(self.world):
            for b in range(self.NBUF):
                if d == self.rank:
                    self.ptrs[b][d] = self._own[b]
                    continue
                raw = (ctypes.c_uint8 * 64).from_buffer_copy(bytes(allh[d, b].numpy().tobytes()))
                ptr = ctypes.c_void_p()
                _lib.check(L.b200_peer_buffer_open(raw, ctypes.byref(ptr)))
                self.ptrs[b][d] = ptr.value
                self._opened.append(ptr.value)
        self._flag = torch.zeros(1, dtype=torch.int32, device=self.device)

    def dst_array(self, b: int):
        return (ctypes.c_void_p * self.world)(*self.ptrs[b])

    def local(self, b: int) -> torch.Tensor:
        return _as_tensor(self._own[b], self.per, self.device)

    def barrier(self) -> None:
        if self.world > 1:
            dist.all_reduce(self._flag, group=self.group)

    def exchange_stream(self) -> "torch.cuda.Stream":
        if getattr(self, "_xstream", None) is None:
            self._xstream = torch.cuda.Stream(device=self.device, priority=-1)
        return self._xstream

    def transform_stream(self) -> "torch.cuda.Stream":
        if getattr(self, "_tstream", None) is None:
            self._tstream = torch.cuda.Stream(device=self.device)
        return self._tstream

    def close(self) -> None:
        torch.cuda.synchronize()
        if self.world > 1:
            dist.barrier(group=self.group)
        L = _lib.lib()
        for p in self._opened:
            _lib.check(L.b200_peer_buffer_close(ctypes.c_void_p(p)))
        for p in self._own:
            _lib.check(L.b200_peer_buffer_free(ctypes.c_void_p(p)))
        self._opened, self._own = [], []


def ntt_rows_exchange(rows: torch.Tensor, dst_ptrs, world: int, rank: int, r_local: int, log_len: int, log_n: int, direction: int,
                      twiddle: bool, row_base: int = 0) -> None:
    """row transforms + exchange in one launch set: rows = this rank's [r_local, 2^log_len] slab (left unchanged); the last
    pass of the transforms stores every output where `exchange_transpose` would put it (b200_ntt_rows_exchange_device)"""
    _lib.check(_lib.lib().b200_ntt_rows_exchange_device(
        ctypes.c_void_p(rows.data_ptr()), dst_ptrs, world, rank, r_local, log_len, log_n, direction, 1 if twiddle else 0, row_base,
        ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)))


def exchange_transpose(src: torch.Tensor, dst_ptrs, world: int, rank: int, r_local: int, c: int, log_n: int, direction: int,
                       twiddle: bool, row_base: int = 0) -> None:
    """one fused exchange: src = this rank's [r_local, c] slab; dst_ptrs[d] = rank d's [c / world, r_local * world] slab"""
    _lib.check(_lib.lib().b200_fr_exchange_transpose_device(
        ctypes.c_void_p(src.data_ptr()), dst_ptrs, world, rank, r_local, c, log_n, direction, 1 if twiddle else 0, row_base,
        ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)))


def exchange_transpose_part(src_slab: torch.Tensor, dst_ptrs, world: int, rank: int, r_local: int, row_off: int, r_count: int,
                            c: int, col_lo: int, col_cnt: int, log_n: int, direction: int, twiddle: bool, row_base: int = 0,
                            cta_limit: int = 0) -> None:
    """a part of one fused exchange: rows [row_off, row_off + r_count) of the slab, columns [col_lo, col_lo + col_cnt)
    of every destination rank's range"""
    _lib.check(_lib.lib().b200_fr_exchange_transpose_part_device(
        ctypes.c_void_p(src_slab.data_ptr()), dst_ptrs, world, rank, r_local, row_off, r_count, c, col_lo, col_cnt, log_n,
        direction, 1 if twiddle else 0, row_base, cta_limit, ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)))


def ntt_distributed_overlapped(block: torch.Tensor, log_n: int, fabric: PeerExchange, direction: int = 0, coset: int = 0,
                               natural_out: bool = True, log_n1: Optional[int] = None, chunks: int = 4,
                               cta_limit: Optional[int] = None) -> torch.Tensor:
    """`ntt_distributed_fused` with every slab cut into `chunks` parts and two streams, so that the exchanges run under
    the local transforms:
        exchange stream : X1(0) | X1(1) | ... | X2(0) X2(1) ...      | X3(0) X3(1) ...
        transform stream:        NTT_B(0) NTT_B(1) ...               NTT_C(0) NTT_C(1) ...
    X1(q) delivers rows chunk q of every rank's B (a column sub-range of every destination), after a barrier NTT_B(q)
    transforms them and X2(q) sends them on (twiddled) while NTT_B(q+1) runs; the N2-point transforms need all of C, so
    one barrier separates the phases; X3(q) follows NTT_C(q).  Same results as the unchunked schedule, bit for bit."""
    ops = CudaOps()
    world, rank = fabric.world, fabric.rank
    n = 1 << log_n
    n1_log = log_n1 if log_n1 is not None else log_n // 2
    n2_log = log_n - n1_log
    n1, n2 = 1 << n1_log, 1 << n2_log
    if n1 % world or n2 % world:
        raise ValueError("both matrix sides must be divisible by the world size")
    per = n // world
    if block.shape[0] != per or fabric.per != per:
        raise ValueError("block and fabric must hold 2^log_n / world elements")
    if coset and direction == 1 and not natural_out:               # checked before anything is enqueued
        raise ValueError("inverse coset scaling needs natural order out")
    if cta_limit is None:                                           # two exchange CTAs per SM leave room for the transforms
        cta_limit = 2 * torch.cuda.get_device_properties(block.device).multi_processor_count
    rb, rc = n2 // world, n1 // world                               # rows of my B slab / of my C slab
    q_n = max(1, min(chunks, rb, rc))
    while rb % q_n or rc % q_n:
        q_n -= 1
    block = block.contiguous()
    cur = torch.cuda.current_stream()
    xs = fabric.exchange_stream()                                   # exchanges + barriers: HIGH priority, so that their CTAs
    ts = fabric.transform_stream()                                  # get SM slots while the (compute-bound) transforms run
    if coset and direction == 0:
        block = block.clone()                                       # the caller's tensor is not modified
        ops.mul_powers(block, log_n, 0, 1, per, 1, rank * per, 0)
    B, C = fabric.local(0), fabric.local(1)
    Bm, Cm = B.view(rb, n1, FR_LIMBS), C.view(rc, n2, FR_LIMBS)
    dst_b, dst_c = fabric.dst_array(0), fabric.dst_array(1)
    ready = torch.cuda.Event()
    ready.record(cur)
    arrived, done_b, done_c = [], [], []
    with torch.cuda.stream(xs):
        xs.wait_event(ready)
        fabric.barrier()                                           # every rank is done with the slabs of the previous call
        for q in range(q_n):                                       # X1(q): B rows [q * rb/Q, ..) on every rank
            exchange_transpose_part(block, dst_b, world, rank, rc, 0, rc, n2, q * (rb // q_n), rb // q_n, log_n, direction, False, 0, cta_limit)
            fabric.barrier()
            ev = torch.cuda.Event()
            ev.record(xs)
            arrived.append(ev)
    with torch.cuda.stream(ts):
        for q in range(q_n):
            ts.wait_event(arrived[q])
            ops.ntt_rows(Bm[q * (rb // q_n):(q + 1) * (rb // q_n)], n1_log, direction)
            ev = torch.cuda.Event()
            ev.record(ts)
            done_b.append(ev)
    ev_c = torch.cuda.Event()
    with torch.cuda.stream(xs):
        for q in range(q_n):                                       # X2(q): my B rows chunk q, twiddled, into everybody's C
            xs.wait_event(done_b[q])
            exchange_transpose_part(B, dst_c, world, rank, rb, q * (rb // q_n), rb // q_n, n1, 0, rc, log_n, direction, True, rank * rb, cta_limit)
        fabric.barrier()                                           # C complete everywhere
        ev_c.record(xs)
    with torch.cuda.stream(ts):
        ts.wait_event(ev_c)
        for q in range(q_n):
            ops.ntt_rows(Cm[q * (rc // q_n):(q + 1) * (rc // q_n)], n2_log, direction)
            ev = torch.cuda.Event()
            ev.record(ts)
            done_c.append(ev)
    if not natural_out:
        cur.wait_event(done_c[-1])
        return Cm
    fin = torch.cuda.Event()
    with torch.cuda.stream(xs):
        for q in range(q_n):                                       # X3(q): my C rows chunk q into everybody's B
            xs.wait_event(done_c[q])
            exchange_transpose_part(C, dst_b, world, rank, rc, q * (rc // q_n), rc // q_n, n2, 0, rb, log_n, direction, False, 0, cta_limit)
        fabric.barrier()
        fin.record(xs)
    cur.wait_event(fin)
    out = B.view(per, FR_LIMBS)
    if coset and direction == 1:
        ops.mul_powers(out, log_n, 1, 1, per, 1, rank * per, 0)
    return out


def ntt_distributed_fused(block: torch.Tensor, log_n: int, fabric: PeerExchange, direction: int = 0, coset: int = 0,
                          natural_out: bool = True, log_n1: Optional[int] = None, fuse_transforms: bool = True) -> torch.Tensor:
    """`ntt_distributed` with the exchanges fused into peer-memory stores.  Returns a VIEW of one of the fabric's
    slabs (valid until the next call on the same fabric): natural-order block, or the k1 slab if not natural_out.
    fuse_transforms: the second and third exchange are the LAST PASS of the row transforms before them (one launch set:
    butterflies, twiddle and peer stores; no local slab written and read again); False keeps them as separate kernels."""
    ops = CudaOps()
    world, rank = fabric.world, fabric.rank
    n = 1 << log_n
    n1_log = log_n1 if log_n1 is not None else log_n // 2
    n2_log = log_n - n1_log
    n1, n2 = 1 << n1_log, 1 << n2_log
    if n1 % world or n2 % world:
        raise ValueError("both matrix sides must be divisible by the world size")
    per = n // world
    if block.shape[0] != per or fabric.per != per:
        raise ValueError("block and fabric must hold 2^log_n / world elements")
    if coset and direction == 1 and not natural_out:               # checked before anything is enqueued
        raise ValueError("inverse coset scaling needs natural order out")
    block = block.contiguous()
    if coset and direction == 0:
        block = block.clone()                                      # the caller's tensor is not modified
        ops.mul_powers(block, log_n, 0, 1, per, 1, rank * per, 0)
    B, C = fabric.local(0), fabric.local(1)
    fabric.barrier()                                               # every rank is done with the slabs of the previous call
    exchange_transpose(block, fabric.dst_array(0), world, rank, n1 // world, n2, log_n, direction, False)
    fabric.barrier()                                               # B = [N2/world, N1]: my columns j2, all j1
    if fuse_transforms:
        ntt_rows_exchange(B, fabric.dst_array(1), world, rank, n2 // world, n1_log, log_n, direction, True, rank * (n2 // world))
    else:
        ops.ntt_rows(B.view(n2 // world, n1, FR_LIMBS), n1_log, direction)
        exchange_transpose(B, fabric.dst_array(1), world, rank, n2 // world, n1, log_n, direction, True, rank * (n2 // world))
    fabric.barrier()                                               # C = [N1/world, N2]: my k1, all j2, twiddled
    if not natural_out:
        ops.ntt_rows(C.view(n1 // world, n2, FR_LIMBS), n2_log, direction)
        return C.view(n1 // world, n2, FR_LIMBS)
    if fuse_transforms:
        ntt_rows_exchange(C, fabric.dst_array(0), world, rank, n1 // world, n2_log, log_n, direction, False)
    else:
        ops.ntt_rows(C.view(n1 // world, n2, FR_LIMBS), n2_log, direction)
        exchange_transpose(C, fabric.dst_array(0), world, rank, n1 // world, n2, log_n, direction, False)
    fabric.barrier()                                               # B = [N2/world, N1]: my k2 range = natural block
    out = B.view(per, FR_LIMBS)
    if coset and direction == 1:
        ops.mul_powers(out, log_n, 1, 1, per, 1, rank * per, 0)
    return out
