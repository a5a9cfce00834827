"""CPU-side verification of the DEVICE NTT pass code (snarkos_b200/csrc/ntt_core.cuh + ntt_plan.h): the per-phase
functions the CUDA kernel executes are run tile by tile on the host (tests/host/ntt_host.cpp) and compared with
the oracle for every decomposition shape (1..4 passes, several tile widths, all four transform kinds)."""
import ctypes
import os
import subprocess

import numpy as np
import pytest

from oracle import bls12_377 as O
from oracle import c_oracle as C
from tests import helpers as H

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)


@pytest.fixture(scope="module")
def hostlib():
    src = os.path.join(HERE, "host", "ntt_host.cpp")
    so = os.path.join(HERE, "host", "libntt_host.so")
    deps = [src] + [os.path.join(ROOT, "snarkos_b200", "csrc", f) for f in ("field.cuh", "ntt_core.cuh", "ntt_plan.h", "ptx_ops.cuh")]
    if not os.path.exists(so) or any(os.path.getmtime(d) > os.path.getmtime(so) for d in deps):
        subprocess.run(["g++", "-O2", "-std=c++17", "-shared", "-fPIC", "-x", "c++", src, "-o", so], check=True)
    return ctypes.CDLL(so)


def run(lib, data, log_n, batch, stride, direction, coset, lens=(), cws=(), nthreads=8, use_tables=1, variant=0):
    buf = np.ascontiguousarray(data, dtype=np.uint64).copy()
    L = (ctypes.c_uint32 * 4)(*(list(lens) + [0] * (4 - len(lens))))
    W = (ctypes.c_uint32 * 4)(*(list(cws) + [0] * (4 - len(cws))))
    rc = lib.host_ntt_variant(buf.ctypes.data_as(ctypes.c_void_p), ctypes.c_uint32(log_n), ctypes.c_uint32(batch),
                              ctypes.c_uint64(stride), ctypes.c_int(direction), ctypes.c_int(coset), ctypes.c_uint32(len(lens)),
                              L, W, ctypes.c_uint32(nthreads), ctypes.c_int(use_tables), ctypes.c_int(variant))
    assert rc == 0
    return buf


SHAPES = [
    (1, (1,), (0,)), (3, (3,), (0,)), (6, (6,), (0,)),
    (6, (3, 3), (0, 0)), (6, (3, 3), (2, 2)), (6, (2, 4), (1, 2)), (7, (4, 3), (3, 1)),
    (6, (2, 2, 2), (0, 0, 0)), (6, (2, 2, 2), (1, 2, 2)), (8, (3, 2, 3), (2, 3, 3)), (9, (3, 3, 3), (2, 1, 3)),
    (8, (2, 2, 2, 2), (1, 2, 1, 2)), (9, (3, 2, 2, 2), (2, 1, 2, 3)),
]


@pytest.mark.parametrize("log_n,lens,cws", SHAPES)
@pytest.mark.parametrize("direction,coset", [(0, 0), (1, 0), (0, 1), (1, 1)])
def test_pass_decomposition_matches_oracle(hostlib, log_n, lens, cws, direction, coset):
    n = 1 << log_n
    rng = O.SplitMix64(100 * log_n + 10 * len(lens) + 2 * direction + coset)
    batch, stride = 2, n + 3
    vals = [rng.below(O.R_MOD, 253) for _ in range(stride * (batch - 1) + n)]
    data = H.ints_to_limbs(vals, 4)
    got = run(hostlib, data, log_n, batch, stride, direction, coset, lens, cws, nthreads=5)
    got2 = run(hostlib, data, log_n, batch, stride, direction, coset, lens, cws, nthreads=7, use_tables=0)
    assert np.array_equal(got, got2)          # per-pass twiddle tables == two-level powers
    want = data.copy()
    for b in range(batch):
        want[b * stride:b * stride + n] = C.ntt(data[b * stride:b * stride + n], log_n, direction=direction, coset=coset)
    assert np.array_equal(got, want)          # includes: padding between polynomials untouched
    # bulk-copy (TMA) variant: interleaved tile filled run by run, padded last-pass layout, half-swapped accesses
    got3 = run(hostlib, data, log_n, batch, stride, direction, coset, lens, cws, nthreads=6, variant=1)
    assert np.array_equal(got3, want)


@pytest.mark.parametrize("log_n", [4, 10, 12, 13, 14])
def test_default_plan(hostlib, log_n):
    n = 1 << log_n
    rng = np.random.default_rng(log_n)
    data = rng.integers(0, 1 << 62, size=(n, 4), dtype=np.uint64)
    data[:, 3] &= (1 << 59) - 1                 # < 2^251 < r : valid Montgomery representatives
    got = run(hostlib, data, log_n, 1, n, 0, 0, nthreads=64)
    assert np.array_equal(got, C.ntt(data, log_n))
    back = run(hostlib, got, log_n, 1, n, 1, 0, nthreads=64)
    assert np.array_equal(back, data)


def test_planner_invariants(hostlib):
    for log_n in range(1, 29):
        npasses = ctypes.c_uint32()
        L = (ctypes.c_uint32 * 4)()
        W = (ctypes.c_uint32 * 4)()
        assert hostlib.host_ntt_plan(ctypes.c_uint32(log_n), ctypes.byref(npasses), L, W) == 0
        k = npasses.value
        assert 1 <= k <= 4 and sum(L[i] for i in range(k)) == log_n
        before = 0
        for i in range(k):
            assert 1 <= L[i] <= 12 and L[i] + W[i] <= 12
            if k == 1:
                assert W[i] == 0
            elif i + 1 < k:
                assert W[i] <= log_n - before - L[i]
            else:
                assert W[i] <= L[0]
            before += L[i]


WC_SHAPES = [
    (10, (8, 2), (2, 2)),              # warp-column strided pass + generic last pass
    (10, (2, 8), (2, 2)),              # generic strided pass + warp-column LAST pass (digit reversal, t-fastest fill)
    (16, (8, 8), (2, 2)),              # both passes warp-column (the 2^16 plan)
    (12, (2, 8, 2), (2, 2, 2)),        # warp-column middle pass
]


@pytest.mark.parametrize("log_n,lens,cws", WC_SHAPES)
@pytest.mark.parametrize("direction,coset", [(0, 0), (1, 0), (0, 1), (1, 1)])
def test_warp_column_passes_match_oracle(hostlib, log_n, lens, cws, direction, coset):
    """ntt_pass_wc_kernel's phases (one warp per 2^8 column, 8 elements per lane, three in-register rounds, swizzled
    shared layout) emulated lane by lane: bit-exact against the oracle for all four transform kinds, first / middle /
    last position of the pass, padded batch stride"""
    n = 1 << log_n
    rng = np.random.default_rng(7000 + log_n + 2 * direction + coset)
    batch, stride = 2, n + 3
    data = H.random_fr_mont_np(rng, (stride * (batch - 1) + n,))
    got = run(hostlib, data, log_n, batch, stride, direction, coset, lens, cws, nthreads=16, variant=2)
    got_shaped = run(hostlib, data, log_n, batch, stride, direction, coset, lens, cws, nthreads=16, variant=3)
    want = data.copy()
    for b in range(batch):
        want[b * stride:b * stride + n] = C.ntt(data[b * stride:b * stride + n], log_n, direction=direction, coset=coset)
    assert np.array_equal(got, want)
    assert np.array_equal(got_shaped, want)      # ntt_pass_shaped_kernel<8, 2>'s phases (compile-time shape)


@pytest.mark.parametrize("log_n,lens,cws", [(10, (7, 3), (3, 3)), (10, (3, 7), (3, 3)), (14, (7, 7), (3, 3)), (13, (3, 7, 3), (3, 3, 3)),
                                            (12, (6, 6), (4, 4)), (13, (7, 6), (3, 4)), (10, (4, 6), (4, 4)),
                                            (11, (9, 2), (2, 2)), (11, (2, 9), (2, 2)), (11, (8, 3), (3, 3)), (11, (3, 8), (3, 3)),
                                            (17, (9, 8), (2, 3))])
@pytest.mark.parametrize("direction,coset", [(0, 0), (1, 1), (0, 1), (1, 0)])
def test_shape_specialised_7_3_and_6_4_match_oracle(hostlib, log_n, lens, cws, direction, coset):
    """ntt_pass_shaped_kernel<7, 3> (odd length: three stage pairs and one single stage) and <6, 4> (16-column tiles)
    emulated thread by thread, with the launcher's choice of the store mode (forward plain: modes 1 / 2; coset and
    inverse: mode 0)"""
    n = 1 << log_n
    rng = np.random.default_rng(8000 + log_n + 2 * direction + coset)
    data = H.random_fr_mont_np(rng, (n,))
    got = run(hostlib, data, log_n, 1, n, direction, coset, lens, cws, nthreads=16, variant=3)
    assert np.array_equal(got, C.ntt(data, log_n, direction=direction, coset=coset))


@pytest.mark.parametrize("log_len,rows,world,log_n_total", [(4, 8, 2, 7), (8, 16, 4, 12), (11, 32, 2, 16), (13, 16, 8, 17), (13, 48, 4, 19), (6, 5, 1, 9)])
@pytest.mark.parametrize("direction,twiddle", [(0, 1), (1, 1), (0, 0)])
def test_rows_transform_with_fused_exchange(hostlib, log_len, rows, world, log_n_total, direction, twiddle):
    """b200_ntt_rows_exchange_device emulated on the host: the last pass of a batch of row transforms stores every output
    (row, col) -- twiddled by w_N^((row_base + row) * col) -- into the transposed slab of the rank that owns the column.
    Checked against the oracle's row transforms + big-int twiddles + an explicit transposition (single pass, two passes,
    odd row counts that shrink the row tile, world = 1)."""
    n = 1 << log_len
    rank, row_base = world - 1, 3 * rows
    rng = np.random.default_rng(9000 + log_len + rows + direction)
    data = H.random_fr_mont_np(rng, (rows * n,))
    c_local, R = n // world, rows * world
    dsts = [np.zeros((c_local * R, 4), dtype=np.uint64) for _ in range(world)]
    ptrs = (ctypes.c_void_p * world)(*[d.ctypes.data_as(ctypes.c_void_p).value for d in dsts])
    rc = hostlib.host_ntt_rows_exchange(data.ctypes.data_as(ctypes.c_void_p), ctypes.c_uint64(rows), ctypes.c_uint32(log_len),
                                        ctypes.c_uint32(world), ctypes.c_uint32(rank), ctypes.c_uint32(log_n_total), ctypes.c_int(direction),
                                        ctypes.c_int(twiddle), ctypes.c_uint64(row_base), ptrs, ctypes.c_uint32(24))
    assert rc == 0
    dom = O.EvaluationDomain(1 << log_n_total)
    w = dom.group_gen_inv if direction else dom.group_gen
    N = 1 << log_n_total
    for r in range(rows):
        out = C.ntt(data[r * n:(r + 1) * n], log_len, direction=direction)
        vals = H.fr_from_mont_array(out)
        if twiddle:
            vals = [v * pow(w, ((row_base + r) * col) % N, O.R_MOD) % O.R_MOD for col, v in enumerate(vals)]
        want = H.fr_mont_array(vals)
        for col in range(0, n, max(1, n // 64)):            # a sample of the columns, every destination rank
            d, cl = col // c_local, col % c_local
            assert np.array_equal(dsts[d][cl * R + rank * rows + r], want[col]), (r, col)
    # every slot written by this rank is its own: rows of other ranks stay zero
    for d in range(world):
        view = dsts[d].reshape(c_local, world, rows, 4)
        for other in range(world):
            if other != rank:
                assert not view[:, other].any()
