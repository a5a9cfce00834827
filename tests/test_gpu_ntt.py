"""GPU parity (bit-exact) of EvaluationDomain::{fft,ifft,coset_fft,coset_ifft}_in_place through the C ABI against
the oracle; golden vectors; size-independent properties at the BASELINE sizes (2^20 x 16, 2^24)."""
import numpy as np
import pytest

from oracle import bls12_377 as O
from oracle import c_oracle as C
from tests import helpers as H

pytestmark = pytest.mark.gpu
KAT = H.load_kat()
KINDS = [(0, 0), (1, 0), (0, 1), (1, 1)]


def dom(n):
    import snarkos_b200 as S
    return S.EvaluationDomain(n)


def run(d, data, direction, coset):
    f = {(0, 0): d.fft_in_place, (1, 0): d.ifft_in_place, (0, 1): d.coset_fft_in_place, (1, 1): d.coset_ifft_in_place}
    return f[(direction, coset)](data)


def test_golden_vectors():
    assert H.fr_from_mont_array(dom(4).fft_in_place(H.fr_mont_array([1, 2, 3, 4]))) == KAT["NTT_4_1234"]
    assert H.fr_from_mont_array(dom(4).coset_fft_in_place(H.fr_mont_array([1, 2, 3, 4]))) == KAT["cosetNTT_4_1234"]
    assert H.fr_from_mont_array(dom(8).fft_in_place(H.fr_mont_array(list(range(1, 9))))) == KAT["NTT_8_1to8"]


@pytest.mark.parametrize("log_n", [0, 1, 2, 3, 5, 8, 10, 11, 12, 13, 14, 16, 17, 19, 20])
@pytest.mark.parametrize("direction,coset", KINDS)
def test_matches_oracle(log_n, direction, coset):
    n = 1 << log_n
    rng = np.random.default_rng(1000 + log_n)
    batch = 3 if log_n <= 16 else 1
    data = H.random_fr_mont_np(rng, (batch, n))
    got = run(dom(n), data, direction, coset)
    want = np.stack([C.ntt(data[b], log_n, direction=direction, coset=coset) for b in range(batch)])
    assert np.array_equal(got, want)


def test_zero_padding_and_edges():
    d = dom(100)                                   # -> size 128
    assert d.size == 128 and d.log_size_of_group == 7
    rng = O.SplitMix64(3)
    x = O.random_fr(rng, 100)
    got = H.fr_from_mont_array(d.fft_in_place(H.fr_mont_array(x)))
    assert got == O.EvaluationDomain(100).fft(x)
    assert H.fr_from_mont_array(d.fft_in_place(H.fr_mont_array([1]))) == [1] * 128            # delta -> ones
    assert H.fr_from_mont_array(d.fft_in_place(H.fr_mont_array([1] * 128))) == [128] + [0] * 127
    ext = [0, 1, O.R_MOD - 1, O.R_MOD - 2] * 32
    assert H.fr_from_mont_array(d.coset_ifft_in_place(d.coset_fft_in_place(H.fr_mont_array(ext)))) == ext
    with pytest.raises(ValueError):
        d.fft_in_place(H.fr_mont_array([1] * 129))


def test_batch_stride_and_device_resident():
    import ctypes
    import torch
    import snarkos_b200 as S
    log_n, batch, n = 12, 5, 1 << 12
    stride = n + 64
    rng = np.random.default_rng(5)
    data = H.random_fr_mont_np(rng, (batch * stride,))
    want = data.copy()
    for b in range(batch):
        want[b * stride:b * stride + n] = C.ntt(data[b * stride:b * stride + n], log_n, direction=0, coset=1)
    t = torch.from_numpy(data.view(np.int64)).cuda()
    S._lib.check(S.lib().b200_ntt_fr_bls12_377_device(ctypes.c_void_p(t.data_ptr()), log_n, batch, stride, 0, 1,
                                                      ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)))
    torch.cuda.synchronize()
    assert np.array_equal(t.cpu().numpy().view(np.uint64), want)       # gaps between polynomials untouched
    # EvaluationDomain on a CUDA tensor, in place
    x = H.random_fr_mont_np(rng, (4, n))
    t = torch.from_numpy(x.view(np.int64)).cuda()
    r = S.EvaluationDomain(n).ifft_in_place(t)
    assert r.data_ptr() == t.data_ptr()
    torch.cuda.synchronize()
    assert np.array_equal(t.cpu().numpy().view(np.uint64), np.stack([C.ntt(x[b], log_n, direction=1) for b in range(4)]))


@pytest.mark.parametrize("log_n,batch", [(20, 16), (24, 1), (26, 1)])
def test_full_size_properties(log_n, batch):
    """BASELINE configs: properties that need no oracle run at full size (intt(ntt(x)) = x, coset round trip,
    linearity, ntt(delta_1) = powers of omega spot-checked), plus one polynomial checked against the oracle."""
    import torch
    import snarkos_b200 as S
    n = 1 << log_n
    d = S.EvaluationDomain(n)
    rng = np.random.default_rng(log_n)
    x = H.random_fr_mont_np(rng, (batch, n))
    dx = torch.from_numpy(x.view(np.int64)).cuda()
    orig = dx.clone()
    d.fft_in_place(dx)
    fwd = dx.clone()
    assert not torch.equal(fwd, orig)
    d.ifft_in_place(dx)
    assert torch.equal(dx, orig)
    d.coset_fft_in_place(dx)
    d.coset_ifft_in_place(dx)
    assert torch.equal(dx, orig)
    # one polynomial against the CPU oracle (multi-threaded C restatement); 2^26 relies on the properties alone
    if log_n <= 24:
        assert np.array_equal(fwd[0].cpu().numpy().view(np.uint64), C.ntt(x[0], log_n))
    # delta at index 1 -> omega^k: spot check a few k against Python big-int powers
    delta = np.zeros((n, 4), dtype=np.uint64)
    delta[1] = H.fr_mont_array([1])[0]
    out = d.fft_in_place(torch.from_numpy(delta.view(np.int64)).cuda()).cpu().numpy().view(np.uint64)
    w = O.EvaluationDomain(n).group_gen
    for k in (0, 1, 2, 12345, n // 2, n - 1):
        assert H.fr_from_mont_array(out[k:k + 1])[0] == pow(w, k, O.R_MOD)


def test_pageable_host_buffers_staged_copies(b200_opt):
    """Ordinary (pageable) caller memory, as a Rust Vec is: uploads and downloads above 8 MiB go through the library's
    pinned staging slots on worker threads (hostcopy.cu) -- several slices per worker, both directions; the result must
    equal the driver's own pageable path bit for bit, and the round trip must be the identity."""
    log_n, batch = 20, 5                                     # 160 MiB: 40 slices of 4 MiB over 8 workers
    n = 1 << log_n
    data = H.random_fr_mont_np(np.random.default_rng(77), (batch, n))
    d = dom(n)
    staged = d.fft_in_place(data.copy())
    b200_opt("staged_copies", 0)
    plain = d.fft_in_place(data.copy())
    b200_opt("staged_copies", 1)
    assert np.array_equal(staged, plain)
    assert np.array_equal(staged[3], C.ntt(data[3], log_n))
    assert np.array_equal(d.ifft_in_place(staged.copy()), data)


def test_host_batch_pipeline_matches_single_shot(b200_opt):
    """Host batches above 64 MiB are pipelined in groups over three streams (upload / transforms / download); ragged
    groups (7 polynomials), a padded batch stride and pinned memory: same bytes as the single-shot path, padding
    between the polynomials untouched."""
    import torch
    log_n, batch, pad = 19, 7, 1000
    n = 1 << log_n
    stride = n + pad
    rng = np.random.default_rng(5)
    flat = H.random_fr_mont_np(rng, (batch * stride,))
    import snarkos_b200 as S
    from snarkos_b200 import _lib
    import ctypes

    def run(buf):
        t = torch.from_numpy(buf.copy().view(np.int64)).pin_memory()
        _lib.check(_lib.lib().b200_ntt_fr_bls12_377(ctypes.c_void_p(t.data_ptr()), log_n, batch, stride, 0, 1))
        return t.numpy().view(np.uint64).copy()

    piped = run(flat)
    b200_opt("ntt_host_pipeline", 0)
    single = run(flat)
    assert np.array_equal(piped, single)
    for b in (0, 3, 6):
        assert np.array_equal(piped[b * stride:b * stride + n], C.ntt(flat[b * stride:b * stride + n], log_n, coset=1))
        assert np.array_equal(piped[b * stride + n:(b + 1) * stride], flat[b * stride + n:(b + 1) * stride])


@pytest.mark.parametrize("log_n,batch", [(1, 3), (5, 2), (10, 3), (12, 1), (13, 2), (16, 2), (20, 1)])
@pytest.mark.parametrize("tile_log", [9, 10, 11])
def test_bulk_copy_tma_variant_matches_oracle(b200_opt, log_n, batch, tile_log):
    """ntt_variant = 1: the tile arrives through cp.async.bulk (TMA, mbarrier completion) into an interleaved,
    double-buffered tile while the previous tile is transformed (persistent CTAs) -- all four transform kinds, padded
    batch stride, bit-exact against the oracle and against the default kernel"""
    b200_opt("ntt_variant", 1)
    b200_opt("ntt_tile_log", tile_log)
    n = 1 << log_n
    stride = n + 5
    rng = np.random.default_rng(1000 + log_n)
    flat = H.random_fr_mont_np(rng, (batch * stride,))
    import ctypes
    from snarkos_b200 import _lib
    for direction, coset in ((0, 0), (1, 0), (0, 1), (1, 1)):
        buf = flat.copy()
        _lib.check(_lib.lib().b200_ntt_fr_bls12_377(buf.ctypes.data_as(ctypes.c_void_p), log_n, batch, stride, direction, coset))
        for b in range(batch):
            want = C.ntt(flat[b * stride:b * stride + n], log_n, direction=direction, coset=coset)
            assert np.array_equal(buf[b * stride:b * stride + n], want), (direction, coset, b)
            assert np.array_equal(buf[b * stride + n:(b + 1) * stride], flat[b * stride + n:(b + 1) * stride])


@pytest.mark.parametrize("log_n,plan", [(10, "8,2"), (10, "2,8"), (12, "2,8,2"), (16, "8,8"), (18, "8,8,2"), (24, "8,8,8")])
def test_warp_column_kernel_matches_generic_and_oracle(b200_opt, log_n, plan):
    """ntt_pass_wc_kernel (one warp per 2^8 column, in-register radix-8 rounds) in first / middle / last position: same
    bytes as the default pass kernel for all four transform kinds, and the oracle up to 2^18 (ntt_variant = 4: the
    kernel is an opt-in variant, measured slower than the default -- see ntt.cu)"""
    import ctypes
    import torch
    from snarkos_b200 import _lib
    b200_opt("ntt_plan", plan)
    n = 1 << log_n
    batch = 2 if log_n <= 18 else 1
    stride = n + (5 if log_n <= 18 else 0)
    rng = np.random.default_rng(4000 + log_n)
    flat = H.random_fr_mont_np(rng, (batch * stride,))
    L = _lib.lib()
    for direction, coset in ((0, 0), (1, 0), (0, 1), (1, 1)):
        outs = []
        for variant in (4, 0):
            b200_opt("ntt_variant", variant)
            buf = flat.copy()
            _lib.check(L.b200_ntt_fr_bls12_377(buf.ctypes.data_as(ctypes.c_void_p), log_n, batch, stride, direction, coset))
            outs.append(buf)
        assert np.array_equal(outs[0], outs[1]), (direction, coset)
        if log_n <= 18:
            for b in range(batch):
                want = C.ntt(flat[b * stride:b * stride + n], log_n, direction=direction, coset=coset)
                assert np.array_equal(outs[0][b * stride:b * stride + n], want), (direction, coset, b)
                assert np.array_equal(outs[0][b * stride + n:(b + 1) * stride], flat[b * stride + n:(b + 1) * stride])
