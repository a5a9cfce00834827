"""Re-entrancy of the C ABI: snarkOS calls the path from many rayon / tokio threads at once
(/root/reference/cli/src/commands/start.rs:616-641).  Eight host threads issue MSMs and NTTs concurrently through the
host-buffer entry points (ctypes drops the GIL for the duration of each call); every result must equal the
single-threaded one."""
import threading

import numpy as np
import pytest

from oracle import c_oracle as C
from tests import helpers as H

pytestmark = pytest.mark.gpu


def test_concurrent_callers():
    import torch
    import snarkos_b200 as S
    n, log_n = 1 << 11, 13
    bases = S.synthetic_bases(n, seed=17)
    torch.cuda.synchronize()
    hb = bases.cpu().numpy()
    rng = np.random.default_rng(0)
    jobs = []
    for t in range(8):
        sc = H.random_scalars_np(rng, n)
        x = H.random_fr_mont_np(rng, (1 << log_n,))
        jobs.append((sc, x, H.jac_bytes_to_affine(C.msm(hb, sc)), C.ntt(x, log_n, coset=1)))
    errors = []

    def work(sc, x, want_msm, want_ntt):
        try:
            d = S.EvaluationDomain(1 << log_n)
            for _ in range(4):
                assert H.jac_bytes_to_affine(S.VariableBase.msm(hb, sc)) == want_msm
                assert np.array_equal(d.coset_fft_in_place(x), want_ntt)
        except Exception as e:  # pragma: no cover
            errors.append(repr(e))

    threads = [threading.Thread(target=work, args=j) for j in jobs]
    for t in threads:
        t.start()
    for t in threads:
        t.join()
    assert not errors, errors


def test_short_lived_threads_do_not_pile_up_streams():
    """tokio's blocking pool creates and retires threads: a thread hands its CUDA streams back when it exits and the next
    thread takes them over, so the number of streams follows the number of CONCURRENT callers, not of threads ever seen
    (thousands of streams slow every stream-ordered allocation: tools/queue_bench.cpp, 3 ms -> 66 ms per batch call)"""
    import torch
    import snarkos_b200 as S
    n = 64
    bases = S.synthetic_bases(n, seed=3)
    torch.cuda.synchronize()
    hb = bases.cpu().numpy()
    sc = H.random_scalars_np(np.random.default_rng(1), n)
    want = H.jac_bytes_to_affine(C.msm(hb, sc))
    errors = []

    def work():
        try:
            assert H.jac_bytes_to_affine(S.VariableBase.msm(hb, sc)) == want
        except Exception as e:  # pragma: no cover
            errors.append(repr(e))

    def round_of(k):
        ts = [threading.Thread(target=work) for _ in range(k)]
        for t in ts:
            t.start()
        for t in ts:
            t.join()

    round_of(8)
    before = S.counter("streams_created")
    for _ in range(10):
        round_of(8)
    assert not errors, errors
    assert S.counter("streams_created") <= before + 8, (before, S.counter("streams_created"))
