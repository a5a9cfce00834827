"""BASELINE configs[3] shape: a validator verifies the transactions of a block on up to 512 threads
(/root/reference/cli/src/commands/start.rs:623-640, node/bft/ledger-service/src/ledger.rs:341-347) and each
verification issues small VariableBase::msm calls.  The coalescing queue (snarkos_b200/csrc/queue.cu) must return, for
every caller, exactly the point a stand-alone call returns -- whatever the batching happened to be."""
import ctypes
import threading

import numpy as np
import pytest

from oracle import bls12_377 as O
from oracle import c_oracle as C
from tests import helpers as H

pytestmark = pytest.mark.gpu


def _inputs(n_tx, per, seed):
    import torch
    import snarkos_b200 as S
    hb = S.synthetic_bases(n_tx * per, seed=seed).cpu().numpy()
    torch.cuda.synchronize()
    sc = H.random_scalars_np(np.random.default_rng(seed), n_tx * per)
    return hb, sc


def test_submit_wait_256_threads_x_40_points():
    import snarkos_b200 as S
    from snarkos_b200 import _lib
    L = _lib.lib()
    n_tx, per = 256, 40
    hb, sc = _inputs(n_tx, per, 606)
    off = np.arange(n_tx + 1, dtype=np.uint64) * per
    want = S.msm_batch(hb, sc, off)                          # per-call results (checked against the oracle elsewhere)
    got = np.zeros((n_tx, 144), dtype=np.uint8)
    errors = []
    start = threading.Barrier(n_tx)

    def tx(m):
        try:
            pts = hb[m * per * 104:(m + 1) * per * 104]
            s = sc[m * per:(m + 1) * per]
            ticket = ctypes.c_uint64(0)
            start.wait()
            _lib.check(L.b200_msm_submit(pts.ctypes.data_as(ctypes.c_void_p), per, s.ctypes.data_as(ctypes.c_void_p), 104,
                                         ctypes.byref(ticket)))
            _lib.check(L.b200_msm_wait(ticket.value, got[m].ctypes.data_as(ctypes.c_void_p)))
        except Exception as e:  # pragma: no cover
            errors.append(repr(e))

    b0, s0 = S.counter("queue_batches"), S.counter("queue_submits")
    threads = [threading.Thread(target=tx, args=(m,)) for m in range(n_tx)]
    for t in threads:
        t.start()
    for t in threads:
        t.join()
    assert not errors, errors[:3]
    assert S.counter("queue_submits") - s0 == n_tx
    assert 1 <= S.counter("queue_batches") - b0 <= n_tx
    for m in range(n_tx):
        assert C.g1_to_affine(got[m]).tobytes() == C.g1_to_affine(want[m]).tobytes(), m
    for m in (0, 17, 255):
        assert H.jac_bytes_to_affine(got[m]) == H.jac_bytes_to_affine(C.msm(hb[m * per * 104:(m + 1) * per * 104], sc[m * per:(m + 1) * per]))


def test_queue_behind_the_unchanged_entry_point(b200_opt):
    """with msm_queue_threshold set, b200_msm_g1_bls12_377 itself routes small calls through the queue: the Rust call
    sites do not change; ragged sizes, an empty MSM, infinity among the bases"""
    import snarkos_b200 as S
    b200_opt("msm_queue_threshold", 128)
    rng = O.SplitMix64(99)
    sizes = [1, 7, 33, 128, 2, 64, 100, 18, 129, 300]        # the last two are above the threshold: direct path
    jobs = []
    for k in sizes:
        pts = O.random_points(rng, k)
        sc = O.random_fr(rng, k)
        if k > 5:
            pts[3] = None
            sc[4] = 0
        jobs.append((H.bases_array(pts), H.scalars_array(sc), O.msm_naive(pts, sc) if k <= 128 else None, pts, sc))
    out = [None] * len(jobs)
    errors = []

    def work(i):
        try:
            out[i] = H.jac_bytes_to_affine(S.VariableBase.msm(jobs[i][0], jobs[i][1]))
        except Exception as e:  # pragma: no cover
            errors.append(repr(e))

    s0 = S.counter("queue_submits")
    threads = [threading.Thread(target=work, args=(i,)) for i in range(len(jobs))]
    for t in threads:
        t.start()
    for t in threads:
        t.join()
    assert not errors, errors
    assert S.counter("queue_submits") - s0 == sum(1 for k in sizes if k <= 128)
    for i, (b, s, want, pts, sc) in enumerate(jobs):
        if want is None:
            want = H.jac_bytes_to_affine(C.msm(b, s))
        assert out[i] == want, sizes[i]


def test_queue_errors():
    import snarkos_b200 as S
    from snarkos_b200 import _lib
    L = _lib.lib()
    out = np.zeros(144, dtype=np.uint8)
    e = L.b200_msm_wait(0xdeadbeef, out.ctypes.data_as(ctypes.c_void_p))
    assert e.code == -4                                     # B200_ERR_BAD_HANDLE
    t = ctypes.c_uint64(0)
    e = L.b200_msm_submit(None, 5, None, 104, ctypes.byref(t))
    assert e.code == -1
    e = L.b200_msm_submit(None, 0, None, 104, ctypes.byref(t))          # empty MSM: infinity
    assert e.code == 0
    _lib.check(L.b200_msm_wait(t.value, out.ctypes.data_as(ctypes.c_void_p)))
    assert H.jac_bytes_to_affine(out) is None
    e = L.b200_msm_wait(t.value, out.ctypes.data_as(ctypes.c_void_p))  # a ticket is waited on once
    assert e.code == -4
