"""b200_init / b200_shutdown: the library is bound to ONE device per process; re-binding to another ordinal
is refused until b200_shutdown; after a shutdown everything cached on the device (tables, resident bases, per-thread
streams, staging slots, the generator constants) is rebuilt on the next use."""
import numpy as np
import pytest

from oracle import c_oracle as C
from tests import helpers as H

pytestmark = pytest.mark.gpu


def test_reinit_and_shutdown_cycle():
    import torch
    import snarkos_b200 as S
    from snarkos_b200 import _lib
    L = _lib.lib()
    S.init(0)
    S.init(0)                                               # idempotent
    S.init(-1)                                              # "whatever is bound"
    e = L.b200_init(torch.cuda.device_count())              # another ordinal while bound: refused, state untouched
    assert e.code == -1
    n, log_n = 1 << 11, 12
    hb = S.synthetic_bases(n, seed=5).cpu().numpy()
    sc = H.random_scalars_np(np.random.default_rng(1), n)
    x = H.random_fr_mont_np(np.random.default_rng(2), (1 << log_n,))
    want_msm = C.g1_to_affine(C.msm(hb, sc)).tobytes()
    want_ntt = C.ntt(x, log_n)
    rb = S.ResidentBases(hb)
    assert C.g1_to_affine(rb.msm(sc)).tobytes() == want_msm
    for cycle in range(2):
        torch.cuda.synchronize()
        L.b200_shutdown()
        # handles die with the shutdown
        with pytest.raises(S.B200Error):
            rb.msm(sc)
        S.init(0)
        assert C.g1_to_affine(S.VariableBase.msm(hb, sc)).tobytes() == want_msm          # fresh per-thread stream
        assert np.array_equal(S.EvaluationDomain(1 << log_n).fft_in_place(x), want_ntt)   # tables rebuilt
        hb2 = S.synthetic_bases(n, seed=5).cpu().numpy()                                  # generator constants re-uploaded
        assert np.array_equal(hb2, hb)
        rb = S.ResidentBases(hb)
        assert C.g1_to_affine(rb.msm(sc)).tobytes() == want_msm
    rb.release()


def test_shutdown_after_queue_and_foreign_streams():
    """Scratch is cached per stream and a cache can outlive its stream: the queue's dispatcher stream is destroyed by
    b200_shutdown before the caches are emptied, and a caller may destroy a stream of its own.  Neither may be touched
    when the cached blocks are handed back (a stream-ordered free on a destroyed stream crashed inside the driver:
    tools/queue_bench.cpp, round 2); b200_release_scratch does the same outside a shutdown."""
    import ctypes
    import torch
    import snarkos_b200 as S
    from snarkos_b200 import _lib
    L = _lib.lib()
    S.init(0)
    n = 64
    hb = S.synthetic_bases(n, seed=9).cpu().numpy()
    sc = H.random_scalars_np(np.random.default_rng(3), n)
    want = C.g1_to_affine(C.msm(hb, sc)).tobytes()
    for cycle in range(2):
        # through the coalescing queue: scratch cached on the dispatcher's stream
        tk = ctypes.c_uint64(0)
        out = np.zeros(144, dtype=np.uint8)
        _lib.check(L.b200_msm_submit(hb.ctypes.data_as(ctypes.c_void_p), n, sc.ctypes.data_as(ctypes.c_void_p), 104, ctypes.byref(tk)))
        _lib.check(L.b200_msm_wait(tk.value, out.ctypes.data_as(ctypes.c_void_p)))
        assert C.g1_to_affine(out).tobytes() == want
        # on a caller's stream that is destroyed right afterwards: its cache stays behind
        st = torch.cuda.Stream()
        with torch.cuda.stream(st):
            dev = S.VariableBase.msm(torch.from_numpy(hb).cuda(), torch.from_numpy(sc.view(np.int64)).cuda())
        st.synchronize()
        assert C.g1_to_affine(dev.cpu().numpy()).tobytes() == want
        del st
        if cycle == 0:
            S.release_scratch()
            assert C.g1_to_affine(S.VariableBase.msm(hb, sc)).tobytes() == want
        torch.cuda.synchronize()
        L.b200_shutdown()
        S.init(0)
        assert C.g1_to_affine(S.VariableBase.msm(hb, sc)).tobytes() == want
