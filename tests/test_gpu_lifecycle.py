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
