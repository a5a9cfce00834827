#!/usr/bin/env python3
"""First-light probe for a B200 box: integer-pipe microbenchmarks (roofline denominators for the MSM) and quick
timings of NTT / MSM at a few sizes.  Writes gpurun_out/probe.json."""
import ctypes
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch

import snarkos_b200 as S


def microbench():
    L = S.lib()
    out = {}
    names = {0: "imad", 1: "imad_wide", 2: "imad_hi", 6: "iadd", 3: "fr_modmul", 4: "fq_modmul", 5: "xyzz_madd"}
    iters = {0: 4096, 1: 4096, 2: 4096, 6: 4096, 3: 512, 4: 256, 5: 64}
    for kind, name in names.items():
        ms = ctypes.c_float()
        ops = ctypes.c_double()
        best = None
        for _ in range(3):
            S._lib.check(L.b200_debug_microbench(kind, iters[kind], ctypes.byref(ms), ctypes.byref(ops)))
            rate = ops.value / (ms.value * 1e-3)
            best = rate if best is None else max(best, rate)
        out[name] = {"ops_per_s": best, "ms": ms.value, "ops": ops.value}
        print(f"{name:12s} {best / 1e9:12.2f} Gop/s", flush=True)
    return out


def time_it(fn, reps=5, warm=2):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    return min(ts), sorted(ts)[len(ts) // 2]


def main():
    torch.cuda.set_device(0)
    S.init(0)
    res = {"device": torch.cuda.get_device_name(0), "microbench": microbench(), "ntt": {}, "msm": {}}
    rng = np.random.default_rng(0)
    for log_n, batch in ((16, 16), (20, 16), (22, 1), (24, 1)):
        n = 1 << log_n
        x = rng.integers(0, 1 << 62, size=(batch, n, 4), dtype=np.uint64)
        x[..., 3] &= np.uint64((1 << 59) - 1)
        t = torch.from_numpy(x.view(np.int64)).cuda()
        d = S.EvaluationDomain(n)
        best, med = time_it(lambda: d.fft_in_place(t))
        res["ntt"][f"2^{log_n}x{batch}"] = {"ms_best": best, "ms_median": med, "gelem_per_s": batch * n / best / 1e6,
                                             "hbm_gbs_algorithmic": 64 * batch * n / best / 1e6}
        print("ntt", log_n, batch, res["ntt"][f"2^{log_n}x{batch}"], flush=True)
        del t
    for log_n in (16, 20, 22, 24):
        n = 1 << log_n
        bases = S.synthetic_bases(n, seed=3)
        sc = rng.integers(0, 1 << 62, size=(n, 4), dtype=np.uint64)
        sc[:, 3] &= np.uint64((1 << 59) - 1)
        dsc = torch.from_numpy(sc.view(np.int64)).cuda()
        best, med = time_it(lambda: S.VariableBase.msm(bases, dsc), reps=3, warm=1)
        res["msm"][f"2^{log_n}"] = {"ms_best": best, "ms_median": med, "mpoints_per_s": n / best / 1e3,
                                    "c": int(S.lib().b200_msm_window_bits(n))}
        print("msm", log_n, res["msm"][f"2^{log_n}"], flush=True)
        del bases, dsc
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    json.dump(res, open(os.path.join(ROOT, "gpurun_out", "probe.json"), "w"), indent=1)


if __name__ == "__main__":
    main()
