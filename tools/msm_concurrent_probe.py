#!/usr/bin/env python3
"""Two large MSMs of a proving burst on two streams against one after the other: the gather-bound denominator pass of one
call can run under the multiplier-bound additions of the other."""
import os, sys, threading, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import snarkos_b200 as S
S.init(0)
log_n = int(os.environ.get("LOG_N", "24"))
n = 1 << log_n
bases = [S.synthetic_bases(n, seed=5 + k) for k in range(2)]
g = torch.Generator(device="cuda"); g.manual_seed(1)
scs = []
for k in range(2):
    sc = torch.randint(-(1 << 63), (1 << 63) - 1, (n, 4), dtype=torch.int64, device="cuda", generator=g)
    sc[:, 3] &= (1 << 60) - 1
    scs.append(sc)
rbs = [S.ResidentBases(b) for b in bases]            # GLV pairs resident, as a prover's SRS would be
streams = [torch.cuda.Stream() for _ in range(2)]
torch.cuda.synchronize()

def one(k, reps):
    with torch.cuda.stream(streams[k]):
        for _ in range(reps):
            rbs[k].msm(scs[k])
        streams[k].synchronize()

for k in range(2):
    one(k, 1)
t0 = time.perf_counter(); one(0, 3); one(1, 3); seq = (time.perf_counter() - t0) / 6
ths = [threading.Thread(target=one, args=(k, 3)) for k in range(2)]
t0 = time.perf_counter()
for t in ths: t.start()
for t in ths: t.join()
conc = (time.perf_counter() - t0) / 6
print(f"2^{log_n}: one after the other {seq * 1e3:.2f} ms per MSM, two streams at once {conc * 1e3:.2f} ms per MSM "
      f"({n / conc / 1e6:.1f} Mpoints/s aggregate, {seq / conc:.3f} x)")
