#!/usr/bin/env python3
"""Do a batched row NTT (compute-bound) and the exchange/transpose kernel (memory-bound) overlap on one GPU?"""
import ctypes, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import snarkos_b200 as S
from snarkos_b200 import dist as D
S.init(0)
ops = D.CudaOps()
rows, cols = 4096, 8192                      # 2^25 elements = 1 GiB
a = torch.randint(0, 1 << 59, (rows * cols, 4), dtype=torch.int64, device="cuda")
b = torch.randint(0, 1 << 59, (rows * cols, 4), dtype=torch.int64, device="cuda")
c = torch.empty_like(b)
dst = (ctypes.c_void_p * 1)(c.data_ptr())
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream(priority=-1)
def ntt():
    ops.ntt_rows(a.view(rows, cols, 4), 13, 0)
def xchg():
    D.exchange_transpose(b, dst, 1, 0, rows, cols, 26, 0, TW, 0)
TW = os.environ.get("TW", "0") == "1"
def timed(fn, reps=3):
    fn(); torch.cuda.synchronize()
    best = 1e9
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize(); e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    return best
def both():
    cur = torch.cuda.current_stream()
    ev = torch.cuda.Event(); ev.record(cur)
    with torch.cuda.stream(s1):
        s1.wait_event(ev); ntt(); e1 = torch.cuda.Event(); e1.record(s1)
    with torch.cuda.stream(s2):
        s2.wait_event(ev); xchg(); e2 = torch.cuda.Event(); e2.record(s2)
    cur.wait_event(e1); cur.wait_event(e2)
print("ntt alone", timed(ntt), "xchg alone", timed(xchg), "both", timed(both), flush=True)
# timeline of one concurrent run: when does each stream's work start / end relative to the fork?
torch.cuda.synchronize()
cur = torch.cuda.current_stream()
T = lambda: torch.cuda.Event(enable_timing=True)
f, a0, a1, b0, b1 = T(), T(), T(), T(), T()
f.record(cur)
with torch.cuda.stream(s1):
    s1.wait_event(f); a0.record(s1); ntt(); a1.record(s1)
with torch.cuda.stream(s2):
    s2.wait_event(f); b0.record(s2); xchg(); b1.record(s2)
torch.cuda.synchronize()
print("ntt  [%.3f, %.3f] ms   xchg [%.3f, %.3f] ms" % (f.elapsed_time(a0), f.elapsed_time(a1), f.elapsed_time(b0), f.elapsed_time(b1)), flush=True)
