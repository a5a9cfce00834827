#!/usr/bin/env python3
"""The fused exchange kernel of the four-step NTT (ntt_exchange_transpose_kernel: transpose + twiddle + stores straight
into the destination ranks' slabs) launched from ONE process, for an ncu capture with NVLink counters: ncu cannot follow
a multi-rank job (it replays every kernel ~40 times, the peers would wait at their barriers), so rank 0's launch of a
WORLD-rank exchange is reproduced here with the other ranks' slabs allocated on the other visible GPUs and peer access
enabled -- the same peer stores over NVLink as in the torchrun run, one writer.  With one visible GPU all slabs are local."""
import ctypes, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import snarkos_b200 as S
from snarkos_b200 import dist as D

log_n = int(os.environ.get("LOG_N", "26"))
ngpu = torch.cuda.device_count()
world = int(os.environ.get("WORLD", str(max(ngpu, 2))))
torch.cuda.set_device(0)
S.init(0)
rt = ctypes.CDLL("libcudart.so.12")
for d in range(1, ngpu):
    rc = rt.cudaDeviceEnablePeerAccess(d, 0)
    print(f"peer access 0 -> {d}: rc={rc}", flush=True)
n = 1 << log_n
n1_log = log_n // 2
n2 = 1 << (log_n - n1_log)
n1 = 1 << n1_log
per = n // world
rc_rows = n1 // world                     # rows of the local slab [rc_rows x n2]
src = torch.randint(0, 1 << 59, (per, 4), dtype=torch.int64, device="cuda:0")
dsts = [torch.empty((per, 4), dtype=torch.int64, device=f"cuda:{r % ngpu}") for r in range(world)]
ptrs = (ctypes.c_void_p * world)(*[t.data_ptr() for t in dsts])
for twiddle in (False, True):
    for _ in range(3):
        D.exchange_transpose(src, ptrs, world, 0, rc_rows, n2, log_n, 0, twiddle, 0)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10):
        D.exchange_transpose(src, ptrs, world, 0, rc_rows, n2, log_n, 0, twiddle, 0)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 10
    remote = per * 32 * (world - (world + ngpu - 1) // ngpu) / world if ngpu > 1 else 0
    print(f"exchange 2^{log_n} world={world} gpus={ngpu} twiddle={int(twiddle)}: {ms:.3f} ms per launch, slab {per * 32 / 2**20:.0f} MiB, "
          f"{per * 32 / ms / 1e6:.0f} GB/s read+written, remote bytes per launch {remote / 2**20:.0f} MiB", flush=True)
