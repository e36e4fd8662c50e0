"""Which stream layout moves host buffers through the kernels fastest?  Same work as pqmf_roundtrip_host_f32 (64 x 2^20 fp32 in and out,
16 MiB row chunks), driven from Python through the torch ops: (A) one stream per chunk slot (the C ABI's layout), (B) one stream per
STAGE (H2D / kernels / D2H) with events between them, (C) copies only, as in tools/pcie_probe.py."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import pqmf_b200 as pq
B, T = 64, 1 << 20
rows = int(sys.argv[1]) if len(sys.argv) > 1 else 4           # rows per chunk (4 = 16 MiB)
nslot = int(sys.argv[2]) if len(sys.argv) > 2 else 4
mod = pq.PQMF(100, 16).cuda()
hx = torch.randn(B, 1, T).clamp_(-1, 1).pin_memory(); ho = torch.empty(B, 1, T).pin_memory()
dx = [torch.empty(rows, 1, T, device="cuda") for _ in range(nslot)]
do = [torch.empty(rows, 1, T, device="cuda") for _ in range(nslot)]
slot_streams = [torch.cuda.Stream() for _ in range(nslot)]
s_in, s_k, s_out = torch.cuda.Stream(), torch.cuda.Stream(), torch.cuda.Stream()
def per_slot(kernels=True, tiny=False, copy_kernel=False, sleep=0):
    for i, r in enumerate(range(0, B, rows)):
        k = i % nslot
        with torch.cuda.stream(slot_streams[k]):
            dx[k].copy_(hx[r:r + rows], non_blocking=True)
            if sleep:           # one thread spinning for `sleep` cycles between the two copies: no HBM traffic, no SMs to speak of
                torch.cuda._sleep(sleep)
                ho[r:r + rows].copy_(dx[k], non_blocking=True)
            elif tiny:            # a kernel that does nothing between the two copies
                dx[k][0, 0, :1].add_(1.0)
                ho[r:r + rows].copy_(dx[k], non_blocking=True)
            elif copy_kernel:   # a plain device copy (HBM-bound, short) between the two copies
                do[k].copy_(dx[k])
                ho[r:r + rows].copy_(do[k], non_blocking=True)
            elif kernels:
                o = mod.inverse(mod(dx[k]))
                ho[r:r + rows].copy_(o, non_blocking=True)
            else:
                ho[r:r + rows].copy_(dx[k], non_blocking=True)
    torch.cuda.synchronize()
def per_stage():
    ev_in = [None] * nslot; ev_k = [None] * nslot; ev_out = [None] * nslot
    for i, r in enumerate(range(0, B, rows)):
        k = i % nslot
        with torch.cuda.stream(s_in):
            if ev_k[k] is not None: s_in.wait_event(ev_k[k])      # the kernels of the chunk that used this slot have read dx[k]
            dx[k].copy_(hx[r:r + rows], non_blocking=True)
            ev_in[k] = torch.cuda.Event(); ev_in[k].record(s_in)
        with torch.cuda.stream(s_k):
            s_k.wait_event(ev_in[k])
            if ev_out[k] is not None: s_k.wait_event(ev_out[k])   # do[k] has been copied out
            y = mod(dx[k])
            torch.ops.pqmf_b200.synthesis  # (same ops as per_slot)
            o = mod.inverse(y)
            do[k].copy_(o)
            ev_k[k] = torch.cuda.Event(); ev_k[k].record(s_k)
        with torch.cuda.stream(s_out):
            s_out.wait_event(ev_k[k])
            ho[r:r + rows].copy_(do[k], non_blocking=True)
            ev_out[k] = torch.cuda.Event(); ev_out[k].record(s_out)
    torch.cuda.synchronize()
def timed(fn, n=8):
    for _ in range(3): fn()
    best = 1e9
    for _ in range(n):
        t0 = time.perf_counter(); fn(); best = min(best, time.perf_counter() - t0)
    return best
for name, fn in (("A: stream per chunk slot", per_slot), ("B: stream per stage", per_stage), ("C: copies only, per slot", lambda: per_slot(False)),
                 ("D: copies + an empty kernel", lambda: per_slot(tiny=True)), ("E: copies + a device copy", lambda: per_slot(copy_kernel=True)),
                 ("F: copies + 100 us spin", lambda: per_slot(sleep=190000)), ("G: copies + 300 us spin", lambda: per_slot(sleep=570000))):
    dt = timed(fn)
    print(f"{name:28s} rows/chunk {rows} slots {nslot}: {dt*1e3:.2f} ms -> {B*T/dt*1e-9:.2f} Gsamples/s = {B*T*4/dt*1e-9:.1f} GB/s per direction", flush=True)

# D2H start offset sweep: spin for a fraction of a chunk's copy time between the H2D and the D2H of a chunk (copies only otherwise)
H_us = rows * T * 4 / 48e3          # chunk copy time at ~48 GB/s, in us
for frac in (0.0, 0.05, 0.1, 0.2, 0.35, 0.5, 0.65, 0.8, 0.9, 0.95, 1.0, 1.05, 1.1, 1.25, 1.5, 2.0):
    cyc = int(frac * H_us * 1900)
    dt = timed(lambda: per_slot(sleep=cyc) if cyc else per_slot(False), n=6)
    print(f"offset {frac:4.2f} x chunk copy time ({frac*H_us:6.0f} us): {dt*1e3:.2f} ms", flush=True)
