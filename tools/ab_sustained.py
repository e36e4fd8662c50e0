"""Sustained (3 s) round trips at the bench shape for the library in the current directory (used to A/B two builds on ONE box)."""
import sys, os, time
sys.path.insert(0, os.getcwd())
import torch
import pqmf_b200 as pq
B, T = 64, 1 << 20
mod = pq.PQMF(100, 16).cuda()
x = (0.5 * torch.randn(B, 1, T, device="cuda")).clamp_(-1, 1)
y = mod(x)
def burst(fn):
    for _ in range(3): fn()
    torch.cuda.synchronize(); best = 1e9
    for _ in range(8):
        fn(); e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(6): fn()
        e1.record(); torch.cuda.synchronize(); best = min(best, e0.elapsed_time(e1) / 6)
    return best
ba, bs = burst(lambda: mod(x)), burst(lambda: mod.inverse(y))
time.sleep(1.0)
torch.cuda.synchronize(); t0 = time.perf_counter(); n = 0
while time.perf_counter() - t0 < 3.0:
    for _ in range(100):
        yy = mod(x); oo = mod.inverse(yy)
    torch.cuda.synchronize(); n += 100
dt = time.perf_counter() - t0
print(f"{os.getcwd()}: burst analysis {ba:.4f} synthesis {bs:.4f} ms; sustained {dt/n*1e3:.4f} ms per round trip = {B*T*n/dt*1e-9:.1f} Gs/s")
