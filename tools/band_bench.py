"""Burst timing of analysis / synthesis for several n_band at 64 x 2^20 samples (tensor-core Hankel kernels vs the direct form)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import pqmf_b200 as pq
B, T = 64, 1 << 20
def timeit(fn, n=5, inner=4):
    for _ in range(2): fn()
    torch.cuda.synchronize()
    best = 1e9
    for _ in range(n):
        fn(); e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(inner): fn()
        e1.record(); torch.cuda.synchronize(); best = min(best, e0.elapsed_time(e1) / inner)
    return best
x = (0.5 * torch.randn(B, 1, T, device="cuda")).clamp_(-1, 1)
for m in (4, 8, 16, 32, 64):
    for exact in (False, True):
        if exact and m not in (8, 32): continue
        mod = pq.PQMF(100, m, exact=exact).cuda()
        y = mod(x)
        ta = timeit(lambda: mod(x)); ts = timeit(lambda: mod.inverse(y))
        ns = B * T
        print(f"n_band {m:2d} exact={exact!s:5}: analysis {ta:.3f} ms ({ns/ta*1e-6:6.1f} Gs/s)  synthesis {ts:.3f} ms ({ns/ts*1e-6:6.1f} Gs/s)  round trip {16*ns/(ta+ts)*1e-6/6552.6*100:5.1f}% of HBM peak")
