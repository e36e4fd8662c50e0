"""Fixed cost vs per-tile cost of a streaming block step: time analysis / synthesis / both (each back to back, 200 launches per
interval) for a range of stream counts at one block length and fit t = a + b * tiles."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import pqmf_b200 as pq
T = int(sys.argv[1]) if len(sys.argv) > 1 else 2048
M = int(sys.argv[2]) if len(sys.argv) > 2 else 16
mod = pq.CachedPQMF(100, M).cuda()
rows = []
def timed(fn, n=200):
    for _ in range(10): fn()
    torch.cuda.synchronize()
    best = 1e9
    for _ in range(5):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(n): fn()
        e1.record(); torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1) / n)
    return best
for B in (444, 888, 1776, 3552, 4096, 7104, 14208, 28416):
    x = (0.5 * torch.randn(B, 1, T, device="cuda")).clamp_(-1, 1)
    mod.reset_stream()
    y = mod.forward_stream(x); o = mod.inverse_stream(y)
    ta = timed(lambda: mod.forward_stream(x))
    ts = timed(lambda: mod.inverse_stream(y))
    tb = timed(lambda: mod.inverse_stream(mod.forward_stream(x)))
    gm = pq.CachedPQMF(100, M).cuda()
    g = pq.StreamGraph(gm, B, T)
    g.x.copy_(x)
    tg = timed(lambda: g.step())
    rows.append((B, ta, ts, tb, tg))
    print(f"streams {B:6d}: analysis {ta*1e3:7.2f} us  synthesis {ts*1e3:7.2f} us  both {tb*1e3:7.2f} us  graph step {tg*1e3:7.2f} us  "
          f"-> {B*T/tb*1e-6:7.1f} Gs/s = {20*B*T/tb*1e-6/6552.6:.3f} of the 20 B/sample roofline", flush=True)
import numpy as np
b = np.array([r[0] for r in rows], float)
for k, name in ((1, "analysis"), (2, "synthesis"), (3, "both")):
    t = np.array([r[k] for r in rows]) * 1e3
    A = np.stack([np.ones_like(b), b], 1)
    (a0, a1), *_ = np.linalg.lstsq(A[3:], t[3:], rcond=None)
    print(f"{name}: t ~ {a0:.2f} us + {a1*1000:.3f} us per 1000 streams (fit on streams >= 3552)")
