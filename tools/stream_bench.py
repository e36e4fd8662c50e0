"""Streaming (cached) mode throughput: config 3 = 4096 streams x 2048-sample blocks, state carried between blocks."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import pqmf_b200 as pq
B, T = (int(sys.argv[1]), int(sys.argv[2])) if len(sys.argv) > 2 else (4096, 2048)
for exact in (False, True):
    mod = pq.CachedPQMF(100, 16, exact=exact).cuda()
    x = (0.5 * torch.randn(B, 1, T, device="cuda")).clamp_(-1, 1)
    mod.reset_stream()
    for _ in range(4):
        y = mod.forward_stream(x); o = mod.inverse_stream(y)
    torch.cuda.synchronize()
    e = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
    n = 50
    ta = ts = 0.0
    for _ in range(n):
        e[0].record(); y = mod.forward_stream(x); e[1].record(); o = mod.inverse_stream(y); e[2].record()
        torch.cuda.synchronize(); ta += e[0].elapsed_time(e[1]); ts += e[1].elapsed_time(e[2])
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        y = mod.forward_stream(x); o = mod.inverse_stream(y)
    e1.record(); torch.cuda.synchronize()
    tt = e0.elapsed_time(e1) / n
    print(f"exact={exact}: {B} streams x {T}: analysis {ta/n:.3f} ms synthesis {ts/n:.3f} ms (timed alone, incl. launch latency); "
          f"back to back {tt:.3f} ms per block step = {B*T/tt*1e-6:.1f} Gsamples/s = {16*B*T/tt*1e-6/6552.6*100:.1f}% of HBM roofline")
