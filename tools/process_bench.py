"""forward() + inverse() vs process() (one op, two whole-batch launches) vs reconstruct() (sub-bands through an L2-sized scratch buffer,
row chunk by row chunk: half the DRAM traffic, more launches) at the bench shape: burst and sustained (power-capped)."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import pqmf_b200 as pq
B, T = 64, 1 << 20
mod = pq.PQMF(100, 16).cuda()
x = (0.5 * torch.randn(B, 1, T, device="cuda")).clamp_(-1, 1)
def two():
    y = mod(x); return mod.inverse(y)
def fused():
    return mod.process(x)[0]
def recon():
    return mod.reconstruct(x)
def sustained(fn, secs=2.0):
    for _ in range(20): fn()
    torch.cuda.synchronize(); t0 = time.perf_counter(); n = 0
    while time.perf_counter() - t0 < secs:
        for _ in range(50): fn()
        torch.cuda.synchronize(); n += 50
    return (time.perf_counter() - t0) / n * 1e3
def burst(fn):
    for _ in range(3): fn()
    torch.cuda.synchronize(); best = 1e9
    for _ in range(8):
        fn(); e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(4): fn()
        e1.record(); torch.cuda.synchronize(); best = min(best, e0.elapsed_time(e1) / 4)
    return best
assert torch.equal(two(), fused()) and torch.equal(two(), recon())
for rep in range(2):
    a, b, c = sustained(two), sustained(fused), sustained(recon)
    print(f"sustained: forward+inverse {a:.4f} ms ({B*T/a*1e-6:.1f} Gs/s)   process {b:.4f} ms ({B*T/b*1e-6:.1f} Gs/s)   reconstruct {c:.4f} ms ({B*T/c*1e-6:.1f} Gs/s, {(a/c-1)*100:+.1f}%)")
a, b, c = burst(two), burst(fused), burst(recon)
print(f"burst:     forward+inverse {a:.4f} ms ({B*T/a*1e-6:.1f} Gs/s)   process {b:.4f} ms ({B*T/b*1e-6:.1f} Gs/s)   reconstruct {c:.4f} ms ({B*T/c*1e-6:.1f} Gs/s, {(a/c-1)*100:+.1f}%)")
