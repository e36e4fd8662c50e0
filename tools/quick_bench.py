"""Quick device-side timing of the two directions (CUDA events, L2-exceeding inputs).  Development aid; bench.py is the contract."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import pqmf_b200 as pq

def timeit(fn, n=10, warm=3):
    for _ in range(warm): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ts = []
    inner = 6   # back-to-back launches per timed interval, so host launch latency hides behind the previous kernel
    for _ in range(n):
        fn(); e0.record()
        for _ in range(inner): fn()
        e1.record(); torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1) / inner)
    ts.sort()
    return ts[0], ts[len(ts)//2]

def main():
    B, T = 64, 1 << 20
    if len(sys.argv) > 2: B, T = int(sys.argv[1]), int(sys.argv[2])
    x = (0.5 * torch.randn(B, 1, T, device="cuda")).clamp_(-1, 1)
    for exact in (False, True):
        mod = pq.PQMF(100, 16, exact=exact).cuda()
        y = mod(x)
        n = 10 if not exact else 3
        ta, tam = timeit(lambda: mod(x), n)
        ts, tsm = timeit(lambda: mod.inverse(y), n)
        ns = B * T
        print(f"exact={exact}: analysis {ta:.3f} ms ({ns/ta*1e-6:.1f} Gs/s, {8*ns/ta*1e-6:.0f} GB/s)  synthesis {ts:.3f} ms ({ns/ts*1e-6:.1f} Gs/s, {8*ns/ts*1e-6:.0f} GB/s)"
              f"  round trip {ns/(ta+ts)*1e-6:.1f} Gs/s = {16*ns/(ta+ts)*1e-6/6552.6*100:.1f}% of 6552.6 GB/s")
    # copy baseline
    a = torch.empty(B * T, device="cuda"); b = torch.empty_like(a)
    tc, _ = timeit(lambda: b.copy_(a), 10)
    print(f"torch copy_ {tc:.3f} ms -> {8*B*T/tc*1e-6:.0f} GB/s")

if __name__ == "__main__":
    main()
