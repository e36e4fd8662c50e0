"""A/B on ONE box: Hankel-4 kernels with the bulk-copy staging ring (default) vs the register-prefetch variant (PQMF_FLAG_NO_PREFETCH),
burst (6 launches per interval, best of 10) and sustained (N back-to-back round trips).  Development aid; bench.py is the contract."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import pqmf_b200 as pq
from pqmf_b200 import _lib


def burst(fn, n=10, inner=6):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    best = 1e9
    for _ in range(n):
        fn()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(inner):
            fn()
        e1.record()
        torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1) / inner)
    return best


def main():
    b, t = 64, 1 << 20
    steps = int(sys.argv[1]) if len(sys.argv) > 1 else 300
    m = int(sys.argv[2]) if len(sys.argv) > 2 else 16
    x = (0.5 * torch.randn(b, 1, t, device="cuda")).clamp_(-1, 1)
    ns = b * t
    outs = {}
    for name, extra in (("l2pf", 0), ("no_l2pf", _lib.PQMF_FLAG_NO_PREFETCH), ("l2pf", 0), ("no_l2pf", _lib.PQMF_FLAG_NO_PREFETCH)):
        mod = pq.PQMF(100, m).cuda()
        mod._flags |= extra
        y = mod(x)
        o = mod.inverse(y)
        outs[name] = (y, o)
        ta, ts = burst(lambda: mod(x)), burst(lambda: mod.inverse(y))
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            mod.inverse(mod(x))
        e1.record()
        torch.cuda.synchronize()
        sus = e0.elapsed_time(e1) / steps
        print(f"{name:8s} n_band {m}: burst analysis {ta:.4f} ms synthesis {ts:.4f} ms -> {16 * ns / (ta + ts) * 1e-6 / 6552.6:.3f} of HBM peak | "
              f"sustained {steps} steps {sus:.4f} ms/step -> {ns / sus * 1e-6:.1f} Gs/s = {16 * ns / sus * 1e-6 / 6552.6:.3f}", flush=True)
    print("bit-identical:", torch.equal(outs["l2pf"][0], outs["no_l2pf"][0]), torch.equal(outs["l2pf"][1], outs["no_l2pf"][1]))


if __name__ == "__main__":
    main()
