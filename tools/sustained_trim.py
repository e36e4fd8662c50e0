"""Sustained (power-capped) throughput of the Hankel-4 kernels as a function of the trimmed correction steps and of CTA pairing.
Accuracy is NOT checked here: this only asks how much time a correction K-step costs once the clocks have settled."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import pqmf_b200 as pq
from pqmf_b200 import _lib
B, T = 64, 1 << 20
mod = pq.PQMF(100, 16).cuda()
x = (0.5 * torch.randn(B, 1, T, device="cuda")).clamp_(-1, 1)
y = mod(x)
base = mod._flags & ~((7 << 17) | (7 << 20))
def run(fn, n=600):
    for _ in range(50): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n
for extra, name in ((0, "pair"), (_lib.PQMF_FLAG_NO_PAIR, "single"), (_lib.PQMF_FLAG_FOLD, "fold")):
    for trim in ((0, 2, 4, 6) if name != "fold" else (0,)):
        fl = base | (trim << 17) | (trim << 20) | extra
        ta = run(lambda: torch.ops.pqmf_b200.analysis(x, mod.hk, mod._tables, T // 16, fl))
        ts = run(lambda: torch.ops.pqmf_b200.synthesis(y, mod.hk, mod._tables, 0, fl))
        print(f"{name:6s} trim {trim}: analysis {ta:.4f} ms  synthesis {ts:.4f} ms  -> round trip {B*T/(ta+ts)*1e-6:.1f} Gs/s = {16*B*T/(ta+ts)*1e-6/6552.6*100:.1f}%")
