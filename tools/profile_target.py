"""Small fixed workload for ncu captures: a few launches of each fast kernel on [16, 1, 2^20] (67 MB > half of L2... the
per-launch numbers under ncu are cold-cache and serialised anyway)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import pqmf_b200 as pq

B, T = (int(sys.argv[1]), int(sys.argv[2])) if len(sys.argv) > 2 else (32, 1 << 20)
torch.manual_seed(0)
x = (0.5 * torch.randn(B, 1, T, device="cuda")).clamp_(-1, 1)
mod = pq.PQMF(100, 16).cuda()
for _ in range(3):
    y = mod(x)
    out = mod.inverse(y)
torch.cuda.synchronize()
print("ok", float(y.abs().max()), float((out - x).abs().max()))
