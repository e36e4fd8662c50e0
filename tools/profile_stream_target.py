"""Small fixed workload for ncu captures of the streaming kernels: config 3, a few block steps."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, pqmf_b200 as pq
mod = pq.CachedPQMF(100, 16).cuda()
x = (0.5 * torch.randn(4096, 1, 2048, device="cuda")).clamp_(-1, 1)
for _ in range(4):
    y = mod.forward_stream(x); o = mod.inverse_stream(y)
torch.cuda.synchronize(); print("ok")
