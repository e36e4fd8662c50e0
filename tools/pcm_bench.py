"""Device-side timing of the PCM instantiations against the fp32 kernels at the bench shape (development aid)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import pqmf_b200 as pq

def burst(fn, n=8, inner=6):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    best = 1e9
    for _ in range(n):
        fn()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(inner): fn()
        e1.record(); torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1) / inner)
    return best

b, t = 64, 1 << 20
mod = pq.PQMF(100, 16).cuda()
x = (0.5 * torch.randn(b, 1, t, device="cuda")).clamp_(-1, 1)
pcm1 = (x[:, 0] * 32767).round().to(torch.int16).reshape(b, t, 1).contiguous()
pcm2 = pcm1.reshape(b // 2, 2, t).transpose(1, 2).contiguous()   # 32 stereo clips
y = mod(x)
print(f"fp32      analysis {burst(lambda: mod(x)):.4f} ms   synthesis {burst(lambda: mod.inverse(y)):.4f} ms")
print(f"pcm mono  analysis {burst(lambda: mod.forward_pcm16(pcm1)):.4f} ms   synthesis {burst(lambda: mod.inverse_pcm16(y)):.4f} ms")
y2 = y.reshape(b // 2, 32, -1)
print(f"pcm stereo analysis {burst(lambda: mod.forward_pcm16(pcm2)):.4f} ms   synthesis {burst(lambda: mod.inverse_pcm16(y2)):.4f} ms   down-mix analysis {burst(lambda: mod.forward_pcm16(pcm2, True)):.4f} ms (32 rows)")
