"""Small workload that touches every kernel family and entry point once (written as a compute-sanitizer target; the tool is closed on this
pool, so it serves as a crash / launch-error smoke run: every call is followed by a synchronize at the end and must not raise)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import pqmf_b200 as pq

torch.manual_seed(0)
dev = "cuda"
def noise(*shape):
    return (0.5 * torch.randn(*shape, device=dev)).clamp_(-1, 1)

# offline: Hankel pairs (>= 96 tiles), fold / Hankel-16 (small), direct form (other n_band, fp32), split bank (n_band 64)
for m, b, t in ((16, 24, 32768), (16, 2, 4096), (8, 24, 32768), (32, 12, 65536), (64, 24, 32768), (4, 25, 32768), (12, 2, 1200)):
    mod = (pq.PQMF(100, m, polyphase=(m != 12))).to(dev)
    x = noise(b, 1, t)
    y = mod(x); o = mod.inverse(y)
    if m != 12:
        mod.reconstruct(x); mod.process(x)
pq.PQMF(100, 16, exact=True).to(dev)(noise(2, 1, 4096))
pq.PQMF(100, 16, fp32=True).to(dev).inverse(noise(2, 16, 256))
# PCM edge, band hand-off
mod = pq.CachedPQMF(100, 16).to(dev)
pcm = torch.randint(-32768, 32767, (12, 32768, 2), device=dev, dtype=torch.int32).to(torch.int16)
y = mod.forward_pcm16(pcm); mod.inverse_pcm16(y); mod.forward_pcm16(pcm, True); mod.forward_pcm16(pcm[:1, :4096]); mod.inverse_pcm16(y[:1, :, :256].contiguous())
bands = [noise(1, 500 + 3 * k) for k in range(16)]
tail = noise(16, 32); full = torch.hann_window(64, device=dev)
mod.inverse_bands(bands, 512, tail, full[:32].unsqueeze(0), full[32:].unsqueeze(0))
# streaming: many streams (Hankel stream kernels, n_band 8 / 16 / 32), few streams (fold / direct), graph replay
for m, s in ((16, 300), (8, 300), (32, 300), (16, 3), (8, 3)):
    c = pq.CachedPQMF(100, m).to(dev)
    for _ in range(3):
        c.process_stream(noise(s, 1, 2048))
g = pq.StreamGraph(pq.CachedPQMF(100, 16).to(dev), 1, 512)
for _ in range(3):
    g.step(noise(1, 1, 512))
# autograd
xg = noise(24, 1, 32768).requires_grad_(True)
mod.inverse(mod(xg)).square().mean().backward()
torch.cuda.synchronize()
print("all kernel families ran:", pq.launch_count(), "launches")
