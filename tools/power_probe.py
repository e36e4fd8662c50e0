"""Board power and SM clock (NVML) during ~3 s of back-to-back analysis + synthesis at the bench shape."""
import sys, os, time, threading
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, pynvml
import pqmf_b200 as pq
from pqmf_b200 import _lib
pynvml.nvmlInit(); h = pynvml.nvmlDeviceGetHandleByIndex(0)
B, T = 64, 1 << 20
mod = pq.PQMF(100, 16).cuda()
x = (0.5 * torch.randn(B, 1, T, device="cuda")).clamp_(-1, 1)
y = mod(x); out = torch.empty_like(x)
samples = []
stop = threading.Event()
def poll():
    while not stop.is_set():
        samples.append((pynvml.nvmlDeviceGetPowerUsage(h) / 1e3, pynvml.nvmlDeviceGetClockInfo(h, pynvml.NVML_CLOCK_SM),
                        pynvml.nvmlDeviceGetCurrentClocksEventReasons(h)))
        time.sleep(0.02)
for name, extra in (("pair", 0), ("single", _lib.PQMF_FLAG_NO_PAIR), ("fold", _lib.PQMF_FLAG_FOLD)):
    fl = mod._flags | extra
    samples.clear(); stop.clear()
    th = threading.Thread(target=poll); th.start()
    torch.cuda.synchronize(); t0 = time.perf_counter(); n = 0
    while time.perf_counter() - t0 < 3.0:
        for _ in range(100):
            yy = torch.ops.pqmf_b200.analysis(x, mod.hk, mod._tables, T // 16, fl)
            oo = torch.ops.pqmf_b200.synthesis(y, mod.hk, mod._tables, 0, fl)
        torch.cuda.synchronize(); n += 100
    dt = time.perf_counter() - t0
    stop.set(); th.join()
    half = samples[len(samples) // 2:]
    pw = sorted(s[0] for s in half); ck = sorted(s[1] for s in half)
    reasons = 0
    for s in half: reasons |= s[2]
    print(f"{name:6s}: {dt/n*1e3:.4f} ms per round trip = {B*T*n/dt*1e-9:.1f} Gs/s | second half of the run: power median {pw[len(pw)//2]:.0f} W max {pw[-1]:.0f} W, "
          f"SM clock median {ck[len(ck)//2]} MHz min {ck[0]}, reasons mask {reasons:#x}")
    time.sleep(2.0)
