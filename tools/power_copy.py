"""Board power during ~3 s of back-to-back device copies (256 MiB read + 256 MiB written per copy) -- what HBM traffic alone costs."""
import time, threading, torch, pynvml
pynvml.nvmlInit(); h = pynvml.nvmlDeviceGetHandleByIndex(0)
a = torch.randn(64 << 20, device="cuda"); b = torch.empty_like(a)
samples = []; stop = threading.Event()
def poll():
    while not stop.is_set():
        samples.append((pynvml.nvmlDeviceGetPowerUsage(h) / 1e3, pynvml.nvmlDeviceGetClockInfo(h, pynvml.NVML_CLOCK_SM))); time.sleep(0.02)
th = threading.Thread(target=poll); th.start()
torch.cuda.synchronize(); t0 = time.perf_counter(); n = 0
while time.perf_counter() - t0 < 3.0:
    for _ in range(200): b.copy_(a)
    torch.cuda.synchronize(); n += 200
dt = time.perf_counter() - t0
stop.set(); th.join()
half = samples[len(samples) // 2:]
pw = sorted(s[0] for s in half); ck = sorted(s[1] for s in half)
print(f"copy: {2 * a.numel() * 4 * n / dt * 1e-9:.0f} GB/s sustained, power median {pw[len(pw)//2]:.0f} W max {pw[-1]:.0f} W, SM clock median {ck[len(ck)//2]} MHz")
