"""How fast can this box move 256 MiB each way over PCIe (pinned host memory)?  Bounds bench.py's e2e number."""
import torch, time
n = 64 << 20
h_in = torch.empty(n, dtype=torch.float32).pin_memory(); h_out = torch.empty(n, dtype=torch.float32).pin_memory()
d_in = torch.empty(n, device="cuda"); d_out = torch.empty(n, device="cuda")
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
def t(fn, reps=5):
    fn(); torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(reps): fn()
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / reps
def h2d():
    with torch.cuda.stream(s1): d_in.copy_(h_in, non_blocking=True)
def d2h():
    with torch.cuda.stream(s2): h_out.copy_(d_out, non_blocking=True)
def both(): h2d(); d2h()
for name, fn in (("H2D", h2d), ("D2H", d2h), ("both", both)):
    dt = t(fn)
    print(f"{name}: {dt*1e3:.2f} ms for 256 MiB{' each way' if name=='both' else ''} -> {n*4/dt*1e-9:.1f} GB/s per direction")

# the e2e pipeline's traffic pattern without its kernels: chunk i goes up, then comes back, on stream i % 4 (what bounds
# pqmf_roundtrip_host_f32 at a given chunk size; the whole-buffer numbers above are the asymptote)
for mib in (4, 16, 32, 64):
    c = (mib << 20) // 4
    streams = [torch.cuda.Stream() for _ in range(4)]
    def pipe():
        for i, o in enumerate(range(0, n, c)):
            with torch.cuda.stream(streams[i % 4]):
                d_in[o:o + c].copy_(h_in[o:o + c], non_blocking=True)
                h_out[o:o + c].copy_(d_in[o:o + c], non_blocking=True)
    dt = t(pipe)
    print(f"chunked pipeline, {mib:2d} MiB chunks, copies only: {dt*1e3:.2f} ms for 256 MiB each way -> {n*4/dt*1e-9:.1f} GB/s per direction")
