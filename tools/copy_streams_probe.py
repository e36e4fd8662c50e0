import torch, time
n = 64 << 20
h_in = torch.empty(n, dtype=torch.float32).pin_memory(); h_out = torch.empty(n, dtype=torch.float32).pin_memory()
d = torch.empty(n, device="cuda")
def t(fn, reps=6):
    fn(); torch.cuda.synchronize()
    best = 1e9
    for _ in range(reps):
        t0 = time.perf_counter(); fn(); torch.cuda.synchronize(); best = min(best, time.perf_counter() - t0)
    return best
for mib in (8, 16, 32):
    c = (mib << 20) // 4
    for nstream in (1, 2, 3, 4, 8):
        ups = [torch.cuda.Stream() for _ in range(nstream)]
        def pipe():
            # chunk i: H2D then D2H on stream i % nstream (in-stream order => D2H(i) right behind H2D(i); with several streams the
            # copies of different chunks overlap on the engines)
            for i, o in enumerate(range(0, n, c)):
                with torch.cuda.stream(ups[i % nstream]):
                    d[o:o + c].copy_(h_in[o:o + c], non_blocking=True)
                    h_out[o:o + c].copy_(d[o:o + c], non_blocking=True)
        dt = t(pipe)
        print(f"{mib:2d} MiB chunks, {nstream} streams: {dt*1e3:.2f} ms ({n*4/dt*1e-9:.1f} GB/s per direction)", flush=True)
