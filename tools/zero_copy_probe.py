"""Experiment: the kernels reading x straight from pinned (mapped) host memory and writing the reconstruction straight back to it,
no staging copies -- against pqmf_roundtrip_host_f32's chunked copy pipeline.

Measured (one B200 box): staged copies 6.38 ms per 64 x 2^20 round trip; analysis reading host x alone 5.48 ms (49 GB/s), synthesis writing
host out alone 5.11 ms (52.5 GB/s) -- no better than the copy engines (55.5 / 57.2 GB/s) -- and with both at once (two streams, each
kernel on half of the SMs through a temporary grid-cap flag) 7.07 ms = 38 GB/s per direction against the copy engines' 49.5.  Not kept."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import pqmf_b200 as pq
from pqmf_b200 import _lib
B, T = 64, 1 << 20
mod = pq.PQMF(100, 16).cuda()
hx = torch.randn(B, T).clamp_(-1, 1).pin_memory(); ho = torch.empty(B, T).pin_memory(); ho2 = torch.empty(B, T).pin_memory()
hk = mod.hk.contiguous(); tab = mod._tables.contiguous()
dy = torch.empty(B, 16, T // 16, device="cuda")
dx = torch.empty(B, T, device="cuda"); do = torch.empty(B, T, device="cuda")
f = int(mod._flags)
st = torch.cuda.current_stream().cuda_stream
def zero_copy():
    rc = _lib.cabi.pqmf_analysis_f32(hx.data_ptr(), dy.data_ptr(), hk.data_ptr(), tab.data_ptr(), B, T, T // 16, 16, 512, f, st); assert rc == 0
    rc = _lib.cabi.pqmf_synthesis_f32(dy.data_ptr(), ho.data_ptr(), hk.data_ptr(), tab.data_ptr(), B, T // 16, 16, 512, 0, f, st); assert rc == 0
    torch.cuda.synchronize()
def read_only():
    rc = _lib.cabi.pqmf_analysis_f32(hx.data_ptr(), dy.data_ptr(), hk.data_ptr(), tab.data_ptr(), B, T, T // 16, 16, 512, f, st); assert rc == 0
    torch.cuda.synchronize()
def write_only():
    rc = _lib.cabi.pqmf_synthesis_f32(dy.data_ptr(), ho.data_ptr(), hk.data_ptr(), tab.data_ptr(), B, T // 16, 16, 512, 0, f, st); assert rc == 0
    torch.cuda.synchronize()
hkc = mod.hk.cpu().contiguous(); tabc = mod._tables.cpu().contiguous()
def staged():
    rc = _lib.cabi.pqmf_roundtrip_host_f32(hx.data_ptr(), None, ho2.data_ptr(), hkc.data_ptr(), tabc.data_ptr(), B, T, 16, 512, 0, f, 0); assert rc == 0
for name, fn in (("staged copies", staged), ("zero copy (both kernels)", zero_copy), ("analysis reading host x", read_only), ("synthesis writing host out", write_only)):
    for _ in range(2): fn()
    t0 = time.perf_counter()
    for _ in range(5): fn()
    dt = (time.perf_counter() - t0) / 5
    print(f"{name:30s}: {dt*1e3:.2f} ms -> {B*T/dt*1e-9:.2f} Gsamples/s", flush=True)
staged(); zero_copy()
print("same result:", torch.equal(ho, ho2))

