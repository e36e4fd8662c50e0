import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, pqmf_b200 as pq
from pqmf_b200 import _lib
B, T = 64, 1 << 20
mod = pq.PQMF(100, 16).cuda()
hx = torch.randint(-20000, 20000, (B, T, 1), dtype=torch.int16).pin_memory(); ho = torch.empty(B, T, 1, dtype=torch.int16).pin_memory()
hk = mod.hk.cpu().contiguous(); tab = mod._tables.cpu().contiguous()
def step():
    rc = _lib.cabi.pqmf_roundtrip_host_pcm16(hx.data_ptr(), None, ho.data_ptr(), hk.data_ptr(), tab.data_ptr(), B, T, 1, 16, 512, 0, int(mod._flags), 0)
    assert rc == 0
for _ in range(3): step()
t0 = time.perf_counter()
for _ in range(10): step()
dt = (time.perf_counter() - t0) / 10
print(f"pcm chunk {os.environ.get('PQMF_HOST_CHUNK_MIB', '8')} MiB slots {os.environ.get('PQMF_HOST_SLOTS','4')}: {dt*1e3:.2f} ms -> {B*T/dt*1e-9:.2f} Gsamples/s, {B*T*2/dt*1e-9:.1f} GB/s each way")
