"""CTA-pair kernels vs the single-CTA kernels: identical arithmetic, so the outputs must match bit for bit."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import pqmf_b200 as pq
from pqmf_b200 import _lib
mod = pq.PQMF(100, 16).cuda()
for B, F in ((64, 65536), (25, 2048 + 516), (97, 515), (1, 512 * 97)):
    x = (0.5 * torch.randn(B, 1, 16 * F, device="cuda")).clamp_(-1, 1)
    y_pair = torch.ops.pqmf_b200.analysis(x, mod.hk, mod._tables, F, mod._flags)
    y_one = torch.ops.pqmf_b200.analysis(x, mod.hk, mod._tables, F, mod._flags | _lib.PQMF_FLAG_NO_PAIR)
    o_pair = torch.ops.pqmf_b200.synthesis(y_one, mod.hk, mod._tables, 0, mod._flags)
    o_one = torch.ops.pqmf_b200.synthesis(y_one, mod.hk, mod._tables, 0, mod._flags | _lib.PQMF_FLAG_NO_PAIR)
    torch.cuda.synchronize()
    print(B, F, "analysis equal:", torch.equal(y_pair, y_one), float((y_pair - y_one).abs().max()), "synthesis equal:", torch.equal(o_pair, o_one), float((o_pair - o_one).abs().max()))
