import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import pqmf_b200 as pq
torch.manual_seed(1)
streams, block, n_blocks = 512, 2048, 8
for exact in (True, False):
    mod = pq.CachedPQMF(100, 16, exact=exact).cuda()
    x = (0.5 * torch.randn(streams, 1, block * n_blocks, device="cuda")).clamp_(-1, 1)
    for trial in range(3):
        mod.reset_stream()
        ys = []
        for i in range(n_blocks):
            yb = mod.forward_stream(x[..., i * block:(i + 1) * block].contiguous())
            bad = torch.isnan(yb)
            if bad.any():
                idx = bad.nonzero()
                print(f"exact={exact} trial {trial} block {i}: {int(bad.sum())} NaNs, rows {sorted(set(idx[:,0].tolist()))[:8]} bands {sorted(set(idx[:,1].tolist()))[:8]} frames {idx[:,2].min().item()}..{idx[:,2].max().item()}")
            ys.append(yb)
        xz = torch.cat([torch.zeros(streams, 1, 256, device="cuda"), x], dim=-1)
        yo = mod.forward(xz)
        bad = torch.isnan(yo)
        print(f"exact={exact} trial {trial}: offline NaNs {int(bad.sum())}", (f"rows {sorted(set(bad.nonzero()[:,0].tolist()))[:8]} frames {bad.nonzero()[:,2].min().item()}..{bad.nonzero()[:,2].max().item()}" if bad.any() else ""),
              "max diff", float((torch.cat(ys, -1) - yo[..., :block * n_blocks // 16]).abs().max()))
