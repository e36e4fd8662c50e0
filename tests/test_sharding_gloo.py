"""N > 1 host logic on CPU: two gloo ranks shard the rows of a batch, each processes only its own rows (the oracle stands
in for the CUDA kernels here), and the gathered result equals the single-process result -- no data-path collective is
needed because rows are independent."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, tmp):
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from oracle import pqmf_oracle as O
    from pqmf_b200.sharding import shard_rows

    hk = np.load(os.path.join(ROOT, "tests", "golden", "bank_M16.npz"))["hk"]
    x = O.audio_like((7, 2048), 123)  # 7 rows over 2 ranks: ragged split
    lo, hi = shard_rows(x.shape[0], world, rank)
    y_local = torch.from_numpy(O.analysis(x[lo:hi], hk))
    dist.barrier()
    # gather only to CHECK the result (the benchmark never gathers)
    sizes = [shard_rows(x.shape[0], world, r) for r in range(world)]
    bufs = [torch.zeros(b - a, 16, 128, dtype=torch.float64) for a, b in sizes]
    dist.all_gather(bufs, y_local) if len({b - a for a, b in sizes}) == 1 else None
    if len({b - a for a, b in sizes}) != 1:  # ragged: exchange through files instead of padding a collective
        np.save(os.path.join(tmp, f"y{rank}.npy"), y_local.numpy())
        dist.barrier()
        bufs = [torch.from_numpy(np.load(os.path.join(tmp, f"y{r}.npy"))) for r in range(world)]
    full = torch.cat(bufs, 0).numpy()
    ref = O.analysis(x, hk)
    ok = torch.tensor([float(np.array_equal(full, ref))])
    dist.all_reduce(ok, op=dist.ReduceOp.MIN)
    t = torch.tensor([float(rank + 1)])
    dist.all_reduce(t, op=dist.ReduceOp.MAX)  # the max-over-ranks reduction bench.py uses for its timing
    assert ok.item() == 1.0 and t.item() == float(world)
    dist.destroy_process_group()


def test_two_rank_row_sharding(tmp_path):
    port = 29600 + (os.getpid() % 300)
    mp.spawn(_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)


def test_shard_rows_partition():
    sys.path.insert(0, ROOT)
    from pqmf_b200.sharding import shard_rows

    for n in (0, 1, 7, 64, 4096, 16384):
        for w in (1, 2, 3, 4, 8):
            spans = [shard_rows(n, w, r) for r in range(w)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(spans[i][1] == spans[i + 1][0] for i in range(w - 1))
            sizes = [b - a for a, b in spans]
            assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        shard_rows(4, 2, 2)


def test_c_abi_sharding_rule_equals_the_python_one():
    """pqmf_roundtrip_host_multi_f32 splits rows with pqmf_shard_rows: the same contiguous split as pqmf_b200.sharding.shard_rows."""
    import ctypes

    sys.path.insert(0, ROOT)
    from pqmf_b200 import _lib
    from pqmf_b200.sharding import shard_rows

    f = _lib.cabi.pqmf_shard_rows
    f.restype = None
    f.argtypes = [ctypes.c_long, ctypes.c_int, ctypes.c_int, ctypes.POINTER(ctypes.c_long), ctypes.POINTER(ctypes.c_long)]
    for n in (0, 1, 7, 37, 64, 4096, 16385):
        for w in (1, 2, 3, 4, 8):
            for r in range(w):
                a, c = ctypes.c_long(), ctypes.c_long()
                f(n, w, r, ctypes.byref(a), ctypes.byref(c))
                lo, hi = shard_rows(n, w, r)
                assert (a.value, a.value + c.value) == (lo, hi)
