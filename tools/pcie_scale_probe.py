"""What does the HOST side of this box give N GPUs at once?  Pure cudaMemcpyAsync traffic, no kernels: every rank copies a pinned
host buffer to its GPU and another GPU buffer back to pinned host memory, concurrently on two streams -- the traffic pattern of
bench.py's `e2e` leg (pqmf_roundtrip_host_f32) -- and the ranks report the aggregate GB/s per direction.  bench.py's e2e at N GPUs
cannot beat these numbers; SCALE's e2e efficiency is to be read against them.

    python tools/pcie_scale_probe.py                      # one GPU
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29571 tools/pcie_scale_probe.py

Variants: plain pinned memory (cudaHostAlloc default) and write-combined source buffers for the H2D side."""
import ctypes
import json
import os
import time

import torch
import torch.distributed as dist


def main():
    world, rank, local = (int(os.environ.get(k, d)) for k, d in (("WORLD_SIZE", "1"), ("RANK", "0"), ("LOCAL_RANK", "0")))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)
    rt = ctypes.CDLL("libcudart.so", mode=ctypes.RTLD_GLOBAL) if False else None
    try:
        rt = ctypes.CDLL(os.path.join(os.path.dirname(torch.__file__), "lib", "libcudart.so.12"))
    except OSError:
        rt = ctypes.CDLL("libcudart.so.12")
    rt.cudaHostAlloc.argtypes = [ctypes.POINTER(ctypes.c_void_p), ctypes.c_size_t, ctypes.c_uint]
    rt.cudaMemcpyAsync.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_size_t, ctypes.c_int, ctypes.c_void_p]
    nbytes = 256 << 20
    d_in, d_out = torch.empty(nbytes, dtype=torch.uint8, device=dev), torch.zeros(nbytes, dtype=torch.uint8, device=dev)
    s_h2d, s_d2h = torch.cuda.Stream(dev), torch.cuda.Stream(dev)
    results = {}
    for name, flags in (("pinned", 0), ("write_combined_src", 4)):
        h_src, h_dst = ctypes.c_void_p(), ctypes.c_void_p()
        assert rt.cudaHostAlloc(ctypes.byref(h_src), nbytes, flags) == 0 and rt.cudaHostAlloc(ctypes.byref(h_dst), nbytes, 0) == 0
        ctypes.memset(h_src, 1, nbytes)
        for mode in ("h2d", "d2h", "both"):
            def issue():
                if mode in ("h2d", "both"):
                    rt.cudaMemcpyAsync(d_in.data_ptr(), h_src, nbytes, 1, ctypes.c_void_p(s_h2d.cuda_stream))
                if mode in ("d2h", "both"):
                    rt.cudaMemcpyAsync(h_dst, d_out.data_ptr(), nbytes, 2, ctypes.c_void_p(s_d2h.cuda_stream))
            for _ in range(2):
                issue()
            torch.cuda.synchronize()
            if world > 1:
                dist.barrier()
            reps = 8
            t0 = time.perf_counter()
            for _ in range(reps):
                issue()
            torch.cuda.synchronize()
            dt = time.perf_counter() - t0
            if world > 1:
                t = torch.tensor([dt], device=dev, dtype=torch.float64)
                dist.all_reduce(t, op=dist.ReduceOp.MAX)
                dt = float(t.item())
            results[f"{name}/{mode}"] = round(world * reps * nbytes / dt * 1e-9, 1)  # aggregate GB/s PER DIRECTION
        rt.cudaFreeHost(h_src)
        rt.cudaFreeHost(h_dst)
    if rank == 0:
        print(json.dumps({"n_gpus": world, "aggregate_GBps_per_direction": results,
                          "cpu_affinity": len(os.sched_getaffinity(0)), "note": "256 MiB per copy, 8 copies per direction, max time over ranks"}))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
