#!/usr/bin/env python
"""Host-side ceiling of the e2e numbers: N ranks (one GPU each, torchrun) copy at the same time, no kernels.
Per step every rank moves what one 4096^2 chain moves (64 MB up, 22 MB down); reported: aggregate GB/s for H2D alone, D2H
alone and both directions together, with ordinary pinned memory and with write-combined pinned memory on the upload side.
Usage: torchrun --nproc-per-node N tools/host_link_probe.py   (or plain python for N = 1)"""
import ctypes as C
import json
import os
import sys
import time

import torch
import torch.distributed as dist

rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
if world > 1:
    sys.stdout.flush()
    saved = os.dup(1)
    os.dup2(2, 1)
    dist.init_process_group("nccl", device_id=dev)
    dist.barrier()
    sys.stdout.flush()
    os.dup2(saved, 1)
rt = C.CDLL("libcudart.so.12")
rt.cudaHostAlloc.argtypes = [C.POINTER(C.c_void_p), C.c_size_t, C.c_uint]
rt.cudaMemcpyAsync.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, C.c_int, C.c_void_p]
UP, DOWN = 64 << 20, 22 << 20


def host_alloc(n, flags):
    p = C.c_void_p()
    assert rt.cudaHostAlloc(C.byref(p), n, flags) == 0
    C.memset(p, 1, n)
    return p


bufs = {"pinned": host_alloc(UP, 1), "write-combined": host_alloc(UP, 1 | 4)}  # portable (| write-combined)
h_down = host_alloc(DOWN, 1)
d_up, d_down = torch.empty(UP, dtype=torch.uint8, device=dev), torch.zeros(DOWN, dtype=torch.uint8, device=dev)
s_up, s_down = torch.cuda.Stream(), torch.cuda.Stream()


def run(kind, up, down, steps=40):
    def go():
        for _ in range(steps):
            if up:
                rt.cudaMemcpyAsync(d_up.data_ptr(), bufs[kind], UP, 1, C.c_void_p(s_up.cuda_stream))
            if down:
                rt.cudaMemcpyAsync(h_down, d_down.data_ptr(), DOWN, 2, C.c_void_p(s_down.cuda_stream))
        torch.cuda.synchronize()
    go()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    go()
    dt = time.perf_counter() - t0
    t = torch.tensor([dt], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return world * steps * ((UP if up else 0) + (DOWN if down else 0)) / float(t[0]) * 1e-9


out = {"n_gpus": world, "per_step_per_rank_mb": {"up": UP >> 20, "down": DOWN >> 20}}
for kind in bufs:
    out[kind] = {"h2d_gbs": run(kind, True, False), "both_gbs": run(kind, True, True)}
out["d2h_gbs"] = run("pinned", False, True)
if rank == 0:
    out["chain_ms_at_both_rate"] = {k: (UP + DOWN) * world / (out[k]["both_gbs"] * 1e9) * 1e3 for k in bufs}
    print(json.dumps(out))
if world > 1:
    dist.destroy_process_group()
