#!/usr/bin/env python
"""vkt_bcn_cuda_compress() of one 4096^2 chain with the source / the destinations in pinned host or device memory.
Usage: e2e_variants.py [size]"""
import ctypes as C
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402
import torch  # noqa: E402

from vierkant_b200 import capi, synth  # noqa: E402

size = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
img = synth.make_texture(size, size, 0)
plan = capi.compress_plan(size, size, True)
L = plan.num_levels
npix = sum(int(plan.level_width[l]) * int(plan.level_height[l]) for l in range(L))
h_src, d_src = torch.from_numpy(img).pin_memory(), torch.from_numpy(img).cuda()
h_out = [torch.empty((int(plan.level_num_blocks[l]), 16), dtype=torch.uint8).pin_memory() for l in range(L)]
d_out = [torch.empty((int(plan.level_num_blocks[l]), 16), dtype=torch.uint8, device="cuda") for l in range(L)]
with capi.BcnContext([0]) as ctx:
    for name, src, outs in (("device src, device dst", d_src, d_out), ("pinned src, device dst", h_src, d_out), ("device src, pinned dst", d_src, h_out),
                            ("pinned src, pinned dst", h_src, h_out)):
        ptrs = (C.c_void_p * L)(*[o.data_ptr() for o in outs])

        def call():
            ctx._check(ctx.lib.vkt_bcn_cuda_compress(ctx.handle, capi.MODE_BC7, src.data_ptr(), size, size, 4, 1, None, ptrs))
        for _ in range(3):
            call()
        ts = []
        for _ in range(20):
            t0 = time.perf_counter()
            call()
            ts.append(time.perf_counter() - t0)
        print(f"{size}^2 {name:24s} mean {np.mean(ts) * 1e3:.3f} ms  best {min(ts) * 1e3:.3f} ms  {npix / np.mean(ts) * 1e-6:.0f} Mpix/s", flush=True)
