#!/usr/bin/env python
"""vkt_bcn_cuda_compress() of a 3-component (RGB) texture -- what stb_image hands vierkant for a JPEG -- against the same texture as RGBA.
Usage: e2e_rgb.py [size]"""
import ctypes as C
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402
import torch  # noqa: E402

from vierkant_b200 import capi, synth  # noqa: E402

size = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
rgba = synth.make_texture(size, size, 0)
plan = capi.compress_plan(size, size, True)
npix = sum(int(plan.level_width[l]) * int(plan.level_height[l]) for l in range(plan.num_levels))
with capi.BcnContext([0]) as ctx:
    res = {}
    for comps in (4, 3):
        src = torch.from_numpy(np.ascontiguousarray(rgba[..., :comps])).pin_memory()
        outs = [torch.empty((int(plan.level_num_blocks[l]), 16), dtype=torch.uint8).pin_memory() for l in range(plan.num_levels)]
        ptrs = (C.c_void_p * plan.num_levels)(*[o.data_ptr() for o in outs])

        def call():
            ctx._check(ctx.lib.vkt_bcn_cuda_compress(ctx.handle, capi.MODE_BC7, src.data_ptr(), size, size, comps, 1, None, ptrs))
        for _ in range(3):
            call()
        ts = []
        for _ in range(10):
            t0 = time.perf_counter()
            call()
            ts.append(time.perf_counter() - t0)
        res[comps] = [o.numpy().copy() for o in outs]
        print(f"{size}^2 comps {comps}: mean {np.mean(ts) * 1e3:.3f} ms  best {min(ts) * 1e3:.3f} ms  {npix / np.mean(ts) * 1e-6:.0f} Mpix/s", flush=True)
    print("RGB blocks == RGBA (alpha 255) blocks:", all(np.array_equal(a, b) for a, b in zip(res[3], res[4])))
