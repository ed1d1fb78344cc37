#!/usr/bin/env python
"""End-to-end time of vkt_bcn_cuda_compress() per texture size and level-0 band schedule (VKT_BCN_BANDS, tuning only).
Usage: e2e_sweep.py [--sizes 1024,2048,4096] [--schemes "default;16,32;8,16,24,32"]"""
import argparse
import ctypes as C
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from vierkant_b200 import capi, synth  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--sizes", default="1024,2048,4096")
ap.add_argument("--schemes", default="default")
ap.add_argument("--reps", type=int, default=20)
a = ap.parse_args()
with capi.BcnContext([0]) as ctx:
    for size in [int(x) for x in a.sizes.split(",")]:
        img = torch.from_numpy(synth.make_texture(size, size, 0)).pin_memory()
        plan = capi.compress_plan(size, size, True)
        outs = [torch.empty((int(plan.level_num_blocks[l]), 16), dtype=torch.uint8).pin_memory() for l in range(plan.num_levels)]
        ptrs = (C.c_void_p * plan.num_levels)(*[o.data_ptr() for o in outs])
        npix = sum(int(plan.level_width[l]) * int(plan.level_height[l]) for l in range(plan.num_levels))

        def call():
            ctx._check(ctx.lib.vkt_bcn_cuda_compress(ctx.handle, capi.MODE_BC7, img.data_ptr(), size, size, 4, 1, None, ptrs))
        ref = None
        for scheme in a.schemes.split(";"):
            if scheme == "default":
                os.environ.pop("VKT_BCN_BANDS", None)
            else:
                os.environ["VKT_BCN_BANDS"] = scheme
            for _ in range(3):
                call()
            best, tot = 1e9, 0.0
            for _ in range(a.reps):
                t0 = time.perf_counter()
                call()
                dt = time.perf_counter() - t0
                best, tot = min(best, dt), tot + dt
            h = hash(b"".join(o.numpy().tobytes() for o in outs))
            ref = h if ref is None else ref
            print(f"size {size} bands {scheme:28s} mean {tot / a.reps * 1e3:.3f} ms  best {best * 1e3:.3f} ms  {npix / (tot / a.reps) * 1e-6:.0f} Mpix/s  same={h == ref}", flush=True)
