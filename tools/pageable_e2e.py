#!/usr/bin/env python
"""vkt_bcn_cuda_compress() with PAGEABLE host buffers (what the C++ drop-in passes: crocore image data in, std::vector
blocks out) against pinned ones.  Usage: pageable_e2e.py [size]"""
import ctypes as C
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402
import torch  # noqa: E402

from vierkant_b200 import capi, synth  # noqa: E402

size = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
src = synth.make_texture(size, size, 0)
plan = capi.compress_plan(size, size, True)
npix = sum(int(plan.level_width[l]) * int(plan.level_height[l]) for l in range(plan.num_levels))
with capi.BcnContext([0]) as ctx:
    ref = None
    for kind in ("pinned", "pageable", "pageable in, pinned out", "pinned in, pageable out"):
        pin_in, pin_out = kind in ("pinned", "pinned in, pageable out"), kind in ("pinned", "pageable in, pinned out")
        t_in = torch.from_numpy(src.copy())
        t_in = t_in.pin_memory() if pin_in else t_in
        outs = [torch.empty((int(plan.level_num_blocks[l]), 16), dtype=torch.uint8) for l in range(plan.num_levels)]
        outs = [o.pin_memory() if pin_out else o for o in outs]
        ptrs = (C.c_void_p * plan.num_levels)(*[o.data_ptr() for o in outs])

        def call():
            ctx._check(ctx.lib.vkt_bcn_cuda_compress(ctx.handle, capi.MODE_BC7, t_in.data_ptr(), size, size, 4, 1, None, ptrs))
        for _ in range(3):
            call()
        ts = []
        for _ in range(10):
            t0 = time.perf_counter()
            call()
            ts.append(time.perf_counter() - t0)
        h = hash(b"".join(o.numpy().tobytes() for o in outs))
        ref = h if ref is None else ref
        print(f"{size}^2 {kind:26s} mean {np.mean(ts) * 1e3:.3f} ms  best {min(ts) * 1e3:.3f} ms  {npix / np.mean(ts) * 1e-6:.0f} Mpix/s  same={h == ref}", flush=True)

# ---- a material batch through vkt_bcn_cuda_compress_batch (what the drop-in's span overload calls): 8 textures of size^2 / 2
if len(sys.argv) > 2 and sys.argv[2] == "batch":
    n, bs = 8, max(size // 2, 256)
    bplan = capi.compress_plan(bs, bs, True)
    bpix = n * sum(int(bplan.level_width[l]) * int(bplan.level_height[l]) for l in range(bplan.num_levels))
    with capi.BcnContext([0]) as ctx:
        for pinned in (True, False):
            imgs = [torch.from_numpy(synth.make_texture(bs, bs, 1 if i % 4 == 3 else 0, seed=i)) for i in range(n)]
            imgs = [t.pin_memory() if pinned else t for t in imgs]
            outs = [[torch.empty((int(bplan.level_num_blocks[l]), 16), dtype=torch.uint8) for l in range(bplan.num_levels)] for _ in range(n)]
            outs = [[o.pin_memory() if pinned else o for o in lv] for lv in outs]
            keep = [(C.c_void_p * bplan.num_levels)(*[o.data_ptr() for o in lv]) for lv in outs]
            srcs = (capi.Source * n)(*[capi.Source(imgs[i].data_ptr(), bs, bs, 4, capi.MODE_BC7, keep[i]) for i in range(n)])

            def bcall():
                ctx._check(ctx.lib.vkt_bcn_cuda_compress_batch(ctx.handle, srcs, n, 1, None))
            for _ in range(2):
                bcall()
            ts = []
            for _ in range(8):
                t0 = time.perf_counter()
                bcall()
                ts.append(time.perf_counter() - t0)
            print(f"batch of {n} x {bs}^2 {'pinned' if pinned else 'pageable':9s} mean {np.mean(ts) * 1e3:.3f} ms  {bpix / np.mean(ts) * 1e-6:.0f} Mpix/s", flush=True)
