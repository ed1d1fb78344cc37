#!/usr/bin/env python
"""vkt_bcn_cuda_compress_batch throughput by texture size and lanes per device (VKT_BCN_BATCH_LANES).
Usage: batch_lanes.py  (prints one line per size x lanes)"""
import ctypes as C
import os
import subprocess
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

if len(sys.argv) > 1:
    import numpy as np
    import torch
    from vierkant_b200 import capi, synth
    size, count = int(sys.argv[1]), int(sys.argv[2])
    imgs = [torch.from_numpy(synth.make_texture(size, size, 1 if i % 4 == 3 else 0, seed=i)).pin_memory() for i in range(count)]
    plan = capi.compress_plan(size, size, True)
    L = plan.num_levels
    outs = [[torch.empty((int(plan.level_num_blocks[l]), 16), dtype=torch.uint8).pin_memory() for l in range(L)] for _ in range(count)]
    keep = [(C.c_void_p * L)(*[o.data_ptr() for o in out]) for out in outs]
    srcs = (capi.Source * count)()
    for i in range(count):
        srcs[i] = capi.Source(imgs[i].data_ptr(), size, size, 4, capi.MODE_BC7, keep[i])
    npix = count * sum(int(plan.level_width[l]) * int(plan.level_height[l]) for l in range(L))
    with capi.BcnContext([0]) as ctx:
        ts = []
        for rep in range(8):
            t0 = time.perf_counter()
            ctx._check(ctx.lib.vkt_bcn_cuda_compress_batch(ctx.handle, srcs, count, 1, None))
            ts.append(time.perf_counter() - t0)
        best = min(ts[2:])
        import hashlib
        h = hashlib.sha1(b"".join(o.numpy().tobytes() for out in outs for o in out)).hexdigest()[:10]
        print(f"{count} x {size}^2  lanes {os.environ.get('VKT_BCN_BATCH_LANES', 'auto'):>4}: {best * 1e3:8.3f} ms  {npix / best * 1e-6:8.0f} Mpixel/s  sha1 {h}")
else:
    for size, count in ((512, 64), (1024, 32), (2048, 16), (4096, 8)):
        for lanes in ("1", "2", "4", "8", None):
            env = dict(os.environ)
            if lanes:
                env["VKT_BCN_BATCH_LANES"] = lanes
            else:
                env.pop("VKT_BCN_BATCH_LANES", None)
            subprocess.run([sys.executable, __file__, str(size), str(count)], env=env)
