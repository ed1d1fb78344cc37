#!/usr/bin/env python
"""vkt_bcn_cuda_encode_bc7 (one pre-resized level, host buffers in and out) with pageable and with pinned buffers.  Usage: encode_pageable.py [size]"""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402
import torch  # noqa: E402

from vierkant_b200 import capi, synth  # noqa: E402

size = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
img = synth.make_texture(size, size, 0)
with capi.BcnContext([0]) as ctx:
    ref = None
    for kind in ("pageable", "pinned"):
        src = torch.from_numpy(img.copy())
        out = torch.empty(((size // 4) ** 2, 16), dtype=torch.uint8)
        if kind == "pinned":
            src, out = src.pin_memory(), out.pin_memory()
        call = lambda: ctx._check(ctx.lib.vkt_bcn_cuda_encode_bc7(ctx.handle, src.data_ptr(), size, size, 4, 0, None, out.data_ptr()))
        for _ in range(3):
            call()
        ts = []
        for _ in range(10):
            t0 = time.perf_counter()
            call()
            ts.append(time.perf_counter() - t0)
        h = hash(out.numpy().tobytes())
        ref = h if ref is None else ref
        print(f"encode_bc7 {size}^2 {kind:8s}: mean {np.mean(ts) * 1e3:.3f} ms  best {min(ts) * 1e3:.3f} ms  {size * size / np.mean(ts) * 1e-6:.0f} Mpix/s  same={h == ref}", flush=True)
