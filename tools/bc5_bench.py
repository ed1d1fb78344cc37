#!/usr/bin/env python
"""BC5 (rgbcx::encode_bc5) kernel on device-resident levels: time and HBM rate against MEASURED_PEAKS.json.  Usage: bc5_bench.py"""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402
import torch  # noqa: E402

from vierkant_b200 import capi, synth  # noqa: E402

peak = 6550.7
try:
    peak = json.load(open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")))["hbm_gbs"]
except Exception:  # noqa: BLE001
    pass
with capi.BcnContext([0]) as ctx:
    for size in (4096, 8192, 16384):
        imgs = [torch.from_numpy(synth.make_texture(size, size, 1, seed=s, rows=(0, size))).cuda() for s in range(2 if size > 8192 else 4)]
        out = torch.empty(((size // 4) ** 2, 16), dtype=torch.uint8, device="cuda")
        st = torch.cuda.current_stream().cuda_stream
        for i in range(3):
            ctx.encode_bc5_device(imgs[i % len(imgs)], size, size, 4, out, 0, st)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        n = 20
        e0.record()
        for i in range(n):
            ctx.encode_bc5_device(imgs[i % len(imgs)], size, size, 4, out, 0, st)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / n
        gbs = size * size * 5.0 / (ms * 1e-3) * 1e-9
        print(f"BC5 {size}^2: {ms * 1e3:.1f} us per level, {size * size / ms * 1e-3:.0f} Mpixel/s, {gbs:.0f} GB/s algorithmic (64 B in + 16 B out per block) = {gbs / peak:.2f} of the measured {peak:.0f} GB/s copy rate", flush=True)
