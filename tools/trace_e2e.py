#!/usr/bin/env python
"""Device timeline of one vkt_bcn_cuda_compress() call (VKT_BCN_TRACE=1) plus the host link rates it runs against.
Usage: trace_e2e.py [--size 4096] [--kind 0]"""
import argparse
import ctypes as C
import os
import sys
import time

os.environ["VKT_BCN_TRACE"] = "1"
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from vierkant_b200 import capi, synth  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--size", type=int, default=4096)
ap.add_argument("--kind", type=int, default=0)
ap.add_argument("--pageable", action="store_true", help="plain (unpinned) host buffers on both sides, as the C++ drop-in passes them")
a = ap.parse_args()

# link rates: pinned host <-> device, 256 MB
h = torch.empty(256 << 20, dtype=torch.uint8).pin_memory()
d = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
for name, fn in (("H2D", lambda: d.copy_(h, non_blocking=True)), ("D2H", lambda: h.copy_(d, non_blocking=True))):
    fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(4):
        fn()
    e1.record()
    torch.cuda.synchronize()
    print(f"{name}: {4 * (256 << 20) / (e0.elapsed_time(e1) * 1e-3) * 1e-9:.1f} GB/s", file=sys.stderr)

img = torch.from_numpy(synth.make_texture(a.size, a.size, a.kind))
img = img if a.pageable else img.pin_memory()
plan = capi.compress_plan(a.size, a.size, True)
outs = [torch.zeros((int(plan.level_num_blocks[l]), 16), dtype=torch.uint8) for l in range(plan.num_levels)]
outs = outs if a.pageable else [o.pin_memory() for o in outs]
ptrs = (C.c_void_p * plan.num_levels)(*[t.data_ptr() for t in outs])
p = capi.default_params()
with capi.BcnContext([0]) as ctx:
    for i in range(4):
        print(f"--- call {i}", file=sys.stderr)
        t0 = time.perf_counter()
        ctx._check(ctx.lib.vkt_bcn_cuda_compress(ctx.handle, capi.MODE_BC7, img.data_ptr(), a.size, a.size, 4, 1, C.byref(p), ptrs))
        print(f"host wall: {(time.perf_counter() - t0) * 1e3:.3f} ms", file=sys.stderr)
