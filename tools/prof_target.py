#!/usr/bin/env python
"""Profiling target: N launches of the BC7 kernel on one device-resident synthetic level (for ncu).
Usage: prof_target.py [--size 4096] [--kind 0] [--launches 3] [--uber 0] [--fb 1]"""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from vierkant_b200 import capi, synth  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--size", type=int, default=4096)
ap.add_argument("--kind", type=int, default=0)
ap.add_argument("--launches", type=int, default=3)
ap.add_argument("--uber", type=int, default=0)
ap.add_argument("--fb", type=int, default=1)
ap.add_argument("--lib", default=None, help="alternative libvierkant_bcn_cuda.so (tuning variants)")
a = ap.parse_args()
if a.lib:
    capi._lib = capi.load_library(os.path.abspath(a.lib))
img = synth.make_texture(a.size, a.size, a.kind)
d_in = torch.from_numpy(img).cuda()
d_out = torch.empty(((a.size // 4) ** 2, 16), dtype=torch.uint8, device="cuda")
p = capi.default_params(uber_level=a.uber, mode17_partition_estimation_filterbank=a.fb)
with capi.BcnContext([0]) as ctx:
    s = torch.cuda.current_stream().cuda_stream
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(a.launches + 1)]
    ev[0].record()
    for i in range(a.launches):
        ctx.encode_bc7_device(d_in, a.size, a.size, 4, d_out, p, 0, s)
        ev[i + 1].record()
    torch.cuda.synchronize()
    ms = [ev[i].elapsed_time(ev[i + 1]) for i in range(a.launches)]
    import hashlib
    digest = hashlib.sha1(d_out.cpu().numpy().tobytes()).hexdigest()[:12]
    print("kernel ms:", ["%.3f" % m for m in ms], "Mpix/s: %.1f" % (a.size * a.size / min(ms) * 1e-3), "sha1", digest)
