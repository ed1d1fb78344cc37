#!/usr/bin/env python
"""Small run of every kernel for compute-sanitizer: BC7 (opaque + alpha, default and uber), BC5, the resize chain through
compress(), and compress_batch.  Usage: compute-sanitizer --tool memcheck|racecheck|synccheck python tools/sanitize_target.py"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from vierkant_b200 import capi, synth  # noqa: E402

with capi.BcnContext([0]) as ctx:
    a, b = synth.make_texture(256, 128, 0), synth.make_texture(132, 68, 1)
    ctx.encode_bc7(a)
    ctx.encode_bc7(b, capi.default_params(uber_level=2, mode17_partition_estimation_filterbank=0))
    ctx.encode_bc5(b)
    ctx.compress(b, capi.MODE_BC7, True)
    ctx.compress_batch([a, b, a[..., :3]], [capi.MODE_BC7, capi.MODE_BC5, capi.MODE_BC7], True)
print("ok")
