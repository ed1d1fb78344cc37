#!/usr/bin/env python
"""The two strip-kernel shapes of a 4096^2 chain through vkt_bcn_cuda_resize_u8 (run under
`ncu --metrics gpu__time_duration.sum --clock-control none -k regex:resize_` for kernel times; VKT_BCN_OLD_STRIP=1 selects the
first version of the strip kernel).  Prints the sha1 of every result: the versions must agree."""
import hashlib
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from vierkant_b200 import capi, synth  # noqa: E402

size = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
img = synth.make_texture(size, size, 0)
with capi.BcnContext([0]) as ctx:
    for ow in (size, size // 2):
        for _ in range(3):
            out = ctx.resize_u8(img, ow, ow)
        print(f"{size} -> {ow}: sha1 {hashlib.sha1(out.tobytes()).hexdigest()[:12]}")
