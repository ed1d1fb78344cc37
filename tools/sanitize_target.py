#!/usr/bin/env python
"""Small run of every kernel for compute-sanitizer: BC7 (opaque + alpha, default, uber and extended variant), BC5, the resize chain through
compress(), and compress_batch.  Usage: compute-sanitizer --tool memcheck|racecheck|synccheck python tools/sanitize_target.py"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from vierkant_b200 import capi, synth  # noqa: E402

with capi.BcnContext([0]) as ctx:
    a, b = synth.make_texture(256, 128, 0), synth.make_texture(132, 68, 1)
    ctx.encode_bc7(a)
    ctx.encode_bc7(b, capi.default_params(uber_level=2, mode17_partition_estimation_filterbank=0))
    # the extended kernel variant: forced selectors + reduced mode-6 quantisation + low-frequency partition weight, both metrics
    ctx.encode_bc7(b, capi.default_params(force_selectors=1, selectors=[3, 3, 2, 2, 1, 1, 0, 0, 0, 1, 2, 3, 3, 2, 1, 0], quant_mode6_endpoints=1,
                                          low_frequency_partition_weight=0.8, uber_level=1))
    ctx.encode_bc7(b, capi.default_params(low_frequency_partition_weight=0.6, perceptual=0, weights=[1, 1, 1, 1]))
    ctx.encode_bc5(b)
    # a chain large enough for the band pipeline (graded level-0 bands, level 1 resized band by band, its own encode lane)
    # (VKT_SAN_BIG=1, memcheck only: 2048^2 reaches the separate level-1 encode lane)
    n = 2048 if os.environ.get("VKT_SAN_BIG") else 512
    ctx.compress(synth.make_texture(2 * n if n < 2048 else n, n, 1), capi.MODE_BC7, True)
    ctx.compress(b, capi.MODE_BC7, True)
    ctx.compress_batch([a, b, a[..., :3]], [capi.MODE_BC7, capi.MODE_BC5, capi.MODE_BC7], True)
    # round 2: deferred destinations, the fused resize (>= 2^20 output samples per call), concurrent calls on lanes, a process shard
    ctx.compress_alloc(b, capi.MODE_BC7, True)
    if os.environ.get("VKT_SAN_BIG"):
        ctx.compress_alloc(synth.make_texture(2048, 1024, 0), capi.MODE_BC7, True)
    # concurrent calls on lanes: memcheck only (racecheck loses track of launches that several host threads issue at once --
    # "Internal Sanitizer Error: Detected a failure to track a kernel launch" -- and reports that as a hazard)
    if not os.environ.get("VKT_SAN_NO_THREADS"):
        import threading
        ts = [threading.Thread(target=lambda im=im: ctx.compress(im, capi.MODE_BC7, True)) for im in (a, b, a, b)]
        for t in ts:
            t.start()
        for t in ts:
            t.join()
    import ctypes as C
    import numpy as np
    plan = capi.compress_plan(256, 128, True)
    lv = [np.zeros((int(plan.level_num_blocks[l]), 16), dtype=np.uint8) for l in range(plan.num_levels)]
    ptrs = (C.c_void_p * plan.num_levels)(*[x.ctypes.data for x in lv])
    hand = np.zeros(max(int(capi.shard_plan(256, 128, True, 2).handover_bytes), 1), dtype=np.uint8)
    workers = [capi.BcnContext([0]) for _ in range(2)]
    for r, w in enumerate(workers):
        w.compress_shard_begin(capi.MODE_BC7, a, 256, 128, 4, True, None, r, 2, ptrs, hand)
    for r, w in enumerate(workers):
        w.compress_shard_end(capi.MODE_BC7, a, 256, 128, 4, True, None, r, 2, ptrs, hand)
    for w in workers:
        w.close()
print("ok")
