import sys, time, os
sys.path.insert(0, '/root/repo')
import numpy as np, torch
from vierkant_b200 import capi, synth
img = synth.make_texture(4096, 4096, 0)
with capi.BcnContext([0]) as ctx:
    for _ in range(3): out = ctx.encode_bc7(img)
    ts = []
    for _ in range(10):
        t0 = time.perf_counter(); out = ctx.encode_bc7(img); ts.append(time.perf_counter() - t0)
    print("encode_bc7 4096^2 pageable numpy in/out: mean %.3f ms best %.3f ms" % (np.mean(ts) * 1e3, min(ts) * 1e3))
