#!/usr/bin/env python
"""Per-call times of the C++ drop-in's vierkant::bcn::compress() (integration/_build/libvkt_dropin_test.so) on a 4096^2 chain,
next to vkt_bcn_cuda_compress with pageable and pinned buffers.  Usage: dropin_e2e.py [size] [calls]"""
import ctypes as C
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402

from vierkant_b200 import capi, synth  # noqa: E402

size = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
calls = int(sys.argv[2]) if len(sys.argv) > 2 else 30
os.environ.setdefault("VIERKANT_BCN_CUDA_DEVICES", "0")
D = C.CDLL(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "integration", "_build", "libvkt_dropin_test.so"))
D.dropin_compress.argtypes = [C.c_void_p, C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint32, C.c_int]
D.dropin_compress.restype = C.c_void_p
D.dropin_result_free.argtypes = [C.c_void_p]
srcs = [synth.make_texture(size, size, 0, seed=0xB200 + i) for i in range(4)]
ts, tf = [], []
for i in range(calls):
    t0 = time.perf_counter()
    r = D.dropin_compress(srcs[i % 4].ctypes.data, size, size, 4, capi.MODE_BC7, 1)
    t1 = time.perf_counter()
    D.dropin_result_free(r)
    t2 = time.perf_counter()
    ts.append((t1 - t0) * 1e3), tf.append((t2 - t1) * 1e3)
print("compress ms:", " ".join(f"{t:.2f}" for t in ts))
print("free     ms:", " ".join(f"{t:.2f}" for t in tf))
print(f"median compress {np.median(ts[3:]):.3f} ms, free {np.median(tf[3:]):.3f} ms")
