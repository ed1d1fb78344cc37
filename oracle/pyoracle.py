"""ctypes bindings for the test oracles.  TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference leg may import this module.
The product path (vierkant_b200/*) never does.

  RefOracle   -> oracle/_ref/libvkt_ref.so  (the unmodified reference, compiled in place by oracle/Makefile)
  PortOracle  -> oracle/libvkt_oracle.so    (this repo's C restatement: bc7_oracle.c, stbir_oracle.c, bc5_oracle.c)
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REF_SO = os.path.join(HERE, "_ref", "libvkt_ref.so")
PORT_SO = os.path.join(HERE, "libvkt_oracle.so")


class Bc7Params(C.Structure):
    """Mirror of bc7enc_compress_block_params (extern/bc7enc_rdo/bc7enc.h:14-75), field by field."""
    _fields_ = [
        ("mode_mask", C.c_uint32),
        ("max_partitions", C.c_uint32),
        ("weights", C.c_uint32 * 4),
        ("uber_level", C.c_uint32),
        ("perceptual", C.c_uint32),
        ("try_least_squares", C.c_uint32),
        ("mode17_partition_estimation_filterbank", C.c_uint32),
        ("force_alpha", C.c_uint32),
        ("force_selectors", C.c_uint32),
        ("selectors", C.c_uint8 * 16),
        ("quant_mode6_endpoints", C.c_uint32),
        ("bias_mode1_pbits", C.c_uint32),
        ("pbit1_weight", C.c_float),
        ("mode1_error_weight", C.c_float),
        ("mode5_error_weight", C.c_float),
        ("mode6_error_weight", C.c_float),
        ("mode7_error_weight", C.c_float),
        ("low_frequency_partition_weight", C.c_float),
    ]


def default_params(**overrides) -> Bc7Params:
    """bc7enc_compress_block_params_init() defaults (bc7enc.h:95-113) + overrides."""
    p = Bc7Params()
    p.mode_mask = 0xFFFFFFFF
    p.max_partitions = 64
    p.weights[:] = [128, 64, 16, 32]
    p.uber_level = 0
    p.perceptual = 1
    p.try_least_squares = 1
    p.mode17_partition_estimation_filterbank = 1
    p.pbit1_weight = 1.0
    p.mode1_error_weight = 1.0
    p.mode5_error_weight = 1.0
    p.mode6_error_weight = 1.0
    p.mode7_error_weight = 1.0
    p.low_frequency_partition_weight = 1.0
    for k, v in overrides.items():
        if k == "weights":
            p.weights[:] = list(v)
        elif k == "selectors":
            p.selectors[:] = list(v)
        else:
            if not hasattr(p, k):
                raise AttributeError(k)
            setattr(p, k, v)
    return p


def _u8(a):
    a = np.ascontiguousarray(a, dtype=np.uint8)
    return a, a.ctypes.data_as(C.POINTER(C.c_uint8))


def build(target: str = "all") -> None:
    """Run oracle/Makefile (gcc/g++ only).  `ref` is a no-op where /root/reference is absent."""
    subprocess.run(["make", "-s", "-C", HERE, target], check=True)


class _BlockOracle:
    """Shared wrapper: both libraries export the same <prefix>_* block-level entry points."""
    prefix = ""
    path = ""

    def __init__(self):
        if not os.path.exists(self.path):
            raise FileNotFoundError(self.path)
        self.lib = C.CDLL(self.path)
        L, p = self.lib, self.prefix
        getattr(L, p + "bc7_encode_blocks").argtypes = [C.POINTER(C.c_uint8), C.c_uint64, C.POINTER(Bc7Params),
                                                        C.POINTER(C.c_uint8), C.c_int]
        getattr(L, p + "bc7_encode_blocks").restype = None
        getattr(L, p + "bc7_unpack_blocks").argtypes = [C.POINTER(C.c_uint8), C.c_uint64, C.POINTER(C.c_uint8)]
        getattr(L, p + "bc7_unpack_blocks").restype = None
        if not hasattr(L, p + "bc5_encode_blocks") or not hasattr(L, p + "resize_u8"):
            return
        getattr(L, p + "bc5_encode_blocks").argtypes = [C.POINTER(C.c_uint8), C.c_uint64, C.POINTER(C.c_uint8)]
        getattr(L, p + "bc5_encode_blocks").restype = None
        getattr(L, p + "resize_u8").argtypes = [C.POINTER(C.c_uint8), C.c_uint32, C.c_uint32, C.c_uint32,
                                                C.POINTER(C.c_uint8), C.c_uint32, C.c_uint32]
        getattr(L, p + "resize_u8").restype = None

    @classmethod
    def available(cls) -> bool:
        return os.path.exists(cls.path)

    def encode_blocks(self, tiles: np.ndarray, params: Bc7Params | None = None, threads: int = 1) -> np.ndarray:
        """tiles: (n, 16, 4) uint8 -> (n, 16) uint8 BC7 blocks."""
        tiles, ptr = _u8(tiles)
        n = tiles.size // 64
        out = np.zeros((n, 16), dtype=np.uint8)
        pp = C.byref(params) if params is not None else None
        getattr(self.lib, self.prefix + "bc7_encode_blocks")(ptr, n, pp, out.ctypes.data_as(C.POINTER(C.c_uint8)),
                                                             threads)
        return out

    def encode_bc5_blocks(self, tiles: np.ndarray) -> np.ndarray:
        tiles, ptr = _u8(tiles)
        n = tiles.size // 64
        out = np.zeros((n, 16), dtype=np.uint8)
        getattr(self.lib, self.prefix + "bc5_encode_blocks")(ptr, n, out.ctypes.data_as(C.POINTER(C.c_uint8)))
        return out

    def unpack_blocks(self, blocks: np.ndarray) -> np.ndarray:
        blocks, ptr = _u8(blocks)
        n = blocks.size // 16
        out = np.zeros((n, 16, 4), dtype=np.uint8)
        getattr(self.lib, self.prefix + "bc7_unpack_blocks")(ptr, n, out.ctypes.data_as(C.POINTER(C.c_uint8)))
        return out

    def resize(self, img: np.ndarray, ow: int, oh: int) -> np.ndarray:
        """(H, W, C) uint8 -> (oh, ow, C) through crocore::Image_<uint8_t>::resize semantics (stbir defaults)."""
        img, ptr = _u8(img)
        h, w, c = img.shape
        out = np.zeros((oh, ow, c), dtype=np.uint8)
        getattr(self.lib, self.prefix + "resize_u8")(ptr, w, h, c, out.ctypes.data_as(C.POINTER(C.c_uint8)), ow, oh)
        return out


class RefOracle(_BlockOracle):
    prefix = "ref_"
    path = REF_SO

    def __init__(self):
        super().__init__()
        L = self.lib
        L.ref_compress.argtypes = [C.POINTER(C.c_uint8), C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint32, C.c_int,
                                   C.c_int]
        L.ref_compress.restype = C.c_void_p
        for name, res in [("num_levels", C.c_uint32), ("base_width", C.c_uint32), ("base_height", C.c_uint32),
                          ("mode", C.c_uint32), ("duration_ms", C.c_int64)]:
            f = getattr(L, "ref_result_" + name)
            f.argtypes = [C.c_void_p]
            f.restype = res
        L.ref_result_level_blocks.argtypes = [C.c_void_p, C.c_uint32]
        L.ref_result_level_blocks.restype = C.c_uint64
        L.ref_result_level_data.argtypes = [C.c_void_p, C.c_uint32]
        L.ref_result_level_data.restype = C.c_void_p
        L.ref_result_free.argtypes = [C.c_void_p]
        L.ref_hardware_concurrency.restype = C.c_uint

    def hardware_concurrency(self) -> int:
        return int(self.lib.ref_hardware_concurrency())

    def compress(self, img: np.ndarray, mode: int = 1, mipmaps: bool = False, threads: int = 0) -> dict:
        """vierkant::bcn::compress() of the reference.  mode: 0 = BC5, 1 = BC7."""
        img, ptr = _u8(img)
        h, w, c = img.shape
        L = self.lib
        r = L.ref_compress(ptr, w, h, c, mode, int(mipmaps), threads)
        try:
            levels = []
            for l in range(L.ref_result_num_levels(r)):
                n = L.ref_result_level_blocks(r, l)
                buf = (C.c_uint8 * (16 * n)).from_address(L.ref_result_level_data(r, l))
                levels.append(np.frombuffer(buf, dtype=np.uint8).reshape(n, 16).copy())
            return {"mode": L.ref_result_mode(r), "base_width": L.ref_result_base_width(r),
                    "base_height": L.ref_result_base_height(r), "levels": levels,
                    "duration_ms": L.ref_result_duration_ms(r)}
        finally:
            L.ref_result_free(r)


class PortOracle(_BlockOracle):
    prefix = "port_"
    path = PORT_SO

    def __init__(self):
        super().__init__()
        L = self.lib
        L.port_compress_num_levels.argtypes = [C.c_uint32, C.c_uint32, C.c_int]
        L.port_compress_num_levels.restype = C.c_uint32
        L.port_compress.argtypes = [C.POINTER(C.c_uint8), C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint32, C.c_int,
                                    C.POINTER(Bc7Params), C.POINTER(C.c_void_p), C.c_int]
        L.port_compress.restype = C.c_uint32

    def compress(self, img: np.ndarray, mode: int = 1, mipmaps: bool = False, threads: int = 1,
                 params: Bc7Params | None = None) -> dict:
        """Restatement of vierkant::bcn::compress() (oracle/compress_oracle.c); same result layout as RefOracle.compress."""
        img, ptr = _u8(img)
        h, w, c = img.shape
        r4 = lambda v: (v + 3) & ~3
        n = int(self.lib.port_compress_num_levels(w, h, int(mipmaps)))
        lw, lh, levels = r4(w), r4(h), []
        for _ in range(n):
            levels.append(np.zeros(((lw // 4) * (lh // 4), 16), dtype=np.uint8))
            lw, lh = r4(max(lw // 2, 1)), r4(max(lh // 2, 1))
        ptrs = (C.c_void_p * n)(*[l.ctypes.data for l in levels])
        pp = C.byref(params) if params is not None else None
        self.lib.port_compress(ptr, w, h, c, mode, int(mipmaps), pp, ptrs, threads)
        return {"mode": mode, "base_width": r4(w), "base_height": r4(h), "levels": levels}
