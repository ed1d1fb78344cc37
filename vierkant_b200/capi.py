"""ctypes binding of the C ABI in include/vierkant_bcn_cuda.h (libvierkant_bcn_cuda.so).

This is the reference-facing boundary seen from Python (tests, bench).  It contains no compute: every call goes to the
CUDA library, and a missing library or a missing GPU raises -- there is no CPU fallback.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

from . import build as _build

OK, ERR_INVALID, ERR_UNSUPPORTED, ERR_CUDA, ERR_NO_DEVICE, ERR_OOM = 0, -1, -2, -3, -4, -5
MODE_BC5, MODE_BC7 = 0, 1  # == vierkant::bcn::CompressionMode


class BcnError(RuntimeError):
    def __init__(self, code: int, msg: str):
        super().__init__(f"vkt_bcn_cuda error {code}: {msg}")
        self.code = code


class Bc7Params(C.Structure):
    """vkt_bc7_params: field-by-field mirror of bc7enc_compress_block_params (bc7enc.h:14-75)."""
    _fields_ = [
        ("mode_mask", C.c_uint32), ("max_partitions", C.c_uint32), ("weights", C.c_uint32 * 4),
        ("uber_level", C.c_uint32), ("perceptual", C.c_uint32), ("try_least_squares", C.c_uint32),
        ("mode17_partition_estimation_filterbank", C.c_uint32), ("force_alpha", C.c_uint32),
        ("force_selectors", C.c_uint32), ("selectors", C.c_uint8 * 16), ("quant_mode6_endpoints", C.c_uint32),
        ("bias_mode1_pbits", C.c_uint32), ("pbit1_weight", C.c_float), ("mode1_error_weight", C.c_float),
        ("mode5_error_weight", C.c_float), ("mode6_error_weight", C.c_float), ("mode7_error_weight", C.c_float),
        ("low_frequency_partition_weight", C.c_float),
    ]


class Image(C.Structure):
    """vkt_bcn_image"""
    _fields_ = [("pixels", C.c_void_p), ("width", C.c_uint32), ("height", C.c_uint32), ("comps", C.c_uint32),
                ("row_stride_bytes", C.c_uint32), ("out_blocks", C.c_void_p)]


class Plan(C.Structure):
    """vkt_bcn_plan"""
    _fields_ = [("base_width", C.c_uint32), ("base_height", C.c_uint32), ("num_levels", C.c_uint32),
                ("level_width", C.c_uint32 * 16), ("level_height", C.c_uint32 * 16),
                ("level_num_blocks", C.c_uint64 * 16)]


class Source(C.Structure):
    """vkt_bcn_source"""
    _fields_ = [("pixels", C.c_void_p), ("width", C.c_uint32), ("height", C.c_uint32), ("comps", C.c_uint32), ("mode", C.c_uint32),
                ("level_blocks", C.POINTER(C.c_void_p))]


class ShardPlan(C.Structure):
    """vkt_bcn_shard_plan"""
    _fields_ = [("num_levels", C.c_uint32), ("sliced_levels", C.c_uint32), ("workers", C.c_uint32), ("handover_bytes", C.c_uint64)]


class Stats(C.Structure):
    _fields_ = [("kernel_launches", C.c_uint64), ("h2d_bytes", C.c_uint64), ("d2h_bytes", C.c_uint64)]


EXPORTS = [
    "vkt_bc7_params_init", "vkt_bcn_cuda_device_count", "vkt_bcn_cuda_create", "vkt_bcn_cuda_destroy",
    "vkt_bcn_cuda_num_devices", "vkt_bcn_cuda_last_error", "vkt_bcn_cuda_encode_bc7", "vkt_bcn_cuda_encode_bc5",
    "vkt_bcn_cuda_encode_batch", "vkt_bcn_cuda_encode_bc7_device", "vkt_bcn_cuda_encode_bc5_device",
    "vkt_bcn_cuda_resize_u8", "vkt_bcn_cuda_compress_plan", "vkt_bcn_cuda_compress", "vkt_bcn_cuda_get_stats",
    "vkt_bcn_cuda_measure_issue_peak", "vkt_bcn_cuda_encode_batch_device", "vkt_bcn_cuda_compress_batch",
    "vkt_bcn_cuda_compress_shard_plan", "vkt_bcn_cuda_compress_shard_rows", "vkt_bcn_cuda_compress_shard_begin",
    "vkt_bcn_cuda_compress_shard_end", "vkt_bcn_cuda_host_register", "vkt_bcn_cuda_host_unregister",
    "vkt_bcn_cuda_compress_alloc", "vkt_bcn_cuda_import_external_fd", "vkt_bcn_cuda_release_external",
]
ALLOC_FN = C.CFUNCTYPE(C.c_void_p, C.c_void_p, C.c_uint32, C.c_size_t)  # vkt_bcn_alloc_fn

_lib = None


def load_library(path: str | None = None) -> C.CDLL:
    """dlopen libvierkant_bcn_cuda.so (in-tree).  Raises if it has not been built."""
    global _lib
    if _lib is not None and path is None:
        return _lib
    path = path or os.environ.get("VKT_BCN_LIB") or _build.CUDA_SO  # (VKT_BCN_LIB: a tuning build of the same ABI)
    if not os.path.exists(path):
        raise FileNotFoundError(f"{path} not built; run `python -c 'import __graft_entry__ as g; g.build()'` "
                                "(the encoder has no CPU fallback)")
    L = C.CDLL(path)
    u8p, u32, vp = C.POINTER(C.c_uint8), C.c_uint32, C.c_void_p
    L.vkt_bc7_params_init.argtypes = [C.POINTER(Bc7Params)]
    L.vkt_bc7_params_init.restype = None
    L.vkt_bcn_cuda_device_count.restype = C.c_int
    L.vkt_bcn_cuda_create.argtypes = [C.POINTER(vp), C.POINTER(C.c_int), C.c_int]
    L.vkt_bcn_cuda_destroy.argtypes = [vp]
    L.vkt_bcn_cuda_destroy.restype = None
    L.vkt_bcn_cuda_num_devices.argtypes = [vp]
    L.vkt_bcn_cuda_last_error.argtypes = [vp]
    L.vkt_bcn_cuda_last_error.restype = C.c_char_p
    L.vkt_bcn_cuda_encode_bc7.argtypes = [vp, vp, u32, u32, u32, u32, C.POINTER(Bc7Params), vp]
    L.vkt_bcn_cuda_encode_bc5.argtypes = [vp, vp, u32, u32, u32, u32, vp]
    L.vkt_bcn_cuda_encode_batch.argtypes = [vp, u32, C.POINTER(Image), u32, C.POINTER(Bc7Params)]
    L.vkt_bcn_cuda_encode_batch_device.argtypes = [vp, C.c_int, u32, C.POINTER(Image), u32, C.POINTER(Bc7Params), vp]
    L.vkt_bcn_cuda_encode_bc7_device.argtypes = [vp, C.c_int, vp, u32, u32, u32, u32, C.POINTER(Bc7Params), vp, vp]
    L.vkt_bcn_cuda_encode_bc5_device.argtypes = [vp, C.c_int, vp, u32, u32, u32, u32, vp, vp]
    L.vkt_bcn_cuda_resize_u8.argtypes = [vp, vp, u32, u32, u32, vp, u32, u32]
    L.vkt_bcn_cuda_compress_plan.argtypes = [u32, u32, C.c_int, C.POINTER(Plan)]
    L.vkt_bcn_cuda_compress.argtypes = [vp, u32, vp, u32, u32, u32, C.c_int, C.POINTER(Bc7Params), C.POINTER(vp)]
    L.vkt_bcn_cuda_compress_batch.argtypes = [vp, C.POINTER(Source), u32, C.c_int, C.POINTER(Bc7Params)]
    L.vkt_bcn_cuda_compress_alloc.argtypes = [vp, u32, vp, u32, u32, u32, C.c_int, C.POINTER(Bc7Params), ALLOC_FN, vp]
    L.vkt_bcn_cuda_compress_shard_plan.argtypes = [u32, u32, C.c_int, u32, C.POINTER(ShardPlan)]
    L.vkt_bcn_cuda_compress_shard_rows.argtypes = [u32, u32, C.c_int, u32, u32, u32, C.POINTER(u32), C.POINTER(u32)]
    shard_args = [vp, u32, vp, u32, u32, u32, C.c_int, C.POINTER(Bc7Params), u32, u32, C.POINTER(vp), vp]
    L.vkt_bcn_cuda_compress_shard_begin.argtypes = shard_args
    L.vkt_bcn_cuda_compress_shard_end.argtypes = shard_args
    L.vkt_bcn_cuda_import_external_fd.argtypes = [vp, C.c_int, C.c_int, C.c_uint64, C.POINTER(vp), C.POINTER(vp)]
    L.vkt_bcn_cuda_release_external.argtypes = [vp, C.c_int, vp, vp]
    L.vkt_bcn_cuda_host_register.argtypes = [vp, vp, C.c_size_t]
    L.vkt_bcn_cuda_host_unregister.argtypes = [vp, vp]
    L.vkt_bcn_cuda_get_stats.argtypes = [vp, C.POINTER(Stats)]
    L.vkt_bcn_cuda_measure_issue_peak.argtypes = [vp, C.c_int, C.POINTER(C.c_double)]
    if path == _build.CUDA_SO:
        _lib = L
    return L


def default_params(**overrides) -> Bc7Params:
    """vkt_bc7_params_init() (== bc7enc_compress_block_params_init, bc7enc.h:95-113) plus keyword overrides."""
    p = Bc7Params()
    load_library().vkt_bc7_params_init(C.byref(p))
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


def compress_plan(width: int, height: int, generate_mipmaps: bool) -> Plan:
    plan = Plan()
    rc = load_library().vkt_bcn_cuda_compress_plan(width, height, int(generate_mipmaps), C.byref(plan))
    if rc:
        raise BcnError(rc, "invalid size")
    return plan


def shard_plan(width: int, height: int, generate_mipmaps: bool, world: int) -> ShardPlan:
    """vkt_bcn_cuda_compress_shard_plan: how ONE chain is split over `world` single-GPU workers."""
    sp = ShardPlan()
    if load_library().vkt_bcn_cuda_compress_shard_plan(width, height, int(generate_mipmaps), world, C.byref(sp)):
        raise BcnError(ERR_INVALID, "invalid shard plan arguments")
    return sp


def shard_rows(width: int, height: int, generate_mipmaps: bool, rank: int, world: int, level: int) -> tuple[int, int]:
    """Block rows [r0, r1) of `level` that worker `rank` of `world` encodes."""
    r0, r1 = C.c_uint32(), C.c_uint32()
    if load_library().vkt_bcn_cuda_compress_shard_rows(width, height, int(generate_mipmaps), rank, world, level, C.byref(r0), C.byref(r1)):
        raise BcnError(ERR_INVALID, "invalid shard rows arguments")
    return int(r0.value), int(r1.value)


def device_count() -> int:
    return int(load_library().vkt_bcn_cuda_device_count())


def _ptr(a) -> int:
    """Address of a numpy array / torch tensor / raw int pointer."""
    if isinstance(a, int):
        return a
    if isinstance(a, np.ndarray):
        return a.ctypes.data
    if hasattr(a, "data_ptr"):
        return int(a.data_ptr())
    raise TypeError(type(a))


class BcnContext:
    """vkt_bcn_ctx wrapper.  devices: list of CUDA ordinals (None = all visible)."""

    def __init__(self, devices: list[int] | None = None):
        self.lib = load_library()
        self.handle = C.c_void_p()
        if devices is None:
            rc = self.lib.vkt_bcn_cuda_create(C.byref(self.handle), None, 0)
        else:
            arr = (C.c_int * len(devices))(*devices)
            rc = self.lib.vkt_bcn_cuda_create(C.byref(self.handle), arr, len(devices))
        if rc:
            raise BcnError(rc, (self.lib.vkt_bcn_cuda_last_error(None) or b"").decode())

    def close(self):
        if getattr(self, "handle", None) and self.handle.value:
            self.lib.vkt_bcn_cuda_destroy(self.handle)
            self.handle = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    def _check(self, rc: int):
        if rc:
            raise BcnError(rc, (self.lib.vkt_bcn_cuda_last_error(self.handle) or b"").decode())

    @property
    def num_devices(self) -> int:
        return int(self.lib.vkt_bcn_cuda_num_devices(self.handle))

    def stats(self) -> dict:
        s = Stats()
        self._check(self.lib.vkt_bcn_cuda_get_stats(self.handle, C.byref(s)))
        return {"kernel_launches": int(s.kernel_launches), "h2d_bytes": int(s.h2d_bytes), "d2h_bytes": int(s.d2h_bytes)}

    def measure_issue_peak(self, slot: int = 0) -> float:
        """Sustained integer issue rate of the device (lane-ops/s), measured by a short probe kernel."""
        v = C.c_double()
        self._check(self.lib.vkt_bcn_cuda_measure_issue_peak(self.handle, slot, C.byref(v)))
        return float(v.value)

    # ---- host buffers -------------------------------------------------------------------------------------------
    def encode_bc7(self, img: np.ndarray, params: Bc7Params | None = None, out: np.ndarray | None = None) -> np.ndarray:
        """img: (H, W, C) uint8, H and W multiples of 4, C in {3, 4} -> (H/4 * W/4, 16) uint8 BC7 blocks."""
        img = np.ascontiguousarray(img, dtype=np.uint8)
        h, w, c = img.shape
        if out is None:
            out = np.empty(((h // 4) * (w // 4), 16), dtype=np.uint8)
        pp = C.byref(params) if params is not None else None
        self._check(self.lib.vkt_bcn_cuda_encode_bc7(self.handle, _ptr(img), w, h, c, 0, pp, _ptr(out)))
        return out

    def encode_bc5(self, img: np.ndarray, out: np.ndarray | None = None) -> np.ndarray:
        img = np.ascontiguousarray(img, dtype=np.uint8)
        h, w, c = img.shape
        if out is None:
            out = np.empty(((h // 4) * (w // 4), 16), dtype=np.uint8)
        self._check(self.lib.vkt_bcn_cuda_encode_bc5(self.handle, _ptr(img), w, h, c, 0, _ptr(out)))
        return out

    def encode_batch(self, mode: int, images: list, outs: list, params: Bc7Params | None = None) -> None:
        """images: list of (H, W, C) uint8 arrays or (ptr, w, h, c) tuples; outs: matching output arrays / pointers."""
        arr = (Image * len(images))()
        keep = []
        for i, (im, o) in enumerate(zip(images, outs)):
            if isinstance(im, tuple):
                ptr, w, h, c = im
            else:
                im = np.ascontiguousarray(im, dtype=np.uint8)
                keep.append(im)
                h, w, c = im.shape
                ptr = _ptr(im)
            arr[i] = Image(ptr, w, h, c, 0, _ptr(o))
        pp = C.byref(params) if params is not None else None
        self._check(self.lib.vkt_bcn_cuda_encode_batch(self.handle, mode, arr, len(images), pp))

    def resize_u8(self, img: np.ndarray, ow: int, oh: int) -> np.ndarray:
        img = np.ascontiguousarray(img, dtype=np.uint8)
        h, w, c = img.shape
        out = np.empty((oh, ow, c), dtype=np.uint8)
        self._check(self.lib.vkt_bcn_cuda_resize_u8(self.handle, _ptr(img), w, h, c, _ptr(out), ow, oh))
        return out

    def compress(self, img: np.ndarray, mode: int = MODE_BC7, generate_mipmaps: bool = False,
                 params: Bc7Params | None = None) -> tuple[Plan, list[np.ndarray]]:
        """Whole vierkant::bcn::compress() chain on the GPU (resize + mips + encode)."""
        img = np.ascontiguousarray(img, dtype=np.uint8)
        h, w, c = img.shape
        plan = compress_plan(w, h, generate_mipmaps)
        levels = [np.empty((int(plan.level_num_blocks[l]), 16), dtype=np.uint8) for l in range(plan.num_levels)]
        ptrs = (C.c_void_p * plan.num_levels)(*[_ptr(l) for l in levels])
        pp = C.byref(params) if params is not None else None
        self._check(self.lib.vkt_bcn_cuda_compress(self.handle, mode, _ptr(img), w, h, c, int(generate_mipmaps), pp, ptrs))
        return plan, levels

    def compress_alloc(self, img: np.ndarray, mode: int = MODE_BC7, generate_mipmaps: bool = False,
                       params: Bc7Params | None = None) -> list[np.ndarray]:
        """vkt_bcn_cuda_compress_alloc: the level arrays are allocated by a callback while the GPU already works."""
        img = np.ascontiguousarray(img, dtype=np.uint8)
        h, w, c = img.shape
        levels = {}

        def alloc(_user, level, nbytes):
            levels[level] = np.zeros((nbytes // 16, 16), dtype=np.uint8)
            return levels[level].ctypes.data

        cb = ALLOC_FN(alloc)
        pp = C.byref(params) if params is not None else None
        self._check(self.lib.vkt_bcn_cuda_compress_alloc(self.handle, mode, _ptr(img), w, h, c, int(generate_mipmaps), pp, cb, None))
        return [levels[l] for l in sorted(levels)]

    def compress_shard_begin(self, mode: int, pixels, width: int, height: int, comps: int, generate_mipmaps: bool,
                             params: Bc7Params | None, rank: int, world: int, level_ptrs, handover) -> None:
        """Worker `rank` of `world`: queue this GPU's row slices of ONE chain (vkt_bcn_cuda_compress_shard_begin).  `level_ptrs` is
        a (c_void_p * levels) array of the (shared) level buffers, `handover` the shared hand-over buffer (or None)."""
        pp = C.byref(params) if params is not None else None
        self._check(self.lib.vkt_bcn_cuda_compress_shard_begin(self.handle, mode, _ptr(pixels), width, height, comps, int(generate_mipmaps),
                                                               pp, rank, world, level_ptrs, None if handover is None else _ptr(handover)))

    def compress_shard_end(self, mode: int, pixels, width: int, height: int, comps: int, generate_mipmaps: bool,
                           params: Bc7Params | None, rank: int, world: int, level_ptrs, handover) -> None:
        pp = C.byref(params) if params is not None else None
        self._check(self.lib.vkt_bcn_cuda_compress_shard_end(self.handle, mode, _ptr(pixels), width, height, comps, int(generate_mipmaps),
                                                             pp, rank, world, level_ptrs, None if handover is None else _ptr(handover)))

    def import_external_fd(self, fd: int, nbytes: int, slot: int = 0) -> tuple[int, int]:
        """Map external memory exported as an opaque fd (a Vulkan staging buffer, VK_KHR_external_memory_fd) into device `slot`.
        Returns (device pointer, handle); CUDA owns the fd afterwards.  Release with release_external()."""
        d_ptr, handle = C.c_void_p(), C.c_void_p()
        self._check(self.lib.vkt_bcn_cuda_import_external_fd(self.handle, slot, fd, nbytes, C.byref(d_ptr), C.byref(handle)))
        return int(d_ptr.value), int(handle.value)

    def release_external(self, d_ptr: int, handle: int, slot: int = 0) -> None:
        self._check(self.lib.vkt_bcn_cuda_release_external(self.handle, slot, d_ptr, handle))

    def host_register(self, buf) -> None:
        """cudaHostRegister (portable) of a numpy array / torch tensor the caller keeps alive."""
        n = buf.nbytes if hasattr(buf, "nbytes") else buf.numel() * buf.element_size()
        self._check(self.lib.vkt_bcn_cuda_host_register(self.handle, _ptr(buf), n))

    def host_unregister(self, buf) -> None:
        self._check(self.lib.vkt_bcn_cuda_host_unregister(self.handle, _ptr(buf)))

    def compress_batch(self, imgs: list, modes: int | list = MODE_BC7, generate_mipmaps: bool = True,
                       params: Bc7Params | None = None) -> list[list[np.ndarray]]:
        """vkt_bcn_cuda_compress_batch: every texture's whole chain, textures pipelined over two lanes per device."""
        imgs = [np.ascontiguousarray(im, dtype=np.uint8) for im in imgs]
        modes = [modes] * len(imgs) if isinstance(modes, int) else list(modes)
        srcs = (Source * len(imgs))()
        outs, keep = [], []
        for i, im in enumerate(imgs):
            h, w, c = im.shape
            plan = compress_plan(w, h, generate_mipmaps)
            levels = [np.empty((int(plan.level_num_blocks[l]), 16), dtype=np.uint8) for l in range(plan.num_levels)]
            ptrs = (C.c_void_p * plan.num_levels)(*[_ptr(l) for l in levels])
            keep.append(ptrs)
            srcs[i] = Source(_ptr(im), w, h, c, modes[i], ptrs)
            outs.append(levels)
        pp = C.byref(params) if params is not None else None
        self._check(self.lib.vkt_bcn_cuda_compress_batch(self.handle, srcs, len(imgs), int(generate_mipmaps), pp))
        return outs

    # ---- device buffers (kernel only) ---------------------------------------------------------------------------
    def encode_bc7_device(self, d_pixels, width: int, height: int, comps: int, d_out, params: Bc7Params | None = None,
                          slot: int = 0, stream: int | None = None, row_stride: int = 0) -> None:
        pp = C.byref(params) if params is not None else None
        self._check(self.lib.vkt_bcn_cuda_encode_bc7_device(self.handle, slot, _ptr(d_pixels), width, height, comps,
                                                            row_stride, pp, _ptr(d_out), stream))

    def make_device_batch(self, images: list, outs: list):
        """(d_pixels, w, h, comps) tuples + device outputs -> a reusable vkt_bcn_image array for encode_batch_device."""
        arr = (Image * len(images))()
        for i, ((px, w, h, c), o) in enumerate(zip(images, outs)):
            arr[i] = Image(_ptr(px), w, h, c, 0, _ptr(o))
        return arr

    def encode_batch_device(self, mode: int, batch, params: Bc7Params | None = None, slot: int = 0, stream: int | None = None) -> None:
        pp = C.byref(params) if params is not None else None
        self._check(self.lib.vkt_bcn_cuda_encode_batch_device(self.handle, slot, mode, batch, len(batch), pp, stream))

    def encode_bc5_device(self, d_pixels, width: int, height: int, comps: int, d_out, slot: int = 0,
                          stream: int | None = None, row_stride: int = 0) -> None:
        self._check(self.lib.vkt_bcn_cuda_encode_bc5_device(self.handle, slot, _ptr(d_pixels), width, height, comps,
                                                            row_stride, _ptr(d_out), stream))
