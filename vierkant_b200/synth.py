"""Deterministic synthetic RGBA8 textures for the BASELINE.json configs (SURVEY.md App. C generator spec).

Pure numpy; used by tests/ and bench.py to make inputs (this is input *specification*, not part of the encoder).
PRNG: s = s*1664525 + 1013904223 (mod 2^32), draw = s >> 24, four draws per pixel in row-major order
(R, G, B noise, then one alpha draw), so RGB is identical across kinds.
"""
from __future__ import annotations

import os

import numpy as np

_A = 1664525
_C = 1013904223
_M = 0xFFFFFFFF

FNV_OFFSET = 0xCBF29CE484222325
FNV_PRIME = 0x100000001B3


def _affine_powers(n: int):
    """(A_k, C_k) for k = 1..n such that state_k = A_k * s0 + C_k (mod 2^32)."""
    a = np.empty(n, dtype=np.uint64)
    c = np.empty(n, dtype=np.uint64)
    ak, ck = 1, 0
    for k in range(n):
        ak = (ak * _A) & _M
        ck = (ck * _A + _C) & _M
        a[k] = ak
        c[k] = ck
    return a, c


def _affine_pow(a: int, c: int, n: int) -> tuple[int, int]:
    """(a, c) composed n times: x -> a x + c applied n times (mod 2^32), by repeated squaring."""
    ra, rc = 1, 0
    while n:
        if n & 1:
            ra, rc = (a * ra) & _M, (a * rc + c) & _M
        a, c = (a * a) & _M, (a * c + c) & _M
        n >>= 1
    return ra, rc


def make_texture(width: int, height: int, kind: int = 0, seed: int | None = None, rows: tuple[int, int] | None = None,
                 out: np.ndarray | None = None) -> np.ndarray:
    """Return an (H, W, 4) uint8 texture.  kind 0 = opaque albedo-like, kind 1 = alpha gradients (right half).
    rows = (y0, y1): only those rows of the texture (the generator's state is jumped to row y0), returned as a
    (y1 - y0, W, 4) array or written into `out` -- how several processes fill one shared image."""
    if seed is None:
        seed = 0xB200 + kind
    W, H = int(width), int(height)
    a_k, c_k = _affine_powers(4 * W)
    # affine map for a whole row (4W draws), to jump from row start to row start
    row_a, row_c = int(a_k[-1]), int(c_k[-1])
    y0, y1 = rows if rows is not None else (0, H)
    full = out if out is not None else np.empty((y1 - y0, W, 4), dtype=np.uint8)
    out = _RowView(full, y0)
    x = np.arange(W, dtype=np.int64)
    ja, jc = _affine_pow(row_a, row_c, y0)
    s = (ja * (seed & _M) + jc) & _M
    for y in range(y0, y1):
        st = (a_k * np.uint64(s) + c_k) & np.uint64(_M)  # states after draws 1..4W
        draws = (st >> np.uint64(24)).astype(np.int64).reshape(W, 4)
        s = (row_a * s + row_c) & _M
        gr = x * 255 // max(W - 1, 1)
        gg = np.full(W, y * 255 // max(H - 1, 1), dtype=np.int64)
        gb = (x + y) * 255 // max(W + H - 2, 1)
        stripe = (((x + 2 * y) // 24) & 1).astype(bool)
        gr = np.where(stripe, 255 - gr, gr)
        gb = np.where(stripe, gb * 3 // 4, gb)
        out[y, :, 0] = np.clip(gr + draws[:, 0] % 25 - 12, 0, 255)
        out[y, :, 1] = np.clip(gg + draws[:, 1] % 25 - 12, 0, 255)
        out[y, :, 2] = np.clip(gb + draws[:, 2] % 25 - 12, 0, 255)
        if kind == 0:
            out[y, :, 3] = 255
        else:
            na = draws[:, 3]
            bx = x >> 2
            by = y >> 2
            grad = np.clip((2 * (x - W // 2) + y) * 255 // (W + H) + na % 9 - 4, 0, 255)
            a = np.where(((bx + by) & 7) == 0, (bx * 37 + by * 11) & 255, np.where(((bx ^ by) & 7) == 1, 0, grad))
            out[y, :, 3] = np.where(x < W // 2, 255, a)
    return full


class _RowView:
    """out[y, :, c] addressing of a row range stored from row y0 on."""

    def __init__(self, arr: np.ndarray, y0: int):
        self.arr, self.y0 = arr, y0

    def __setitem__(self, key, value):
        y, xs, c = key
        self.arr[y - self.y0, xs, c] = value


def to_blocks(img: np.ndarray) -> np.ndarray:
    """(H, W, 4) image with H, W multiples of 4 -> (H/4 * W/4, 16, 4) row-major 4x4 tiles (get_block order)."""
    H, W, C = img.shape
    assert H % 4 == 0 and W % 4 == 0 and C == 4
    return np.ascontiguousarray(img.reshape(H // 4, 4, W // 4, 4, 4).transpose(0, 2, 1, 3, 4).reshape(-1, 16, 4))


def fnv1a64_words(blocks: np.ndarray) -> int:
    """Word-wise FNV-1a over the uint64 view of BC blocks (SURVEY.md App. C hash)."""
    words = np.ascontiguousarray(blocks).view(np.uint64).ravel()
    h = FNV_OFFSET
    mask = (1 << 64) - 1
    for v in words.tolist():
        h = ((h ^ v) * FNV_PRIME) & mask
    return h


def mode_histogram(blocks: np.ndarray) -> dict:
    """BC7 mode = index of the lowest set bit of byte 0."""
    b0 = np.ascontiguousarray(blocks).view(np.uint8).reshape(-1, 16)[:, 0].astype(np.int64)
    low = b0 & -b0
    modes = np.where(low > 0, np.log2(np.maximum(low, 1)).astype(np.int64), 8)
    vals, counts = np.unique(modes, return_counts=True)
    return {int(v): int(c) for v, c in zip(vals, counts)}


def checkerboard_4x4(comps: int = 4) -> np.ndarray:
    """The 4x4 test image of the reference's tests/TestCompressionBC7.cpp:5-8, reinterpreted with `comps` channels
    exactly as the tests do (the same 64 bytes; with comps=3 only the first 48 are the image)."""
    words = np.array([0xFFFFFFFF, 0xFF000000, 0xFFFFFFFF, 0xFF000000,
                      0xFF000000, 0xFFFFFFFF, 0x00000000, 0xFFFFFFFF,
                      0xFFFFFFFF, 0x00000000, 0xFFFFFFFF, 0xFF000000,
                      0xFF000000, 0xFFFFFFFF, 0xFF000000, 0xFFFFFFFF], dtype="<u4")
    raw = words.view(np.uint8)
    return raw[: 16 * comps].reshape(4, 4, comps).copy()


def _fill_job(args):
    path, offset, W, H, kind, seed, a, b = args
    m = np.memmap(path, dtype=np.uint8, mode="r+", offset=offset + a * W * 4, shape=(b - a, W, 4))
    make_texture(W, H, kind, seed, rows=(a, b), out=m)
    m.flush()
    return b - a


def fill_shared(path: str, offset: int, width: int, height: int, kind: int, seed: int | None, rows: tuple[int, int],
                procs: int = 1) -> None:
    """Rows [rows[0], rows[1]) of make_texture(width, height, kind, seed) written into the file mapping `path` (a
    hostshare.SharedBuffer) at byte `offset` + row * width * 4, by `procs` child processes (the generator is GIL-bound
    numpy; the children are plain `python -m vierkant_b200.synth` runs, independent of the caller's __main__).  How the
    workers of a sharded job fill one shared source image, each its own rows."""
    import subprocess
    import sys
    y0, y1 = rows
    procs = max(1, min(procs, (y1 - y0 + 63) // 64))
    if seed is None:
        seed = 0xB200 + kind
    cuts = [y0 + (y1 - y0) * k // procs for k in range(procs + 1)]
    jobs = [(path, offset, width, height, kind, seed, a, b) for a, b in zip(cuts[:-1], cuts[1:]) if b > a]
    if procs == 1:
        for j in jobs:
            _fill_job(j)
        return
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    children = [subprocess.Popen([sys.executable, "-m", "vierkant_b200.synth", *map(str, j)], cwd=root) for j in jobs]
    bad = [c.args for c in children if c.wait(timeout=900) != 0]
    if bad:
        raise RuntimeError(f"fill_shared: {len(bad)} generator processes failed")


if __name__ == "__main__":  # python -m vierkant_b200.synth <path> <offset> <W> <H> <kind> <seed> <row0> <row1>
    import sys
    _a = sys.argv[1:]
    _fill_job((_a[0], *map(int, _a[1:8])))
