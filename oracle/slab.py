"""Reference pixels of a row slab of a deep mip level without filtering the whole image.  TEST INFRASTRUCTURE ONLY
(same rule as pyoracle.py: only tests/, smoke() and bench.py's checking / cpu_baseline legs import this).

vierkant::bcn::compress() derives level l from level l-1 with stbir (src/texture_block_compression.cpp:99-101), whole image
by whole image, single-threaded: 16384^2 takes the CPU ~14 s before a single block is encoded.  For power-of-two textures
the filter is shift-invariant along y -- at 1:1 and 2:1 the sample positions are exact in float, so the coefficients of an
output row depend only on its distance to the image edge -- and the reference's own resize of a band of rows [a, b) gives
bit-identical rows to its resize of the whole image once the band reaches 8 (level 0) / 16 (2:1) input rows past the rows
asked for, or ends at the image edge (tests/test_oracle_pinning.py::test_slab_chain_equals_whole_image_resize pins that
against oracle/_ref).  So the filtered rows [y0, y1) of level L cost a chain of L+1 small resizes.
"""
from __future__ import annotations

import numpy as np


def level_rows(oracle, src_rows, width: int, height: int, level: int, y0: int, y1: int) -> np.ndarray:
    """Rows [y0, y1) of mip level `level` (level 0 = the 1:1 Mitchell pass) of a width x height RGBA8 power-of-two texture,
    filtered by `oracle.resize` (RefOracle or PortOracle).  src_rows(a, b) returns source rows [a, b) as an (b-a, W, C) array."""
    assert width & (width - 1) == 0 and height & (height - 1) == 0, "power-of-two textures only"
    # row ranges needed per level, from the target level down to the source
    need = [None] * (level + 1)
    need[level] = (y0, y1)
    for l in range(level, 0, -1):
        h_prev = height >> (l - 1)
        a, b = need[l]
        need[l - 1] = (max(0, 2 * a - 16), min(h_prev, 2 * b + 16))
    a0, b0 = need[0]
    sa, sb = max(0, a0 - 8), min(height, b0 + 8)
    cur = oracle.resize(np.ascontiguousarray(src_rows(sa, sb)), width, sb - sa)  # level 0 rows [sa, sb); valid inside [a0, b0)
    cur_a = sa
    for l in range(1, level + 1):
        pa, pb = need[l - 1]
        band = np.ascontiguousarray(cur[pa - cur_a:pb - cur_a])
        w = width >> l
        cur = oracle.resize(band, w, (pb - pa) // 2)  # level l rows [pa / 2, pb / 2)
        cur_a = pa // 2
    return np.ascontiguousarray(cur[y0 - cur_a:y1 - cur_a])


def _bands(n_rows: int, parts: int, quantum: int = 4) -> list[tuple[int, int]]:
    """[0, n_rows) cut into at most `parts` bands whose edges are multiples of `quantum`."""
    q = (n_rows + quantum - 1) // quantum
    parts = max(1, min(parts, q))
    edges = sorted({min(n_rows, (q * k // parts) * quantum) for k in range(parts + 1)})
    return [(a, b) for a, b in zip(edges[:-1], edges[1:]) if b > a]


def resize_banded(oracle, img: np.ndarray, ow: int, oh: int, threads: int) -> np.ndarray:
    """oracle.resize(img, ow, oh) for the two cases vierkant::bcn::compress() produces on power-of-two textures (1:1 and
    2:1), computed as independent row bands on `threads` host threads (the ctypes call releases the GIL).  Byte-identical
    to the whole-image call -- see the module docstring; anything else falls through to the plain call."""
    h, w, c = img.shape
    pot = lambda v: v & (v - 1) == 0
    ratio = h // oh if oh and h % oh == 0 else 0
    if threads <= 1 or not (pot(h) and pot(w)) or ratio not in (1, 2) or w // ow != ratio or oh < 64:
        return oracle.resize(img, ow, oh)
    from concurrent.futures import ThreadPoolExecutor
    out = np.empty((oh, ow, c), dtype=np.uint8)
    halo = 8 if ratio == 1 else 16

    def job(band):
        a, b = band
        ia, ib = max(0, ratio * a - halo), min(h, ratio * b + halo)
        ia -= ia % ratio  # keep the band's phase: an even first input row at 2:1
        part = oracle.resize(np.ascontiguousarray(img[ia:ib]), ow, (ib - ia) // ratio)
        out[a:b] = part[a - ia // ratio:b - ia // ratio]

    with ThreadPoolExecutor(threads) as ex:
        list(ex.map(job, _bands(oh, threads * 2, 4)))
    return out


def chain_levels(oracle, img: np.ndarray, num_levels: int, threads: int) -> list[np.ndarray]:
    """The pixel images of every level of vierkant::bcn::compress() for a power-of-two RGBA8 texture (level 0 = the 1:1 pass,
    src/texture_block_compression.cpp:99-101,141-146), each level from the previous one, bands in parallel."""
    h, w, _ = img.shape
    levels, prev = [], img
    for l in range(num_levels):
        r4 = lambda v: (v + 3) & ~3
        lw, lh = (w, h) if l == 0 else (r4(max(levels[-1].shape[1] // 2, 1)), r4(max(levels[-1].shape[0] // 2, 1)))
        prev = resize_banded(oracle, prev, lw, lh, threads)
        levels.append(prev)
    return levels
