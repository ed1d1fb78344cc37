"""Host-side shard plan of the BC7 path: independent units, no collective.

The same rule the C++ shim applies inside one multi-device context (vierkant_b200/csrc/chain_plan.h, used by
compress_chain in resize_core.cuh), stated once in Python so that one-process-per-GPU drivers (the gloo tests, a
torchrun driver that wants to split ONE chain over ranks) shard identically: the leading levels -- those with at least
4 block rows and 128 pixel rows per worker -- are split evenly by block rows; the small levels after them all go to
worker 0 (it continues the chain from the last split level); an image too small to split at all is worker 0's alone.
No compute happens here.
"""
from __future__ import annotations


def round4(v: int) -> int:
    return (v + 3) & ~3


def chain_dims(width: int, height: int, mipmaps: bool = True) -> list[tuple[int, int]]:
    """Level sizes of vierkant::bcn::compress (texture_block_compression.cpp:80-86,141-146)."""
    w, h = round4(width), round4(height)
    levels = max(0, max(w, h).bit_length() - 1 - 2) + 1 if mipmaps else 1
    out = []
    for _ in range(levels):
        out.append((w, h))
        w, h = round4(max(w // 2, 1)), round4(max(h // 2, 1))
    return out


def chain_split(heights: list[int], world: int) -> tuple[int, int]:
    """(M, workers): levels [0, M) are split by block rows over `workers` workers, levels [M, L) belong to worker 0.
    Mirrors vkt::chain_split (chain_plan.h)."""
    if world <= 1:
        return len(heights), 1
    m = 0
    while m < len(heights) and heights[m] // 4 >= 4 * world and heights[m] >= 128 * world:
        m += 1
    return (len(heights), 1) if m == 0 else (m, world)


def level_rows(level: int, block_rows: int, rank: int, world: int, split: tuple[int, int] | None = None) -> tuple[int, int]:
    """Block-row range [r0, r1) of `level` (which has `block_rows` rows of blocks) that worker `rank` of `world` encodes.
    `split` = chain_split(...) of the chain the level belongs to (default: the level is judged on its own)."""
    m, workers = split if split is not None else chain_split([block_rows * 4], world)
    if split is None:
        m = 1 if workers > 1 else 0
        level = 0
    if workers == 1 or level >= m:
        return (0, block_rows) if rank == 0 else (block_rows, block_rows)
    return block_rows * rank // workers, block_rows * (rank + 1) // workers


def shard_plan(width: int, height: int, rank: int, world: int, mipmaps: bool = True) -> list[dict]:
    """Per level: size, this worker's block rows and the byte range of its blocks inside the level's block array."""
    dims = chain_dims(width, height, mipmaps)
    split = chain_split([h for _, h in dims], world)
    plan = []
    for l, (w, h) in enumerate(dims):
        r0, r1 = level_rows(l, h // 4, rank, world, split)
        plan.append({"level": l, "width": w, "height": h, "rows": (r0, r1), "block_range": (r0 * (w // 4), r1 * (w // 4)),
                     "sliced": l < split[0] and split[1] > 1})
    return plan
