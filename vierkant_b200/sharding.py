"""Host-side shard plan of the BC7 path: independent units, no collective.

The same rule the C++ shim applies inside one multi-device context (vierkant_b200/csrc/bcn_cuda.cu,
vkt_bcn_cuda_encode_batch / resize_core.cuh compress_chain), stated once in Python so that one-process-per-GPU drivers
(bench.py under torchrun, the gloo tests) shard identically: every level's block rows are split evenly over the G
workers; levels with fewer than 4 * G block rows go to a single worker, rotating by level.  No compute happens here.
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


def level_rows(level: int, block_rows: int, rank: int, world: int) -> tuple[int, int]:
    """Block-row range [r0, r1) of `level` (which has `block_rows` rows of blocks) that worker `rank` of `world` encodes."""
    if block_rows < 4 * world:
        return (0, block_rows) if level % world == rank else (block_rows, block_rows)
    return block_rows * rank // world, block_rows * (rank + 1) // world


def shard_plan(width: int, height: int, rank: int, world: int, mipmaps: bool = True) -> list[dict]:
    """Per level: size, this worker's block rows and the byte range of its blocks inside the level's block array."""
    plan = []
    for l, (w, h) in enumerate(chain_dims(width, height, mipmaps)):
        r0, r1 = level_rows(l, h // 4, rank, world)
        plan.append({"level": l, "width": w, "height": h, "rows": (r0, r1), "block_range": (r0 * (w // 4), r1 * (w // 4))})
    return plan
