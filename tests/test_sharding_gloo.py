"""Multi-worker host logic on CPU (gloo, world_size 2): the shard plan covers every block exactly once, sharded
encoding + gather reproduces the single-worker result bit for bit, and the max-over-ranks timing reduction works.
The block encoder used here is the oracle (this is a test of the host-side plan, not of the CUDA path)."""
import os
import socket

import numpy as np
import pytest

from vierkant_b200 import sharding, synth


@pytest.mark.parametrize("w,h,world", [(4096, 4096, 8), (1024, 512, 2), (124, 84, 4), (4, 4, 8), (16384, 16384, 8), (260, 12, 3)])
def test_plan_covers_every_block_once(w, h, world):
    dims = sharding.chain_dims(w, h)
    for l, (lw, lh) in enumerate(dims):
        rows = lh // 4
        seen = np.zeros(rows, dtype=np.int32)
        for r in range(world):
            r0, r1 = sharding.level_rows(l, rows, r, world)
            seen[r0:r1] += 1
        assert (seen == 1).all(), (l, lw, lh)


def test_chain_dims_match_the_reference_formula():
    for w, h in [(512, 256), (123, 81), (64, 128), (4096, 4096), (5, 3), (1, 1)]:
        w4, h4 = sharding.round4(w), sharding.round4(h)
        n = max(0, int(np.log2(max(w4, h4)) - 2)) + 1  # texture_block_compression.cpp:84-85
        dims = sharding.chain_dims(w, h)
        assert len(dims) == n and dims[0] == (w4, h4) and dims[-1][0] >= 4 and dims[-1][1] >= 4


def _worker(rank, world, port, q):
    import torch
    import torch.distributed as dist
    from oracle import pyoracle
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        oracle = pyoracle.PortOracle()
        img = synth.make_texture(128, 64, 1, seed=5)
        # every worker derives the level images itself (the cheap part), encodes only its rows of every level
        pieces, prev = [], img
        for p in sharding.shard_plan(128, 64, rank, world):
            prev = oracle.resize(prev, p["width"], p["height"])
            r0, r1 = p["rows"]
            if r1 > r0:
                tiles = synth.to_blocks(np.ascontiguousarray(prev[4 * r0:4 * r1]))
                pieces.append((p["level"], p["block_range"], oracle.encode_blocks(tiles)))
        gathered = [None] * world
        dist.all_gather_object(gathered, pieces)
        t = torch.tensor([1.0 + rank], dtype=torch.float64)  # stand-in for this rank's device time
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        if rank == 0:
            want = oracle.compress(img, 1, True)
            levels = [np.zeros_like(l) for l in want["levels"]]
            filled = [np.zeros(l.shape[0], dtype=np.int32) for l in levels]
            for plist in gathered:
                for level, (b0, b1), blocks in plist:
                    levels[level][b0:b1] = blocks
                    filled[level][b0:b1] += 1
            ok = all((f == 1).all() for f in filled) and all(np.array_equal(a, b) for a, b in zip(levels, want["levels"]))
            q.put((ok, float(t.item())))
    finally:
        dist.destroy_process_group()


def test_sharded_chain_gathers_to_the_single_worker_result(port_oracle):
    import torch.multiprocessing as mp
    world = 2
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(120)
        assert p.exitcode == 0
    ok, tmax = q.get(timeout=5)
    assert ok
    assert tmax == float(world)
