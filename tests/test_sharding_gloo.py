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
    plans = [sharding.shard_plan(w, h, r, world) for r in range(world)]
    for l, (lw, lh) in enumerate(dims):
        rows = lh // 4
        seen = np.zeros(rows, dtype=np.int32)
        for r in range(world):
            r0, r1 = plans[r][l]["rows"]
            seen[r0:r1] += 1
        assert (seen == 1).all(), (l, lw, lh)


@pytest.mark.parametrize("w,h,world", [(16384, 16384, 8), (8192, 8192, 8), (8192, 8192, 2), (4096, 4096, 4), (1000, 520, 2), (64, 64, 2),
                                        (16384, 16384, 1), (2050, 2046, 3), (4, 4, 8)])
def test_c_abi_shard_plan_matches_the_python_mirror(cuda_lib, w, h, world):
    """vkt_bcn_cuda_compress_shard_plan / _rows (what a process-per-GPU driver asks the library) == sharding.py."""
    from vierkant_b200 import capi
    dims = sharding.chain_dims(w, h)
    m, workers = sharding.chain_split([lh for _, lh in dims], world)
    sp = capi.shard_plan(w, h, True, world)
    assert (sp.num_levels, sp.workers) == (len(dims), workers)
    assert sp.sliced_levels == (m if workers > 1 else 0)
    tail = workers > 1 and m < len(dims)
    assert sp.handover_bytes == (dims[m - 1][0] * dims[m - 1][1] * 4 if tail else 0)
    for r in range(world):
        for l, e in enumerate(sharding.shard_plan(w, h, r, world)):
            assert capi.shard_rows(w, h, True, r, world, l) == e["rows"]


@pytest.fixture(scope="module")
def emul_lib():
    import ctypes as C
    import subprocess
    here = os.path.dirname(os.path.abspath(__file__))
    subprocess.run(["make", "-s", "-C", os.path.join(here, "host_emul")], check=True)
    lib = C.CDLL(os.path.join(here, "host_emul", "libvkt_emul.so"))
    u32p = C.POINTER(C.c_uint32)
    lib.emul_chain_plan.argtypes = [u32p, C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint32, u32p, u32p, u32p, u32p, u32p, u32p]
    return lib


@pytest.mark.parametrize("w,h,world", [(4096, 4096, 8), (4096, 4096, 2), (8192, 8192, 8), (1000, 520, 2), (2050, 2046, 4), (1024, 4096, 3),
                                        (16384, 16384, 8), (512, 512, 2), (260, 12, 3), (123, 81, 2), (2048, 2048, 5)])
def test_multi_device_chain_plan(emul_lib, w, h, world):
    """The C++ plan compress() uses on several devices (chain_plan.h, with the real stbir tap ranges): the Python mirror
    agrees on the split; own rows partition every sliced level; the rows a device produces at level l cover every tap
    of the rows it produces at level l + 1; the source rows it uploads shrink towards 1 / G of the image."""
    import ctypes as C
    dims = sharding.chain_dims(w, h)
    heights = (C.c_uint32 * len(dims))(*[lh for _, lh in dims])
    m_py, workers_py = sharding.chain_split([lh for _, lh in dims], world)
    covered = [np.zeros(lh // 4, dtype=np.int32) for _, lh in dims]
    uploads = []
    for g in range(world):
        arr = [(C.c_uint32 * 16)() for _ in range(4)]
        s0, s1 = C.c_uint32(), C.c_uint32()
        rc = emul_lib.emul_chain_plan(heights, len(dims), h, world, g, arr[0], arr[1], arr[2], arr[3], C.byref(s0), C.byref(s1))
        assert rc >= 0, "need[] does not cover the taps of the next level"
        m, workers = rc & 255, rc >> 8
        assert (m, workers) == (m_py, workers_py)
        if g >= workers:
            continue
        py = sharding.shard_plan(w, h, g, world)
        for l in range(m):
            own, need = (arr[0][l], arr[1][l]), (arr[2][l], arr[3][l])
            assert py[l]["rows"] == own
            assert need[0] <= own[0] * 4 and need[1] >= own[1] * 4 and need[1] <= dims[l][1]
            covered[l][own[0]:own[1]] += 1
        uploads.append(s1.value - s0.value)
    for l in range(m_py):
        assert (covered[l] == 1).all()
    if workers_py > 1:
        assert max(uploads) <= h / workers_py * 1.45 + 16  # own slice + halo, not the whole image


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
        img = synth.make_texture(256, 512, 1, seed=5)  # levels 512 and 256 rows are split over the 2 workers, the rest is worker 0's
        # every worker derives the level images itself (the cheap part), encodes only its rows of every level
        pieces, prev = [], img
        for p in sharding.shard_plan(256, 512, rank, world):
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


def _shared_chain_worker(rank, world, port, q):
    """One worker of the strong-scaling driver's host side (bench_strong.SharedChain + hostshare.FlagBarrier), on CPU: the
    oracle stands in for the GPU, everything else -- shared source filled by rows, shared level arrays written in place at
    the C-ABI shard rows, hand-over of the last sliced level, flag barrier -- is the code the bench runs."""
    import torch.distributed as dist
    import bench_strong
    from oracle import pyoracle
    from vierkant_b200 import capi
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        W, H = 256, 1024  # levels 1024, 512 and 256 rows are sliced over two workers, the tail is worker 0's
        sc = bench_strong.SharedChain(f"vkt_test_{port}", W, H, world, rank, dist.barrier)
        try:
            synth.fill_shared(sc.bufs["src"].path, 0, W, H, 1, 77, (H * rank // world, H * (rank + 1) // world), procs=2)
            dist.barrier()
            oracle = pyoracle.PortOracle()
            M = int(sc.sp.sliced_levels)
            assert M >= 2 and int(sc.sp.workers) == world
            # every worker derives the level images (the cheap part) and writes ONLY its own block rows of the sliced levels
            px, prev = [], np.ascontiguousarray(sc.src)
            for (lw, lh) in sc.level_dims:
                prev = oracle.resize(prev, lw, lh)
                px.append(prev)
            for l in range(M):
                r0, r1 = capi.shard_rows(W, H, True, rank, world, l)
                bx = sc.level_dims[l][0] // 4
                sc.levels[l][r0 * bx:r1 * bx] = oracle.encode_blocks(synth.to_blocks(np.ascontiguousarray(px[l][4 * r0:4 * r1])))
            # hand-over: this worker's pixel rows of the last sliced level, then the flag barrier (no data moves through it)
            lw, lh = sc.level_dims[M - 1]
            r0, r1 = capi.shard_rows(W, H, True, rank, world, M - 1)
            hand = sc.handover[:lw * lh * 4].reshape(lh, lw, 4)
            hand[4 * r0:4 * r1] = px[M - 1][4 * r0:4 * r1]
            sc.barrier.arrive(1)
            if rank == 0:
                sc.barrier.wait_all(1, timeout_s=60)
                prev = np.ascontiguousarray(hand)  # every worker's rows are in place now
                for l in range(M, sc.L):
                    prev = oracle.resize(prev, *sc.level_dims[l])
                    sc.levels[l][:] = oracle.encode_blocks(synth.to_blocks(prev))
            sc.barrier.arrive(2)
            sc.barrier.wait_all(2, timeout_s=60)
            if rank == 0:
                want = oracle.compress(synth.make_texture(W, H, 1, seed=77), 1, True)
                ok = all(np.array_equal(a, b) for a, b in zip(sc.levels, want["levels"]))
                q.put(bool(ok))
            dist.barrier()
        finally:
            sc.close()
    finally:
        dist.destroy_process_group()


def test_shared_chain_and_flag_barrier_gather_one_chain(port_oracle, cuda_lib):
    """world_size 2, gloo: the shared-memory gather of bench_strong (one chain, block rows written in place by their owners,
    tail levels by worker 0 after the flag barrier) reproduces the single-worker chain byte for byte."""
    import torch.multiprocessing as mp
    world = 2
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_shared_chain_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(180)
        assert p.exitcode == 0
    assert q.get(timeout=5) is True


def test_fill_shared_matches_make_texture(tmp_path):
    from vierkant_b200 import hostshare
    buf = hostshare.SharedBuffer(f"vkt_test_fill_{os.getpid()}", 64 * 200 * 4, True)
    try:
        synth.fill_shared(buf.path, 0, 64, 200, 1, 9, (0, 120), procs=3)
        synth.fill_shared(buf.path, 0, 64, 200, 1, 9, (120, 200), procs=1)
        assert np.array_equal(buf.array.reshape(200, 64, 4), synth.make_texture(64, 200, 1, seed=9))
    finally:
        buf.close()
