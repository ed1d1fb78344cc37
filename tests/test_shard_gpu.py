"""ONE chain split over several single-GPU workers (vkt_bcn_cuda_compress_shard_begin / _end), the process-per-GPU form of
SURVEY.md 8e.  The workers of a plan are independent contexts, so a one-GPU box can run all of them one after the other
on its only device: begin for every worker, then end for every worker (the "barrier" in between is the loop itself).  The
union of what they write into the shared level buffers must be byte-identical to vkt_bcn_cuda_compress on one device and
to the oracle -- this is the row-slicing / halo-recompute / hand-over path that `bench.py --gpus N` times as `strong`."""
import ctypes as C
import os

import numpy as np
import pytest

from vierkant_b200 import capi, synth

pytestmark = pytest.mark.gpu


def run_sharded(img, world, mode=capi.MODE_BC7, mips=True, params=None, device_buffers=False, pinned=False):
    """All `world` workers on device 0, sharing one set of level buffers.  Returns the list of level block arrays."""
    import torch
    h, w, c = img.shape
    plan = capi.compress_plan(w, h, mips)
    sp = capi.shard_plan(w, h, mips, world)
    sizes = [int(plan.level_num_blocks[l]) for l in range(plan.num_levels)]
    keep = []
    if device_buffers:
        src = torch.from_numpy(img).cuda()
        levels = [torch.full((n, 16), 0xAB, dtype=torch.uint8, device="cuda") for n in sizes]
        torch.cuda.synchronize()
    elif pinned:
        src = torch.from_numpy(img).pin_memory()
        levels = [torch.full((n, 16), 0xAB, dtype=torch.uint8).pin_memory() for n in sizes]
    else:
        src = np.ascontiguousarray(img)
        levels = [np.full((n, 16), 0xAB, dtype=np.uint8) for n in sizes]
    handover = np.zeros(max(int(sp.handover_bytes), 1), dtype=np.uint8)
    ptrs = (C.c_void_p * len(levels))(*[capi._ptr(l) for l in levels])
    ctxs = [capi.BcnContext([0]) for _ in range(world)]
    try:
        for r, cx in enumerate(ctxs):
            cx.compress_shard_begin(mode, src, w, h, c, mips, params, r, world, ptrs, handover)
        for r, cx in enumerate(ctxs):
            cx.compress_shard_end(mode, src, w, h, c, mips, params, r, world, ptrs, handover)
    finally:
        for cx in ctxs:
            cx.close()
    if device_buffers or pinned:
        return [l.cpu().numpy() for l in levels]
    return levels


@pytest.mark.parametrize("w,h,world", [(1024, 1024, 2), (1024, 1024, 4), (1000, 520, 2), (2048, 1024, 8), (64, 64, 2), (512, 2048, 3)])
def test_shard_union_matches_single_device(ctx, w, h, world):
    img = synth.make_texture(w, h, 1, seed=w + 3 * h + world)
    _, want = ctx.compress(img, capi.MODE_BC7, True)
    got = run_sharded(img, world)
    assert len(got) == len(want)
    for l, (a, b) in enumerate(zip(got, want)):
        assert np.array_equal(a, b), f"level {l}: {(a != b).any(axis=1).sum()} of {len(b)} blocks differ"


def test_shard_matches_oracle_and_known_hash(ctx, port_oracle):
    """The same 2048^2 chain the reference's compress() hash is known for (tests/golden/known_answers.json), over 4 workers."""
    import json
    with open(os.path.join(os.path.dirname(__file__), "golden", "known_answers.json")) as f:
        known = [e for e in json.load(f)["compress"] if e["size"] == 2048 and e["kind"] == 0]
    img = synth.make_texture(2048, 2048, 0)
    got = run_sharded(img, 4, pinned=True)
    if known:
        assert "%016x" % synth.fnv1a64_words(np.concatenate(got)) == known[0]["fnv1a64"]
    small = synth.make_texture(640, 512, 1, seed=5)
    want = port_oracle.compress(small, 1, True, threads=os.cpu_count() or 1)
    for a, b in zip(run_sharded(small, 2), want["levels"]):
        assert np.array_equal(a, b)


def test_shard_device_resident_source_and_destinations(ctx):
    """Source and level buffers in HBM (the `strong.value` arm of bench.py): the source is read in place."""
    img = synth.make_texture(1536, 1024, 1, seed=21)
    _, want = ctx.compress(img, capi.MODE_BC7, True)
    for a, b in zip(run_sharded(img, 4, device_buffers=True), want):
        assert np.array_equal(a, b)


def test_shard_with_parameters_and_bc5(ctx):
    """C3 / C5 style parameters through the sharded path, and the BC5 branch."""
    img = synth.make_texture(1024, 512, 1, seed=8)
    for kw in (dict(mode17_partition_estimation_filterbank=0), dict(uber_level=4, mode17_partition_estimation_filterbank=0)):
        p = capi.default_params(**kw)
        _, want = ctx.compress(img, capi.MODE_BC7, True, p)
        for a, b in zip(run_sharded(img, 2, params=p), want):
            assert np.array_equal(a, b)
    _, want = ctx.compress(img, capi.MODE_BC5, True)
    for a, b in zip(run_sharded(img, 2, mode=capi.MODE_BC5), want):
        assert np.array_equal(a, b)


def test_shard_workers_write_disjoint_rows(ctx):
    """A worker touches only its own block rows (the rest of the shared buffer keeps its fill pattern)."""
    img = synth.make_texture(1024, 1024, 0, seed=2)
    h, w, c = img.shape
    world, rank = 4, 2
    plan = capi.compress_plan(w, h, True)
    sp = capi.shard_plan(w, h, True, world)
    levels = [np.full((int(plan.level_num_blocks[l]), 16), 0xCD, dtype=np.uint8) for l in range(plan.num_levels)]
    handover = np.zeros(int(sp.handover_bytes), dtype=np.uint8)
    ptrs = (C.c_void_p * len(levels))(*[l.ctypes.data for l in levels])
    with capi.BcnContext([0]) as cx:
        cx.compress_shard_begin(capi.MODE_BC7, img, w, h, c, True, None, rank, world, ptrs, handover)
        cx.compress_shard_end(capi.MODE_BC7, img, w, h, c, True, None, rank, world, ptrs, handover)
    _, want = ctx.compress(img, capi.MODE_BC7, True)
    for l in range(plan.num_levels):
        r0, r1 = capi.shard_rows(w, h, True, rank, world, l)
        bx = int(plan.level_width[l]) // 4
        assert np.array_equal(levels[l][r0 * bx:r1 * bx], want[l][r0 * bx:r1 * bx])
        untouched = np.concatenate([levels[l][:r0 * bx], levels[l][r1 * bx:]])
        assert (untouched == 0xCD).all()
    assert handover.any()  # rank 2 left its rows of the last sliced level there


def test_shard_argument_errors(ctx):
    img = synth.make_texture(64, 64, 0)
    out = np.zeros((256, 16), dtype=np.uint8)
    ptrs = (C.c_void_p * 1)(out.ctypes.data)
    with capi.BcnContext([0]) as cx:
        with pytest.raises(capi.BcnError):
            cx.compress_shard_end(capi.MODE_BC7, img, 64, 64, 4, False, None, 0, 2, ptrs, None)  # no begin
        with pytest.raises(capi.BcnError):
            cx.compress_shard_begin(capi.MODE_BC7, img, 64, 64, 4, False, None, 2, 2, ptrs, None)  # rank >= world
        cx.compress_shard_begin(capi.MODE_BC7, img, 64, 64, 4, False, None, 0, 2, ptrs, None)
        with pytest.raises(capi.BcnError):
            cx.compress_shard_begin(capi.MODE_BC7, img, 64, 64, 4, False, None, 0, 2, ptrs, None)  # begin twice
        cx.compress_shard_end(capi.MODE_BC7, img, 64, 64, 4, False, None, 0, 2, ptrs, None)
    assert np.array_equal(out, ctx.encode_bc7(ctx.resize_u8(img, 64, 64)))


def test_axis_cache_survives_more_sizes_than_it_holds(ctx, port_oracle):
    """More than 256 distinct (in, out) resize axes through one context: entries are evicted, never while in use
    (the old cache cleared itself on a miss and left dangling tap tables behind)."""
    rng = np.random.default_rng(5)
    img = synth.make_texture(96, 64, 1, seed=4)
    for i in range(150):
        ow, oh = 4 + i, 200 - i  # 300 new axes
        got = ctx.resize_u8(img, ow, oh)
        if i % 25 == 0:
            assert np.array_equal(got, port_oracle.resize(img, ow, oh))
    small = synth.make_texture(250, 130, 1, seed=6)  # a chain right after the churn: its axes are fetched one by one
    want = port_oracle.compress(small, 1, True, threads=os.cpu_count() or 1)
    _, levels = ctx.compress(small, capi.MODE_BC7, True)
    for a, b in zip(levels, want["levels"]):
        assert np.array_equal(a, b)
