"""Every BASELINE.json config at FULL size through compress() (the reference-facing call: source in, stbir-exact chain + encode
on the GPU, every level's blocks out), against the unmodified reference (oracle/_ref) -- not against the C port:

  C1 / C2   1024^2 / 4096^2 opaque + mips, defaults        the reference's own vierkant::bcn::compress(), every block
  C3        8192^2 alpha gradients + mips, 64 partitions, filterbank off     reference resize chain + bc7enc_compress_block, every block
  C4        material batch (8 x 4096^2 per GPU, every 4th with alpha) through compress_batch: every texture, every block
  C5        16384^2 + mips, uber 4, filterbank off: a 2048-row slab of level 0, rows around every 1/8 boundary of levels 0..2,
            levels >= 3 complete (the CPU needs ~90 us per uber-4 block; BASELINE.md section 3 allows the slab)

vierkant::bcn::compress() has no parameter argument, so C3 / C5 take the reference's pixels (its own resize, row bands in
parallel -- oracle/slab.py, pinned to the whole-image call) and its bc7enc_compress_block with the config's parameters."""
import os

import numpy as np
import pytest

from vierkant_b200 import capi, synth

pytestmark = pytest.mark.gpu
THREADS = os.cpu_count() or 1


def _mismatches(got, want):
    return sum(int((np.asarray(a).reshape(-1, 16) != np.asarray(b).reshape(-1, 16)).any(axis=1).sum()) for a, b in zip(got, want))


@pytest.mark.parametrize("size", [1024, 4096])
def test_c1_c2_whole_chain_equals_reference_compress(ctx, ref_oracle, size):
    img = synth.make_texture(size, size, 0, seed=0xB200)
    want = ref_oracle.compress(img, 1, True, THREADS)["levels"]
    _, got = ctx.compress(img, capi.MODE_BC7, True)
    assert len(got) == len(want) and sum(a.shape[0] for a in got) == sum(a.shape[0] for a in want)
    assert _mismatches(got, want) == 0
    assert _mismatches(ctx.compress_alloc(img, capi.MODE_BC7, True), want) == 0  # the drop-in's entry point


def test_c3_whole_chain_equals_reference(ctx, ref_oracle):
    import bench_strong
    cfg = bench_strong.STRONG["c3"]
    img = synth.make_texture(cfg["base"], cfg["base"], cfg["kind"])
    plan, got = ctx.compress(img, capi.MODE_BC7, True, capi.default_params(**cfg["params"]))
    dims = [(int(plan.level_width[l]), int(plan.level_height[l])) for l in range(plan.num_levels)]
    rep = bench_strong.reference_sample("c3", cfg, img, got, dims, THREADS)
    assert rep["against"].startswith("unmodified reference")
    assert rep["blocks"] == sum(int(plan.level_num_blocks[l]) for l in range(plan.num_levels)) == 5592405
    assert rep["mismatched_blocks"] == 0
    assert set(synth.mode_histogram(got[0])) == {1, 5, 6, 7}


def test_c4_material_batch_equals_reference_compress(ctx, ref_oracle):
    imgs = [synth.make_texture(4096, 4096, 1 if i % 4 == 3 else 0, seed=0xB200 + i) for i in range(8)]
    got = ctx.compress_batch(imgs, capi.MODE_BC7, True)
    for i, im in enumerate(imgs):
        want = ref_oracle.compress(im, 1, True, THREADS)["levels"]
        assert _mismatches(got[i], want) == 0, f"texture {i}"


def test_c5_slab_and_deep_levels_equal_reference(ctx, ref_oracle):
    import bench_strong
    cfg = bench_strong.STRONG["c5"]
    size = cfg["base"]
    img = np.empty((size, size, 4), dtype=np.uint8)
    from concurrent.futures import ThreadPoolExecutor
    rows = [(a, min(a + 1024, size)) for a in range(0, size, 1024)]
    with ThreadPoolExecutor(min(THREADS, 16)) as ex:  # (numpy-bound; threads only overlap the memory traffic)
        list(ex.map(lambda r: synth.make_texture(size, size, cfg["kind"], rows=r, out=img[r[0]:r[1]]), rows))
    plan, got = ctx.compress(img, capi.MODE_BC7, True, capi.default_params(**cfg["params"]))
    dims = [(int(plan.level_width[l]), int(plan.level_height[l])) for l in range(plan.num_levels)]
    rep = bench_strong.reference_sample("c5", cfg, img, got, dims, THREADS)
    assert rep["against"].startswith("unmodified reference")
    assert rep["blocks"] >= 2_600_000 and rep["blocks_per_level"][0] >= 2_000_000
    assert all(rep["blocks_per_level"][l] == (dims[l][0] // 4) * (dims[l][1] // 4) for l in range(3, len(dims)))
    assert rep["mismatched_blocks"] == 0
