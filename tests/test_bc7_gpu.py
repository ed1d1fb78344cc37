"""Parity of the CUDA path (through the C ABI) against the oracle.  Bit-exact: the work is integer/byte output."""
import json
import os

import numpy as np
import pytest

from cases import INVALID_CASES, PARAM_CASES, RDO_CASES, edge_tiles, random_params, tiles_to_image
from oracle.pyoracle import default_params as oracle_params
from vierkant_b200 import capi, synth

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(__file__), "golden")


def gpu_params(kw):
    return capi.default_params(**kw)


@pytest.fixture(scope="module")
def golden():
    return np.load(os.path.join(GOLD, "bc7_blocks.npz"))


ALL_CASES = {**PARAM_CASES, **RDO_CASES}


@pytest.mark.parametrize("case", sorted(ALL_CASES))
def test_golden_vectors(ctx, golden, case):
    """Committed reference outputs (tests/golden/make_golden.py) for every parameter case."""
    tiles = golden["tiles"]
    n = tiles.shape[0]
    bx = 8
    pad = (-n) % bx
    t = np.concatenate([tiles, np.repeat(tiles[-1:], pad, axis=0)]) if pad else tiles
    got = ctx.encode_bc7(tiles_to_image(t, bx), gpu_params(ALL_CASES[case]))[:n]
    assert np.array_equal(got, golden["blocks_" + case])


@pytest.mark.parametrize("case", sorted(ALL_CASES))
@pytest.mark.parametrize("kind", [0, 1])
def test_synthetic_texture_matches_oracle(ctx, port_oracle, case, kind):
    img = synth.make_texture(256, 256, kind)
    want = port_oracle.encode_blocks(synth.to_blocks(img), oracle_params(**ALL_CASES[case]), threads=os.cpu_count() or 1)
    got = ctx.encode_bc7(img, gpu_params(ALL_CASES[case]))
    mism = int((got != want).any(axis=1).sum())
    assert mism == 0, f"{mism} / {len(want)} blocks differ"


def test_edge_tiles_match_reference_build(ctx, ref_oracle):
    """Against the unmodified reference itself (oracle/_ref), not just the restatement."""
    tiles = edge_tiles(77, 128)
    img = tiles_to_image(tiles, 32)
    for kw in (dict(), dict(uber_level=4, mode17_partition_estimation_filterbank=0), dict(perceptual=0, weights=[1, 1, 1, 1])):
        want = ref_oracle.encode_blocks(tiles, oracle_params(**kw), threads=os.cpu_count() or 1)
        assert np.array_equal(ctx.encode_bc7(img, gpu_params(kw)), want)


@pytest.mark.parametrize("w,h", [(4, 4), (8, 4), (4, 64), (260, 12), (124, 84), (36, 4)])
def test_ragged_and_tiny_shapes(ctx, port_oracle, w, h):
    img = synth.make_texture(w, h, 1, seed=w * 131 + h)
    want = port_oracle.encode_blocks(synth.to_blocks(img))
    assert np.array_equal(ctx.encode_bc7(img), want)


def test_three_component_image_gets_opaque_alpha(ctx, port_oracle):
    """get_block injects alpha = 255 for 3-component images (texture_block_compression.cpp:39-60)."""
    rgba = synth.make_texture(64, 128, 1)
    rgb = np.ascontiguousarray(rgba[..., :3])
    opaque = rgba.copy()
    opaque[..., 3] = 255
    want = port_oracle.encode_blocks(synth.to_blocks(opaque))
    assert np.array_equal(ctx.encode_bc7(rgb), want)


def test_row_stride(ctx, port_oracle):
    img = synth.make_texture(64, 32, 1)
    padded = np.zeros((32, 64 * 4 + 48), dtype=np.uint8)
    padded[:, : 64 * 4] = img.reshape(32, -1)
    out = np.empty((8 * 16, 16), dtype=np.uint8)
    rc = ctx.lib.vkt_bcn_cuda_encode_bc7(ctx.handle, padded.ctypes.data, 64, 32, 4, padded.shape[1], None, out.ctypes.data)
    assert rc == 0
    assert np.array_equal(out, port_oracle.encode_blocks(synth.to_blocks(img)))


def test_batch_of_levels(ctx, port_oracle):
    imgs = [synth.make_texture(s, s // 2, k, seed=s) for s, k in [(128, 0), (64, 1), (32, 1), (16, 0), (8, 1)]]
    outs = [np.empty(((i.shape[0] // 4) * (i.shape[1] // 4), 16), dtype=np.uint8) for i in imgs]
    ctx.encode_batch(capi.MODE_BC7, imgs, outs)
    for i, o in zip(imgs, outs):
        assert np.array_equal(o, port_oracle.encode_blocks(synth.to_blocks(i)))


@pytest.mark.parametrize("name", sorted(INVALID_CASES))
def test_invalid_parameters_fail_loudly(ctx, name):
    with pytest.raises(capi.BcnError) as e:
        ctx.encode_bc7(synth.make_texture(8, 8, 0), gpu_params(INVALID_CASES[name]))
    assert e.value.code == capi.ERR_INVALID


def test_rdo_knobs_match_reference_build(ctx, ref_oracle):
    """Forced selectors, reduced mode-6 quantisation and the low-frequency partition weight against the unmodified reference
    itself (oracle/_ref), on edge-case tiles."""
    tiles = edge_tiles(91, 96)
    img = tiles_to_image(tiles, 32)
    for name in sorted(RDO_CASES):
        want = ref_oracle.encode_blocks(tiles, oracle_params(**RDO_CASES[name]), threads=os.cpu_count() or 1)
        assert np.array_equal(ctx.encode_bc7(img, gpu_params(RDO_CASES[name])), want), name


@pytest.mark.parametrize("w,h,c", [(6, 4, 4), (4, 4, 2), (0, 4, 4)])
def test_invalid_arguments(ctx, w, h, c):
    buf = np.zeros(max(w * h * c, 16), dtype=np.uint8)
    out = np.zeros(64, dtype=np.uint8)
    rc = ctx.lib.vkt_bcn_cuda_encode_bc7(ctx.handle, buf.ctypes.data, w, h, c, 0, None, out.ctypes.data)
    assert rc == capi.ERR_INVALID


def test_known_answers_1024(ctx):
    """SURVEY.md App. C inputs at the BASELINE configs[0] size: reference-derived hash and mode histogram."""
    with open(os.path.join(GOLD, "known_answers.json")) as f:
        known = json.load(f)
    for entry in known["direct"]:
        img = synth.make_texture(entry["size"], entry["size"], entry["kind"])
        b = ctx.encode_bc7(img, gpu_params(entry["params"]))
        assert synth.mode_histogram(b) == {int(k): v for k, v in entry["modes"].items()}
        assert "%016x" % synth.fnv1a64_words(b) == entry["fnv1a64"], entry


def test_device_resident_entry_point(ctx, port_oracle):
    torch = pytest.importorskip("torch")
    img = synth.make_texture(128, 64, 1)
    d_in = torch.from_numpy(img).cuda()
    d_out = torch.empty(((64 // 4) * (128 // 4), 16), dtype=torch.uint8, device="cuda")
    ctx.encode_bc7_device(d_in, 128, 64, 4, d_out, stream=torch.cuda.current_stream().cuda_stream)
    torch.cuda.synchronize()
    assert np.array_equal(d_out.cpu().numpy(), port_oracle.encode_blocks(synth.to_blocks(img)))


def test_full_size_properties_4096(ctx, port_oracle):
    """BASELINE configs[1] size.  The oracle cannot encode 1 M blocks in seconds, so check size-independent properties:
    (i) encoding is a pure per-block function: any aligned crop encodes to the same blocks as the full image;
    (ii) a random sample of blocks matches the oracle; (iii) decoded PSNR is sane; (iv) determinism across runs."""
    img = synth.make_texture(4096, 4096, 1)
    full = ctx.encode_bc7(img)
    assert np.array_equal(full, ctx.encode_bc7(img))
    bx = 4096 // 4
    crop = np.ascontiguousarray(img[1024:1024 + 256, 2048:2048 + 512])
    got = ctx.encode_bc7(crop)
    rows = (np.arange(256 // 4) + 1024 // 4)[:, None] * bx + (np.arange(512 // 4) + 2048 // 4)[None, :]
    assert np.array_equal(got, full[rows.ravel()])
    rng = np.random.default_rng(3)
    idx = rng.choice(full.shape[0], 20000, replace=False)
    tiles = synth.to_blocks(img)[idx]
    assert np.array_equal(full[idx], port_oracle.encode_blocks(tiles, threads=os.cpu_count() or 1))
    dec = port_oracle.unpack_blocks(full[idx]).astype(np.float64)
    mse = ((dec - tiles.astype(np.float64)) ** 2).mean()
    assert 10 * np.log10(255 ** 2 / mse) > 30.0


@pytest.mark.parametrize("name,size,kind,kw,sample", [
    ("c3", 8192, 1, dict(max_partitions=64, mode17_partition_estimation_filterbank=0), 12000),
    ("c5", 2048, 0, dict(uber_level=4, max_partitions=64, mode17_partition_estimation_filterbank=0), 4000),
])
def test_full_size_properties_other_configs(ctx, port_oracle, name, size, kind, kw, sample):
    """BASELINE configs[2] (8192^2, alpha, every partition a candidate) at full size and the parameter set of configs[4]
    (uber 4) on a 2048^2 level: crops encode like the whole, a random sample of blocks matches the oracle, repeatable."""
    img = synth.make_texture(size, size, kind)
    full = ctx.encode_bc7(img, gpu_params(kw))
    bx = size // 4
    y0, x0 = size // 2 - 128, size // 2 - 256  # straddles the opaque / alpha halves of kind 1
    crop = np.ascontiguousarray(img[y0:y0 + 256, x0:x0 + 512])
    rows = (np.arange(256 // 4) + y0 // 4)[:, None] * bx + (np.arange(512 // 4) + x0 // 4)[None, :]
    assert np.array_equal(ctx.encode_bc7(crop, gpu_params(kw)), full[rows.ravel()])
    rng = np.random.default_rng(11)
    idx = rng.choice(full.shape[0], sample, replace=False)
    # tiles of the sampled blocks without materialising every tile of the image
    by, bxs = idx // bx, idx % bx
    tiles = np.stack([img[4 * y:4 * y + 4, 4 * x:4 * x + 4].reshape(16, 4) for y, x in zip(by, bxs)])
    assert np.array_equal(full[idx], port_oracle.encode_blocks(tiles, oracle_params(**kw), threads=os.cpu_count() or 1))
    if kind == 1:
        modes = synth.mode_histogram(full)
        assert set(modes) == {1, 5, 6, 7}  # the alpha half exercises every alpha mode


def test_concurrent_calls_on_one_context(ctx, port_oracle):
    """vierkant's loader may call compress() from several host threads (SURVEY.md 8b "Threading"): one context, four
    threads, mixed entry points; every result must be the single-threaded one."""
    import threading
    imgs = [synth.make_texture(256, 128, i & 1, seed=40 + i) for i in range(4)]
    want_blocks = [ctx.encode_bc7(im) for im in imgs]
    want_chain = [ctx.compress(im, capi.MODE_BC7, True)[1] for im in imgs]
    errors = []

    def work(i):
        try:
            for _ in range(6):
                if not np.array_equal(ctx.encode_bc7(imgs[i]), want_blocks[i]):
                    errors.append(("blocks", i))
                _, lv = ctx.compress(imgs[i], capi.MODE_BC7, True)
                if not all(np.array_equal(a, b) for a, b in zip(lv, want_chain[i])):
                    errors.append(("chain", i))
        except Exception as e:  # noqa: BLE001
            errors.append((repr(e), i))

    threads = [threading.Thread(target=work, args=(i,)) for i in range(4)]
    for t in threads:
        t.start()
    for t in threads:
        t.join()
    assert not errors
    assert np.array_equal(want_blocks[1], port_oracle.encode_blocks(synth.to_blocks(imgs[1])))


def test_random_parameter_sets_match_the_reference(ctx, ref_oracle):
    """Fuzz over the whole parameter space: 96 random (valid) parameter sets, each on edge-case tiles plus a slice of the alpha
    texture, GPU against bc7enc_compress_block of the unmodified reference -- every knob in combination, not one at a time."""
    rng = np.random.default_rng(2026)
    tex = synth.to_blocks(synth.make_texture(128, 128, 1, seed=7))
    for case in range(96):
        kw = random_params(rng)
        tiles = np.ascontiguousarray(np.concatenate([edge_tiles(1000 + case, 24), tex[(case * 64) % 1000:(case * 64) % 1000 + 160]]))
        n = tiles.shape[0]
        pad = (-n) % 16
        t = np.concatenate([tiles, np.repeat(tiles[-1:], pad, axis=0)]) if pad else tiles
        got = ctx.encode_bc7(tiles_to_image(t, 16), gpu_params(kw))[:n]
        want = ref_oracle.encode_blocks(tiles, oracle_params(**kw), threads=os.cpu_count() or 1)
        bad = int((got != want).any(axis=1).sum())
        assert bad == 0, f"case {case}: {bad} of {n} blocks differ with {kw}"
