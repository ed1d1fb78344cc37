"""One context over several B200s (run with gpurun --gpus N): block rows / levels are split over the devices and gathered
with async copies; the result must be byte-identical to the single-device one.  Skipped on single-GPU boxes."""
import numpy as np
import pytest

from vierkant_b200 import capi, synth

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def multi_ctx(cuda_lib):
    n = capi.device_count()
    if n < 2:
        pytest.skip("needs at least 2 CUDA devices")
    c = capi.BcnContext(list(range(n)))
    yield c
    c.close()


@pytest.fixture(scope="module")
def known_2048():
    import json
    import os
    with open(os.path.join(os.path.dirname(__file__), "golden", "known_answers.json")) as f:
        for e in json.load(f)["compress"]:
            if e["size"] == 2048 and e["kind"] == 0:
                return e["fnv1a64"]
    pytest.skip("no 2048^2 known answer")


def test_encode_rows_split_over_devices(ctx, multi_ctx):
    img = synth.make_texture(1024, 512, 1, seed=3)
    assert multi_ctx.num_devices >= 2
    assert np.array_equal(multi_ctx.encode_bc7(img), ctx.encode_bc7(img))
    assert np.array_equal(multi_ctx.encode_bc5(img), ctx.encode_bc5(img))


def test_compress_chain_split_over_devices(ctx, multi_ctx):
    img = synth.make_texture(1000, 520, 1, seed=4)
    _, a = multi_ctx.compress(img, capi.MODE_BC7, True)
    _, b = ctx.compress(img, capi.MODE_BC7, True)
    assert len(a) == len(b)
    for x, y in zip(a, b):
        assert np.array_equal(x, y)


def test_compress_chain_big_split_over_devices(ctx, multi_ctx, known_2048):
    """2048^2 + mips (graded band pipeline on every device, each encoding its own block rows): identical to one device
    and to the reference's hash."""
    img = synth.make_texture(2048, 2048, 0)
    _, a = multi_ctx.compress(img, capi.MODE_BC7, True)
    _, b = ctx.compress(img, capi.MODE_BC7, True)
    for x, y in zip(a, b):
        assert np.array_equal(x, y)
    assert "%016x" % synth.fnv1a64_words(np.concatenate(a)) == known_2048


def test_compress_batch_over_devices(ctx, multi_ctx):
    """Whole chains of several textures, round-robin over the devices (two lanes each): identical to one device."""
    imgs = [synth.make_texture(512 >> (i % 3), 256, i & 1, seed=30 + i) for i in range(11)]
    a = multi_ctx.compress_batch(imgs, capi.MODE_BC7, True)
    b = ctx.compress_batch(imgs, capi.MODE_BC7, True)
    for la, lb in zip(a, b):
        assert len(la) == len(lb)
        for x, y in zip(la, lb):
            assert np.array_equal(x, y)


def test_batch_of_textures_over_devices(ctx, multi_ctx):
    imgs = [synth.make_texture(256 >> (i % 3), 128, i & 1, seed=10 + i) for i in range(9)]
    outs_a = [np.empty(((i.shape[0] // 4) * (i.shape[1] // 4), 16), dtype=np.uint8) for i in imgs]
    outs_b = [np.empty_like(o) for o in outs_a]
    multi_ctx.encode_batch(capi.MODE_BC7, imgs, outs_a)
    ctx.encode_batch(capi.MODE_BC7, imgs, outs_b)
    for x, y in zip(outs_a, outs_b):
        assert np.array_equal(x, y)
