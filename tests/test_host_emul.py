"""The device-side block search (vierkant_b200/csrc/bc7_core.cuh) compiled as host C++ must match the oracle: this
keeps the search logic checked on machines without a GPU.  The emulation library is test infrastructure only."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

from cases import INVALID_CASES, PARAM_CASES, RDO_CASES, edge_tiles, random_params
from oracle.pyoracle import Bc7Params, default_params
from vierkant_b200 import synth

HERE = os.path.dirname(os.path.abspath(__file__))


@pytest.fixture(scope="module")
def emul():
    subprocess.run(["make", "-s", "-C", os.path.join(HERE, "host_emul")], check=True)
    lib = C.CDLL(os.path.join(HERE, "host_emul", "libvkt_emul.so"))
    lib.emul_bc7_encode_blocks.argtypes = [C.POINTER(C.c_uint8), C.c_uint64, C.POINTER(Bc7Params), C.POINTER(C.c_uint8), C.c_int]
    lib.emul_resize_u8.argtypes = [C.c_void_p, C.c_uint32, C.c_uint32, C.c_uint32, C.c_void_p, C.c_uint32, C.c_uint32]

    def encode(tiles, params):
        tiles = np.ascontiguousarray(tiles, dtype=np.uint8)
        n = tiles.size // 64
        out = np.zeros((n, 16), dtype=np.uint8)
        rc = lib.emul_bc7_encode_blocks(tiles.ctypes.data_as(C.POINTER(C.c_uint8)), n, C.byref(params),
                                        out.ctypes.data_as(C.POINTER(C.c_uint8)), 4)
        return rc, out

    def resize(img, ow, oh):
        img = np.ascontiguousarray(img, dtype=np.uint8)
        h, w, c = img.shape
        out = np.zeros((oh, ow, c), dtype=np.uint8)
        assert lib.emul_resize_u8(img.ctypes.data, w, h, c, out.ctypes.data, ow, oh) == 0
        return out
    encode.resize = resize
    encode.lib = lib
    return encode


def test_uber_selector_maps_match_reference_expression(emul):
    """bc7_core.cuh generates the uber-level selector rescaling table with integer arithmetic at compile time; it must be
    the reference's float expression (bc7enc.cpp:1399) for every (max selector, ly, hy, selector)."""
    assert emul.lib.emul_uber_map_mismatches() == 0


ALL_CASES = {**PARAM_CASES, **RDO_CASES}


@pytest.mark.parametrize("case", sorted(ALL_CASES))
def test_device_logic_matches_oracle(emul, port_oracle, case):
    tiles = np.concatenate([edge_tiles(21, 30), synth.to_blocks(synth.make_texture(96, 96, 1, seed=5))])
    p = default_params(**ALL_CASES[case])
    rc, got = emul(tiles, p)
    assert rc == 0
    assert np.array_equal(got, port_oracle.encode_blocks(tiles, p, threads=4))


@pytest.mark.parametrize("case", sorted(RDO_CASES))
def test_rdo_knobs_match_reference_build(emul, ref_oracle, case):
    """The extended variant (forced selectors, reduced mode-6 quantisation, low-frequency partition weight) against the
    unmodified reference itself."""
    tiles = np.concatenate([edge_tiles(33, 24), synth.to_blocks(synth.make_texture(64, 64, 0, seed=8))])
    p = default_params(**RDO_CASES[case])
    rc, got = emul(tiles, p)
    assert rc == 0
    assert np.array_equal(got, ref_oracle.encode_blocks(tiles, p, threads=4))


@pytest.mark.parametrize("case", sorted(INVALID_CASES))
def test_invalid_parameters_are_rejected(emul, case):
    rc, _ = emul(edge_tiles(1, 1), default_params(**INVALID_CASES[case]))
    assert rc == -1


def test_resize_decode_is_the_exact_quotient(emul):
    """The resize passes decode u8 / 255.0f with one Newton step instead of a division: exact for all 256 values."""
    assert emul.lib.emul_resize_decode_mismatches() == 0


def test_resize_strip_kernel_arithmetic_is_exact(emul):
    """The strip kernels' conversion-free sample arithmetic: float(v) from the 2^23 + v bit pattern and the two-instruction
    quotient for all 256 values; the two round-toward-zero additions of the encode against (int)((double) s + 0.5) for
    EVERY float in [0, 255] (1.13e9 values)."""
    import ctypes as C
    assert emul.lib.emul_resize_decode_split_mismatches() == 0
    emul.lib.emul_resize_encode_mismatches.restype = C.c_longlong
    emul.lib.emul_resize_encode_mismatches.argtypes = [C.c_uint32]
    assert emul.lib.emul_resize_encode_mismatches(1) == 0


# (w, h, ow, oh, strip, y0, y1) -- y1 == 0: all rows.  One thread per row (first == last), two, a partial warp, more than one CTA
# of column groups; heights that are no multiple of the strip or of the unrolled period (3 / 4 rows); row ranges as the band loop
# of compress() issues them; the strips the library picks (16, 8, 4).
STRIP_CASES = [(4, 4, 4, 4, 16, 0, 0), (8, 8, 4, 4, 8, 0, 0), (4, 64, 4, 64, 4, 0, 0), (8, 128, 4, 64, 16, 0, 0), (8, 6, 8, 6, 4, 0, 0),
               (16, 10, 8, 5, 8, 0, 0), (516, 20, 516, 20, 16, 0, 0), (1032, 40, 516, 20, 8, 0, 0), (260, 12, 260, 12, 4, 0, 0),
               (12, 300, 12, 300, 16, 0, 0), (24, 600, 12, 300, 16, 0, 0), (64, 37, 64, 37, 16, 0, 0), (128, 74, 64, 37, 16, 0, 0),
               (512, 3, 512, 3, 16, 0, 0), (1028, 2, 514, 1, 8, 0, 0), (8, 4, 8, 4, 1, 0, 0), (8, 8, 4, 4, 1, 0, 0),
               (256, 96, 256, 96, 16, 17, 63), (256, 96, 256, 96, 4, 0, 5), (512, 192, 256, 96, 16, 31, 96), (512, 192, 256, 96, 8, 5, 6),
               (2048, 24, 2048, 24, 16, 0, 0), (2056, 48, 1028, 24, 16, 0, 0)]


@pytest.mark.parametrize("w,h,ow,oh,strip,y0,y1", STRIP_CASES)
def test_resize_strip_thread_matches_oracle(emul, port_oracle, w, h, ow, oh, strip, y0, y1):
    """The per-thread body of the CUDA strip kernels (resize_strip.h), run thread by thread on the host -- surplus threads of the
    last CTA included -- against the oracle: the same bytes inside the requested rows, nothing written outside them."""
    lib = emul.lib
    lib.emul_resize_strip.argtypes = [C.c_void_p, C.c_uint32, C.c_uint32, C.c_void_p, C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint32, C.c_int]
    y1 = y1 or oh
    rng = np.random.default_rng(13 * w + h + strip)
    for img in (rng.integers(0, 256, (h, w, 4), dtype=np.uint8), np.where(rng.random((h, w, 4)) < 0.5, 0, 255).astype(np.uint8)):
        want = port_oracle.resize(img, ow, oh)
        out = np.full((oh, ow, 4), 0xAB, dtype=np.uint8)
        assert lib.emul_resize_strip(img.ctypes.data, w, h, out.ctypes.data, ow, oh, y0, y1, strip) == 0
        assert np.array_equal(out[y0:y1], want[y0:y1]), int((out[y0:y1] != want[y0:y1]).sum())
        assert (out[:y0] == 0xAB).all() and (out[y1:] == 0xAB).all()


def test_resize_tap_lists_ascend(emul):
    """The vertical CUDA pass visits input rows in ascending order and feeds each to the outputs whose next tap it is: every
    tap list must ascend (repeats allowed: clamped margins), for reductions, enlargements and the 1:1 Mitchell pass."""
    pairs = [(n, n) for n in (4, 5, 8, 123, 124, 4096)] + [(n, (n // 2 + 3) & ~3) for n in (8, 12, 20, 84, 124, 520, 4096, 16384)]
    pairs += [(123, 124), (81, 84), (4, 512), (4, 256), (1000, 3), (1000, 4), (17, 20), (300, 150), (7, 8), (50, 200), (50, 30), (33, 36)]
    for i, o in pairs:
        assert emul.lib.emul_resize_axis_order_violations(i, o) == 0, (i, o)


RESIZE_SHAPES = [(64, 64, 64, 64, 4), (64, 64, 32, 32, 4), (123, 81, 124, 84, 4), (256, 128, 16, 8, 4), (60, 36, 60, 36, 3),
                 (4, 4, 512, 256, 4), (100, 52, 52, 28, 4), (124, 84, 64, 44, 4), (8, 4, 4, 4, 4), (4, 4, 4, 4, 4), (12, 20, 8, 12, 3),
                 (33, 7, 36, 8, 1), (50, 50, 200, 30, 2), (17, 300, 20, 150, 4), (640, 8, 320, 4, 4), (1024, 16, 512, 8, 4),
                 (2048, 4, 2048, 4, 4), (1000, 4, 3, 4, 4)]


@pytest.mark.parametrize("w,h,ow,oh,c", RESIZE_SHAPES)
def test_resize_tap_lists_match_oracle(emul, port_oracle, w, h, ow, oh, c):
    """The per-output tap lists the CUDA resize passes consume (resize_axis.h), evaluated on the CPU in kernel order."""
    img = synth.make_texture(w, h, 1, seed=3 * w + h)[..., :c]
    assert np.array_equal(emul.resize(img, ow, oh), port_oracle.resize(img, ow, oh))


def test_bc4_reciprocal_count_equals_the_threshold_compares(emul):
    """bc5_core.cuh counts the thresholds a texel reaches with one multiplication by a rounded-up reciprocal; the reference compares
    seven times (rgbcx.cpp:2655-2683).  Exhaustive over every delta (1..255) and every numerator the encoder can produce."""
    assert emul.lib.emul_bc4_count_mismatches() == 0


def test_bc5_host_build_matches_the_oracle(emul, port_oracle):
    rng = np.random.default_rng(17)
    tiles = np.concatenate([rng.integers(0, 256, size=(4000, 16, 4), dtype=np.uint8),
                            np.repeat(rng.integers(0, 256, size=(500, 1, 4), dtype=np.uint8), 16, axis=1),        # solid blocks
                            (rng.integers(0, 2, size=(1500, 16, 4)) * rng.integers(1, 256, size=(1500, 1, 4))).astype(np.uint8),  # two-valued
                            synth.to_blocks(synth.make_texture(128, 128, 1, seed=3))])
    tiles = np.ascontiguousarray(tiles)
    out = np.zeros((tiles.shape[0], 16), dtype=np.uint8)
    emul.lib.emul_bc5_encode_blocks.argtypes = [C.c_void_p, C.c_uint64, C.c_void_p]
    emul.lib.emul_bc5_encode_blocks.restype = None
    emul.lib.emul_bc5_encode_blocks(tiles.ctypes.data, tiles.shape[0], out.ctypes.data)
    assert np.array_equal(out, port_oracle.encode_bc5_blocks(tiles))


def test_random_parameter_sets_host_build_matches_the_port(emul, port_oracle):
    """The parameter fuzz of tests/test_bc7_gpu.py on the host build of the device code (no GPU needed): 32 random valid
    parameter sets, every knob in combination, against the oracle."""
    rng = np.random.default_rng(2026)
    for case in range(32):
        kw = random_params(rng)
        tiles = edge_tiles(1000 + case, 10)
        p = default_params(**kw)
        rc, got = emul(tiles, p)
        assert rc == 0, kw
        want = port_oracle.encode_blocks(tiles, p, threads=os.cpu_count() or 1)
        bad = int((got != want).any(axis=1).sum())
        assert bad == 0, f"case {case}: {bad} of {len(tiles)} blocks differ with {kw}"


def test_filterbank_bin_order_is_a_permutation_and_shrinks_the_union():
    """bc7_core.cuh sorts a CTA's blocks by kBinOfKey[key iteration] before the filterbank phase: the constant must be a
    permutation of 0..13 (a repeated bin would merge two keys' counts, a missing one leave a hole -- harmless for the result,
    which never depends on which lane scores a block, but not what is meant), and on uniformly distributed keys the simulated
    union a warp pays for must be well below that of the plain iteration order (tools/bin_order.py; 16.98 -> 15.31 candidates,
    slowest warp of a CTA 20.95 -> 18.00)."""
    import importlib.util
    import os
    spec = importlib.util.spec_from_file_location("bin_order", os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tools", "bin_order.py"))
    bo = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(bo)
    order = bo.shipped_order()
    assert sorted(order) == list(range(14))
    cand = bo.candidate_sets()
    keys = bo.uniform_keys(n=300, seed=11)
    plain, shipped = bo.cost(list(range(14)), keys, cand), bo.cost(order, keys, cand)
    assert shipped[0] < plain[0] - 1.2 and shipped[1] < plain[1] - 2.0, (plain, shipped)
