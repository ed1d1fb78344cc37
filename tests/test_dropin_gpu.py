"""The C++20 drop-in vierkant::bcn::compress() (integration/texture_block_compression_cuda.cpp, compiled against the
reference's own headers) run through the reference's tests/TestCompressionBC7.cpp cases: the assertions of its check()
(:40-57) plus what the reference never pinned -- the bytes, against the reference build itself (oracle/_ref)."""
import ctypes as C
import os

import numpy as np
import pytest

from vierkant_b200 import synth

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SO = os.path.join(ROOT, "integration", "_build", "libvkt_dropin_test.so")


def _load():
    if not os.path.exists(SO):
        pytest.skip("integration/_build/libvkt_dropin_test.so not built (needs the reference headers)")
    L = C.CDLL(SO)
    L.dropin_compress.argtypes = [C.c_void_p, C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint32, C.c_int]
    L.dropin_compress.restype = C.c_void_p
    for name, res in [("num_levels", C.c_uint32), ("base_width", C.c_uint32), ("base_height", C.c_uint32), ("mode", C.c_uint32),
                      ("duration_ms", C.c_int64)]:
        f = getattr(L, "dropin_result_" + name)
        f.argtypes, f.restype = [C.c_void_p], res
    L.dropin_result_level_blocks.argtypes, L.dropin_result_level_blocks.restype = [C.c_void_p, C.c_uint32], C.c_uint64
    L.dropin_result_level_data.argtypes, L.dropin_result_level_data.restype = [C.c_void_p, C.c_uint32], C.c_void_p
    L.dropin_result_free.argtypes = [C.c_void_p]
    return L


def test_dropin_library_exports():
    """CPU-only: the drop-in links against the C ABI and exports its entry points (no compute call)."""
    L = _load()
    assert all(hasattr(L, n) for n in ("dropin_compress", "dropin_result_level_data", "dropin_result_free"))


def _compress(L, img, mode, mips):
    img = np.ascontiguousarray(img, dtype=np.uint8)
    h, w, c = img.shape
    r = L.dropin_compress(img.ctypes.data, w, h, c, mode, int(mips))
    try:
        levels = []
        for l in range(L.dropin_result_num_levels(r)):
            n = L.dropin_result_level_blocks(r, l)
            buf = (C.c_uint8 * (16 * n)).from_address(L.dropin_result_level_data(r, l))
            levels.append(np.frombuffer(buf, dtype=np.uint8).reshape(n, 16).copy())
        return {"mode": L.dropin_result_mode(r), "base_width": L.dropin_result_base_width(r), "base_height": L.dropin_result_base_height(r),
                "duration_ms": L.dropin_result_duration_ms(r), "levels": levels}
    finally:
        L.dropin_result_free(r)


def _round4(v):
    return (v + 3) & ~3


def _num_levels(w, h):  # tests/TestCompressionBC7.cpp:13-19
    return max(0, int(np.log2(max(_round4(w), _round4(h))) - 2)) + 1


def _num_blocks(w, h, level):  # tests/TestCompressionBC7.cpp:21-38
    w, h = _round4(w), _round4(h)
    for _ in range(level):
        w, h = _round4(max(w // 2, 1)), _round4(max(h // 2, 1))
    return (w // 4) * (h // 4)


CASES = [("CompressionBC5.basic", 4, 512, 256, 0, False), ("CompressionBC7.basic", 4, 512, 256, 1, False),
         ("CompressionBC7.missing_alpha", 3, 64, 128, 1, False), ("CompressionBC7.mips", 4, 512, 256, 1, True),
         ("CompressionBC7.odd_size", 4, 123, 81, 1, True)]


@pytest.mark.gpu
@pytest.mark.parametrize("name,comps,w,h,mode,mips", CASES)
def test_reference_test_cases_through_the_dropin(port_oracle, name, comps, w, h, mode, mips):
    L = _load()
    img = port_oracle.resize(synth.checkerboard_4x4(comps), w, h)  # the tests resize the 4x4 checkerboard to w x h
    r = _compress(L, img, mode, mips)
    # check(), tests/TestCompressionBC7.cpp:40-57
    assert r["mode"] == mode
    assert r["duration_ms"] > 0
    assert (r["base_width"], r["base_height"]) == (_round4(w), _round4(h))
    assert len(r["levels"]) == (_num_levels(w, h) if mips else 1)
    for l, blocks in enumerate(r["levels"]):
        assert blocks.shape[0] == _num_blocks(w, h, l)
    # what the reference's tests never pinned: the bytes
    want = port_oracle.compress(img, mode, mips, threads=os.cpu_count() or 1)
    for got, ref in zip(r["levels"], want["levels"]):
        assert np.array_equal(got, ref)


@pytest.mark.gpu
def test_dropin_matches_reference_build(ref_oracle):
    L = _load()
    img = synth.make_texture(200, 120, 1, seed=77)
    got, want = _compress(L, img, 1, True), ref_oracle.compress(img, 1, True, 0)
    assert len(got["levels"]) == len(want["levels"])
    for a, b in zip(got["levels"], want["levels"]):
        assert np.array_equal(a, b)


@pytest.mark.gpu
def test_dropin_batch_overload_matches_single_calls():
    """vierkant::bcn::compress(std::span<const compress_info_t>) (integration/texture_block_compression_batch.hpp): the
    textures of a model in one call return what the per-texture calls return."""
    L = _load()
    L.dropin_compress_pair.argtypes = [C.c_void_p, C.c_uint32, C.c_uint32, C.c_void_p, C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint32, C.c_int, C.c_int]
    L.dropin_compress_pair.restype = C.c_void_p
    a = np.ascontiguousarray(synth.make_texture(200, 120, 1, seed=5))
    b = np.ascontiguousarray(synth.make_texture(64, 256, 0, seed=6))
    for which, img in ((0, a), (1, b)):
        r = L.dropin_compress_pair(a.ctypes.data, 200, 120, b.ctypes.data, 64, 256, 4, 1, 1, which)
        try:
            want = _compress(L, img, 1, True)
            assert L.dropin_result_num_levels(r) == len(want["levels"]) and L.dropin_result_duration_ms(r) > 0
            for l, ref in enumerate(want["levels"]):
                n = L.dropin_result_level_blocks(r, l)
                buf = (C.c_uint8 * (16 * n)).from_address(L.dropin_result_level_data(r, l))
                assert np.array_equal(np.frombuffer(buf, dtype=np.uint8).reshape(n, 16), ref)
        finally:
            L.dropin_result_free(r)


REF_GTEST = os.path.join(ROOT, "integration", "_build", "ref_gtest_bc7")
REF_GTEST_CASES = ["CompressionBC5.", "  basic", "CompressionBC7.", "  basic", "  missing_alpha", "  mips", "  odd_size"]


def test_reference_gtest_binary_lists_the_reference_cases():
    """CPU-only: the reference's tests/TestCompressionBC7.cpp, compiled unmodified against the drop-in (integration/Makefile),
    carries exactly the reference's five cases (no compute: --gtest_list_tests)."""
    import subprocess
    if not os.path.exists(REF_GTEST):
        pytest.skip("integration/_build/ref_gtest_bc7 not built (needs the reference tree)")
    out = subprocess.run([REF_GTEST, "--gtest_list_tests"], capture_output=True, text=True, timeout=60)
    assert out.returncode == 0, out.stderr
    assert [l.rstrip() for l in out.stdout.splitlines() if l.strip() and not l.startswith("Running main()")] == REF_GTEST_CASES


@pytest.mark.gpu
def test_reference_gtest_passes_against_the_dropin():
    """The reference's own contract test (tests/TestCompressionBC7.cpp:59-131), unmodified source, googletest as vendored by the
    reference, vierkant::bcn::compress() = the CUDA drop-in: all five cases must pass on the GPU."""
    import subprocess
    if not os.path.exists(REF_GTEST):
        pytest.skip("integration/_build/ref_gtest_bc7 not built (needs the reference tree)")
    out = subprocess.run([REF_GTEST], capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stdout[-2000:] + out.stderr[-2000:]
    assert "[  PASSED  ] 5 tests." in out.stdout
