"""The C-ABI library loads and exports every symbol include/vierkant_bcn_cuda.h declares (no compute without a GPU)."""
import os
import re

import pytest

from vierkant_b200 import capi

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_functions():
    text = open(os.path.join(ROOT, "include", "vierkant_bcn_cuda.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(vkt_bc[n7]_\w+)\s*\(", text)))


def test_header_functions_are_exported(cuda_lib):
    names = declared_functions()
    assert len(names) >= 15
    missing = [n for n in names if not hasattr(cuda_lib, n)]
    assert not missing, missing
    assert sorted(capi.EXPORTS) == names


def test_params_init_matches_bc7enc_defaults(cuda_lib):
    p = capi.default_params()
    assert p.mode_mask == 0xFFFFFFFF and p.max_partitions == 64 and list(p.weights) == [128, 64, 16, 32]
    assert p.uber_level == 0 and p.perceptual == 1 and p.try_least_squares == 1
    assert p.mode17_partition_estimation_filterbank == 1 and p.force_alpha == 0 and p.force_selectors == 0
    assert p.quant_mode6_endpoints == 0 and p.bias_mode1_pbits == 0
    assert [p.pbit1_weight, p.mode1_error_weight, p.mode5_error_weight, p.mode6_error_weight, p.mode7_error_weight,
            p.low_frequency_partition_weight] == [1.0] * 6


def test_params_layout_matches_oracle_mirror():
    import ctypes
    from oracle import pyoracle
    assert ctypes.sizeof(capi.Bc7Params) == ctypes.sizeof(pyoracle.Bc7Params)
    assert [f[0] for f in capi.Bc7Params._fields_] == [f[0] for f in pyoracle.Bc7Params._fields_]


@pytest.mark.parametrize("w,h,mips,levels", [(512, 256, False, 1), (512, 256, True, 8), (123, 81, True, 5), (64, 128, False, 1),
                                             (4, 4, True, 1), (1, 1, True, 1), (1024, 1024, True, 9), (16384, 16384, True, 13)])
def test_compress_plan_follows_reference_level_rules(cuda_lib, w, h, mips, levels):
    """texture_block_compression.cpp:80-86,141-146 and tests/TestCompressionBC7.cpp:13-38."""
    plan = capi.compress_plan(w, h, mips)
    r4 = lambda v: (v + 3) & ~3
    assert (plan.base_width, plan.base_height) == (r4(w), r4(h))
    assert plan.num_levels == levels
    bw, bh = r4(w), r4(h)
    for l in range(plan.num_levels):
        assert (plan.level_width[l], plan.level_height[l]) == (bw, bh)
        assert plan.level_num_blocks[l] == (bw // 4) * (bh // 4)
        bw, bh = r4(max(bw // 2, 1)), r4(max(bh // 2, 1))


def test_no_cpu_fallback_without_device(cuda_lib):
    if capi.device_count() > 0:
        pytest.skip("a GPU is present")
    with pytest.raises(capi.BcnError) as e:
        capi.BcnContext()
    assert e.value.code == capi.ERR_NO_DEVICE
