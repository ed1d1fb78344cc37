"""The device-side block search (vierkant_b200/csrc/bc7_core.cuh) compiled as host C++ must match the oracle: this
keeps the search logic checked on machines without a GPU.  The emulation library is test infrastructure only."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

from cases import PARAM_CASES, edge_tiles
from oracle.pyoracle import Bc7Params, default_params
from vierkant_b200 import synth

HERE = os.path.dirname(os.path.abspath(__file__))


@pytest.fixture(scope="module")
def emul():
    subprocess.run(["make", "-s", "-C", os.path.join(HERE, "host_emul")], check=True)
    lib = C.CDLL(os.path.join(HERE, "host_emul", "libvkt_emul.so"))
    lib.emul_bc7_encode_blocks.argtypes = [C.POINTER(C.c_uint8), C.c_uint64, C.POINTER(Bc7Params), C.POINTER(C.c_uint8), C.c_int]

    def encode(tiles, params):
        tiles = np.ascontiguousarray(tiles, dtype=np.uint8)
        n = tiles.size // 64
        out = np.zeros((n, 16), dtype=np.uint8)
        rc = lib.emul_bc7_encode_blocks(tiles.ctypes.data_as(C.POINTER(C.c_uint8)), n, C.byref(params),
                                        out.ctypes.data_as(C.POINTER(C.c_uint8)), 4)
        return rc, out
    return encode


@pytest.mark.parametrize("case", sorted(PARAM_CASES))
def test_device_logic_matches_oracle(emul, port_oracle, case):
    tiles = np.concatenate([edge_tiles(21, 30), synth.to_blocks(synth.make_texture(96, 96, 1, seed=5))])
    p = default_params(**PARAM_CASES[case])
    rc, got = emul(tiles, p)
    assert rc == 0
    assert np.array_equal(got, port_oracle.encode_blocks(tiles, p, threads=4))


@pytest.mark.parametrize("kw", [dict(force_selectors=1), dict(quant_mode6_endpoints=1), dict(low_frequency_partition_weight=0.5)])
def test_unsupported_knobs_are_rejected(emul, kw):
    rc, _ = emul(edge_tiles(1, 1), default_params(**kw))
    assert rc == -2
