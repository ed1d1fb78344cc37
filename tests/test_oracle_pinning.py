"""Pin the oracles (CPU, no GPU): the C restatement must reproduce the unmodified reference bit for bit, and both must
reproduce the committed golden vectors (generated from the reference by tests/golden/make_golden.py)."""
import json
import os

import numpy as np
import pytest

from cases import ORACLE_ONLY_CASES, PARAM_CASES, edge_tiles
from oracle.pyoracle import default_params
from vierkant_b200 import synth

GOLD = os.path.join(os.path.dirname(__file__), "golden")
ALL_CASES = {**PARAM_CASES, **ORACLE_ONLY_CASES}


@pytest.fixture(scope="module")
def golden():
    return np.load(os.path.join(GOLD, "bc7_blocks.npz"))


@pytest.fixture(scope="module")
def known():
    with open(os.path.join(GOLD, "known_answers.json")) as f:
        return json.load(f)


@pytest.mark.parametrize("case", sorted(ALL_CASES))
def test_port_matches_golden(port_oracle, golden, case):
    got = port_oracle.encode_blocks(golden["tiles"], default_params(**ALL_CASES[case]))
    assert np.array_equal(got, golden["blocks_" + case])


@pytest.mark.parametrize("case", sorted(ALL_CASES))
def test_ref_matches_golden(ref_oracle, golden, case):
    got = ref_oracle.encode_blocks(golden["tiles"], default_params(**ALL_CASES[case]))
    assert np.array_equal(got, golden["blocks_" + case])


@pytest.mark.parametrize("case", sorted(ALL_CASES))
def test_port_matches_ref_on_fresh_inputs(port_oracle, ref_oracle, case):
    tiles = np.concatenate([edge_tiles(1234, 40), synth.to_blocks(synth.make_texture(128, 128, 1, seed=99))])
    p = default_params(**ALL_CASES[case])
    assert np.array_equal(port_oracle.encode_blocks(tiles, p, threads=4), ref_oracle.encode_blocks(tiles, p, threads=4))


def test_port_decoder_matches_golden(port_oracle, golden):
    assert np.array_equal(port_oracle.unpack_blocks(golden["blocks_defaults"]), golden["decoded_defaults"])


def test_port_threads_do_not_change_output(port_oracle):
    tiles = edge_tiles(5, 20)
    assert np.array_equal(port_oracle.encode_blocks(tiles, None, threads=1), port_oracle.encode_blocks(tiles, None, threads=7))


def test_known_answer_direct_1024_defaults(port_oracle, known):
    """mode histograms of the reference on the 1024^2 synthetic textures (the survey's hash values
    could not be reproduced with its stated hash; the histograms and our own reference-derived hashes pin the data)."""
    for entry in known["direct"][:2]:
        img = synth.make_texture(entry["size"], entry["size"], entry["kind"])
        b = port_oracle.encode_blocks(synth.to_blocks(img), default_params(**entry["params"]), threads=os.cpu_count() or 1)
        assert synth.mode_histogram(b) == {int(k): v for k, v in entry["modes"].items()}
        assert "%016x" % synth.fnv1a64_words(b) == entry["fnv1a64"]
    assert known["direct"][0]["modes"] == {"1": 57578, "6": 7958}
    assert known["direct"][1]["modes"] == {"1": 28763, "5": 1992, "6": 17933, "7": 16848}


def test_known_answer_uber4(port_oracle, known):
    entry = known["direct"][5]
    assert entry["modes"] == {"1": 15482, "6": 902}  # SURVEY.md App. C row "direct 512^2 kind 0 uber 4"
    img = synth.make_texture(entry["size"], entry["size"], entry["kind"])
    b = port_oracle.encode_blocks(synth.to_blocks(img), default_params(**entry["params"]), threads=os.cpu_count() or 1)
    assert "%016x" % synth.fnv1a64_words(b) == entry["fnv1a64"]
