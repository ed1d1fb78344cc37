"""Generate the committed golden vectors from the UNMODIFIED reference (oracle/_ref/libvkt_ref.so).

Run in the build container (where /root/reference exists):   python tests/golden/make_golden.py
Outputs (small, committed):
  tests/golden/bc7_blocks.npz       edge-case tiles + reference BC7 blocks for every parameter case
  tests/golden/known_answers.json   mode histograms / FNV hashes of the reference on the SURVEY.md App. C inputs and on
                                    the reference's own test shapes (tests/TestCompressionBC7.cpp), plus stbir resizes
"""
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.dirname(HERE))

from cases import RDO_CASES, PARAM_CASES, edge_tiles  # noqa: E402
from oracle.pyoracle import RefOracle, build, default_params  # noqa: E402
from vierkant_b200 import synth  # noqa: E402


def main():
    build("ref")
    ref = RefOracle()
    tiles = np.concatenate([edge_tiles(7, 16), synth.to_blocks(synth.make_texture(32, 32, 0)),
                            synth.to_blocks(synth.make_texture(32, 32, 1))])
    out = {"tiles": tiles}
    for name, kw in {**PARAM_CASES, **RDO_CASES}.items():
        out["blocks_" + name] = ref.encode_blocks(tiles, default_params(**kw), threads=4)
    out["bc5_blocks"] = ref.encode_bc5_blocks(tiles)
    out["decoded_defaults"] = ref.unpack_blocks(out["blocks_defaults"])
    np.savez_compressed(os.path.join(HERE, "bc7_blocks.npz"), **out)

    ka = {"direct": [], "compress": [], "reference_tests": [], "resize": []}
    for size, kind, kw in [(1024, 0, {}), (1024, 1, {}), (1024, 0, dict(mode17_partition_estimation_filterbank=0)),
                           (1024, 0, dict(uber_level=4, mode17_partition_estimation_filterbank=0)),
                           (1024, 1, dict(uber_level=4, mode17_partition_estimation_filterbank=0)),
                           (512, 0, dict(uber_level=4, mode17_partition_estimation_filterbank=0))]:
        img = synth.make_texture(size, size, kind)
        b = ref.encode_blocks(synth.to_blocks(img), default_params(**kw), threads=8)
        ka["direct"].append({"size": size, "kind": kind, "params": kw, "blocks": int(b.shape[0]),
                             "fnv1a64": "%016x" % synth.fnv1a64_words(b), "modes": synth.mode_histogram(b)})
    for size, kind in [(1024, 0), (1024, 1), (2048, 0)]:
        r = ref.compress(synth.make_texture(size, size, kind), 1, True, 8)
        allb = np.concatenate(r["levels"])
        ka["compress"].append({"size": size, "kind": kind, "levels": len(r["levels"]), "blocks": int(allb.shape[0]),
                               "fnv1a64": "%016x" % synth.fnv1a64_words(allb)})
    # the reference's own test shapes: 4x4 checkerboard resized to w x h (tests/TestCompressionBC7.cpp:59-131)
    for name, comps, w, h, mode, mips in [("CompressionBC5.basic", 4, 512, 256, 0, False), ("CompressionBC7.basic", 4, 512, 256, 1, False),
                                          ("CompressionBC7.missing_alpha", 3, 64, 128, 1, False), ("CompressionBC7.mips", 4, 512, 256, 1, True),
                                          ("CompressionBC7.odd_size", 4, 123, 81, 1, True)]:
        img = ref.resize(synth.checkerboard_4x4(comps), w, h)
        r = ref.compress(img, mode, mips, 0)
        allb = np.concatenate(r["levels"])
        ka["reference_tests"].append({"name": name, "comps": comps, "w": w, "h": h, "mode": mode, "mips": mips,
                                      "base": [r["base_width"], r["base_height"]], "level_blocks": [int(l.shape[0]) for l in r["levels"]],
                                      "fnv1a64": "%016x" % synth.fnv1a64_words(allb),
                                      "modes": synth.mode_histogram(allb) if mode == 1 else {}})
    # stbir: 1:1 Mitchell, 2:1 Mitchell, odd-size upsample (Catmull-Rom), strong downsample, 3 components
    for kind, w, h, ow, oh, comps in [(0, 64, 64, 64, 64, 4), (1, 64, 64, 32, 32, 4), (0, 123, 81, 124, 84, 4), (1, 256, 128, 16, 8, 4),
                                      (0, 60, 36, 60, 36, 3), (0, 4, 4, 512, 256, 4), (1, 100, 52, 52, 28, 4)]:
        img = synth.make_texture(w, h, kind)[..., :comps]
        o = ref.resize(img, ow, oh)
        ka["resize"].append({"kind": kind, "w": w, "h": h, "ow": ow, "oh": oh, "comps": comps,
                             "fnv1a64_bytes": "%016x" % synth.fnv1a64_words(np.frombuffer(o.tobytes() + b"\0" * (-o.size % 8), dtype=np.uint8))})
    with open(os.path.join(HERE, "known_answers.json"), "w") as f:
        json.dump(ka, f, indent=1)
    print("wrote", os.path.join(HERE, "bc7_blocks.npz"), os.path.getsize(os.path.join(HERE, "bc7_blocks.npz")), "bytes")


if __name__ == "__main__":
    main()
