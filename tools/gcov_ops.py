#!/usr/bin/env python
"""gcov_ops.py -- pins the numerator of the ALU roofline: algorithmic scalar ops per block and pixel
for every BASELINE config, from gcov line counts of the REFERENCE's own bc7enc.cpp on that config's own (stbir-filtered)
inputs.  Measurement tooling: compiles /root/reference/extern/bc7enc_rdo/bc7enc.cpp where it lies (-O0 --coverage, into a
temporary directory), so it only runs where the reference tree exists; its output, profiles/gcov_ops.json, is committed and
read by bench.py.

Op model (integer/FP32 arithmetic only, loads / stores / moves excluded, pixel-side and candidate-side YCbCr
hoisted), per block:
    17 D_rgb + 21 D_rgba                    selector-search distance evaluations              (bc7enc.cpp:802, :790)
  + 24 K_rgb + 29 K_rgba                    palette colours built, N per evaluate_solution    (:686 / channels + 2 per call)
  + 37 E1 + 35 E7                           estimator pixel evaluations                       (:1505, :1648)
  + 223 S1 + 6 E1 + 113 S7 + 8 E7           estimator subset set-up (223 + 6n, 113 + 8n)      (:1443, :1581)
  + 125 C + 4 n_mean + 15 n_cov + 62 n_ipca + 13.5 n_mean
                                            colour cell: mean, PCA, projection                (:1149, :1186, :1164)
  + 290 F                                   find_optimal_solution (endpoint quantisation)     (:868)
  + 40 L + 15 n_ls                          least-squares refits (rgb, rgba, a)               (:351/:364, :287/:303, :410/:426)
  + 6 * 16 * 4 A5                           mode-5 alpha level search passes                  (:2092 / 16)
  + 1200                                    per block: gather, pixel YCbCr, dispatch, encode_bc7_block
With the event rates App. D measured on unfiltered 1024^2 inputs this reproduces its totals (4.7e4 opaque / 4.7e4 alpha /
2.1e5 uber 4).  Usage: python tools/gcov_ops.py [--configs c2,c3,c5] [--blocks 120000]
"""
from __future__ import annotations

import argparse
import json
import os
import re
import subprocess
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import pyoracle, slab  # noqa: E402
from vierkant_b200 import synth  # noqa: E402

REF = "/root/reference/extern/bc7enc_rdo"
CONFIGS = {
    "c1": dict(base=1024, kind=0, uber=0, parts=64, fb=1),
    "c2": dict(base=4096, kind=0, uber=0, parts=64, fb=1),
    "c3": dict(base=8192, kind=1, uber=0, parts=64, fb=0),
    "c5": dict(base=16384, kind=0, uber=4, parts=64, fb=0),
    "c4_alpha": dict(base=4096, kind=1, uber=0, parts=64, fb=1),  # the every-fourth texture of the material batch
}
LINES = {"D_rgb": 802, "D_rgba": 790, "pal_ch": 686, "eval": 645, "E1": 1505, "E7": 1648, "S1": 1443, "S7": 1581, "part": 1801,
         "ccc": 1101, "n_mean": 1149, "n_cov": 1186, "n_ipca": 1164, "fos": 868, "ls_rgb": 351, "n_ls_rgb": 364, "ls_rgba": 287,
         "n_ls_rgba": 303, "ls_a": 410, "n_ls_a": 426, "a5_px": 2092}


def sample_tiles(cfg: dict, want_blocks: int, threads: int):
    """Tiles of the config's stbir-filtered chain: the same stride of block rows from every level (>= 1 row per level)."""
    base = cfg["base"]
    oracle = pyoracle.RefOracle()
    # a full-width strip of the texture is enough: rows are statistically alike, and the chain of a strip is the chain's strip
    # (oracle/slab.py); 16384^2 would otherwise take a minute to generate.  Strip = the top 2048 rows (or the whole image).
    rows = min(base, 2048)
    img = synth.make_texture(base, base, cfg["kind"], rows=(0, rows))
    n_levels = max(0, base.bit_length() - 1 - 2) + 1
    tiles, weights = [], []
    prev, w, h = img, base, rows
    total_blocks = sum((max(base >> l, 4) // 4) ** 2 for l in range(n_levels))
    for l in range(n_levels):
        lw = max(base >> l, 4)
        lh = max(h if l == 0 else prev.shape[0] // 2, 4)
        if prev.shape[0] < 8 and l > 0:
            break
        # bottom rows of a strip's level differ from the full image's (edge clamp): drop the last 16 rows of every strip level
        cur = oracle.resize(prev, lw, lh) if lh < 64 else slab.resize_banded(oracle, prev, lw, lh, threads)
        usable = lh if rows == base else max(4, (lh - 16) // 4 * 4)
        t = synth.to_blocks(np.ascontiguousarray(cur[:usable]))
        level_blocks = (lw // 4) ** 2
        take = max(1, int(round(want_blocks * level_blocks / total_blocks)))
        step = max(1, t.shape[0] // take)
        bx = lw // 4
        # whole block rows, evenly spread
        nrows = t.shape[0] // bx
        row_step = max(1, int(round(nrows / max(1, take // bx)))) if take >= bx else nrows
        pick = t.reshape(nrows, bx, 16, 4)[::row_step] if take >= bx else t.reshape(nrows, bx, 16, 4)[:1, ::max(1, bx // take)]
        pick = pick.reshape(-1, 16, 4)
        tiles.append(pick)
        weights.append((l, level_blocks, pick.shape[0]))
        prev = cur
    return np.concatenate(tiles), weights


def gcov_counts(workdir: str) -> dict:
    subprocess.run(["gcov", "-o", workdir, os.path.join(REF, "bc7enc.cpp")], cwd=workdir, check=True, capture_output=True)
    counts = {}
    with open(os.path.join(workdir, "bc7enc.cpp.gcov")) as f:
        for line in f:
            m = re.match(r"\s*([0-9#=\-*]+)\*?:\s*(\d+):", line)
            if m and m.group(1)[0].isdigit():
                counts[int(m.group(2))] = int(m.group(1).rstrip("*"))
    return counts


def run_kind(workdir: str, tiles_path: str, kind: str, cfg: dict) -> tuple[int, dict]:
    for f in os.listdir(workdir):
        if f.endswith(".gcda"):
            os.remove(os.path.join(workdir, f))
    out = subprocess.run([os.path.join(workdir, "harness"), tiles_path, kind, str(cfg["uber"]), str(cfg["parts"]), str(cfg["fb"])],
                         cwd=workdir, check=True, capture_output=True, text=True)
    n = int(out.stdout.split()[0])
    c = gcov_counts(workdir) if n else {}
    return n, {k: c.get(v, 0) for k, v in LINES.items()}


def ops_from_events(n: int, e: dict, alpha: bool) -> tuple[float, dict]:
    """Events per block and ops per block of one block kind."""
    if n == 0:
        return 0.0, {}
    ev = {k: v / n for k, v in e.items()}
    # palette colours: N per evaluate_solution call = interpolated ones (line 686 runs once per channel) + the two endpoints
    if alpha:
        calls5 = ev["D_rgb"] / 64.0                     # mode 5 is the only RGB search of an alpha block: 16 px x 4 colours
        k_rgb = 4.0 * calls5
        k_rgba = (ev["pal_ch"] - 3.0 * 2.0 * calls5) / 4.0 + 2.0 * (ev["eval"] - calls5)
    else:
        k_rgb, k_rgba = ev["pal_ch"] / 3.0 + 2.0 * ev["eval"], 0.0
    ls_calls = ev["ls_rgb"] + ev["ls_rgba"] + ev["ls_a"]
    n_ls = ev["n_ls_rgb"] + ev["n_ls_rgba"] + ev["n_ls_a"]
    parts = {
        "selector_search": 17 * ev["D_rgb"] + 21 * ev["D_rgba"],
        "palette": 24 * k_rgb + 29 * k_rgba,
        "estimator_pixels": 37 * ev["E1"] + 35 * ev["E7"],
        "estimator_setup": 223 * ev["S1"] + 6 * ev["E1"] + 113 * ev["S7"] + 8 * ev["E7"],
        "colour_cell": 125 * ev["ccc"] + 17.5 * ev["n_mean"] + 15 * ev["n_cov"] + 62 * ev["n_ipca"],
        "quantise": 290 * ev["fos"],
        "least_squares": 40 * ls_calls + 15 * n_ls,
        "mode5_alpha": 6 * 4 * ev["a5_px"] if alpha else 0.0,
        "per_block": 1200.0,
    }
    ev.update({"K_rgb": k_rgb, "K_rgba": k_rgba})
    return float(sum(parts.values())), {"events_per_block": {k: round(v, 3) for k, v in ev.items()}, "ops_parts": {k: round(v, 1) for k, v in parts.items()}}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--configs", default="c1,c2,c3,c5,c4_alpha")
    ap.add_argument("--blocks", type=int, default=120000)
    ap.add_argument("--out", default=os.path.join(ROOT, "profiles", "gcov_ops.json"))
    a = ap.parse_args()
    if not os.path.exists(os.path.join(REF, "bc7enc.cpp")):
        raise SystemExit("gcov_ops.py needs the reference tree (/root/reference)")
    pyoracle.build("ref")
    threads = os.cpu_count() or 1
    result = {"method": "gcov line counts of the reference's bc7enc.cpp (-O0 --coverage, compiled in place) on tiles sampled from each config's "
                        "stbir-filtered chain; op model of SURVEY.md App. D (tools/gcov_ops.py docstring)",
              "lines": LINES, "configs": {}}
    with tempfile.TemporaryDirectory() as wd:
        subprocess.run(["g++", "-std=c++17", "-O0", "--coverage", "-w", "-I", REF, "-c", os.path.join(REF, "bc7enc.cpp"), "-o", os.path.join(wd, "bc7enc.o")],
                       check=True, cwd=wd)
        subprocess.run(["g++", "-std=c++17", "-O1", "-w", "-I", REF, os.path.join(ROOT, "tools", "gcov", "gcov_harness.cpp"), os.path.join(wd, "bc7enc.o"),
                        "--coverage", "-o", os.path.join(wd, "harness")], check=True, cwd=wd)
        for name in [c for c in a.configs.split(",") if c]:
            cfg = CONFIGS[name]
            tiles, weights = sample_tiles(cfg, a.blocks, threads)
            path = os.path.join(wd, f"{name}.bin")
            tiles.tofile(path)
            n_o, e_o = run_kind(wd, path, "opaque", cfg)
            n_a, e_a = run_kind(wd, path, "alpha", cfg)
            ops_o, d_o = ops_from_events(n_o, e_o, False)
            ops_a, d_a = ops_from_events(n_a, e_a, True)
            n = n_o + n_a
            ops = (ops_o * n_o + ops_a * n_a) / n
            result["configs"][name] = {
                "input": f"{cfg['base']}x{cfg['base']} synthetic kind {cfg['kind']}, stbir-filtered chain, uber {cfg['uber']}, "
                         f"{cfg['parts']} partitions, filterbank {'on' if cfg['fb'] else 'off'}",
                "blocks": n, "opaque_blocks": n_o, "alpha_blocks": n_a,
                "ops_per_block": ops, "ops_per_pixel": ops / 16.0,
                "opaque": {"ops_per_block": ops_o, **d_o}, "alpha": {"ops_per_block": ops_a, **d_a},
                "levels_sampled": [{"level": l, "level_blocks": lb, "sampled": s} for l, lb, s in weights],
            }
            print(name, f"{n} blocks ({n_a} alpha): {ops:.0f} ops/block = {ops / 16:.0f} ops/pixel (opaque {ops_o:.0f}, alpha {ops_a:.0f})", flush=True)
    cfgs = result["configs"]
    if "c2" in cfgs and "c4_alpha" in cfgs:  # C4: three opaque textures (as C2) to one with alpha gradients, all at defaults
        ops = 0.75 * cfgs["c2"]["ops_per_block"] + 0.25 * cfgs["c4_alpha"]["ops_per_block"]
        cfgs["c4"] = {"input": "material batch: 3 x (4096x4096 kind 0) : 1 x (4096x4096 kind 1), defaults -- 0.75 c2 + 0.25 c4_alpha",
                      "blocks": cfgs["c2"]["blocks"] + cfgs["c4_alpha"]["blocks"], "ops_per_block": ops, "ops_per_pixel": ops / 16.0}
    with open(a.out, "w") as f:
        json.dump(result, f, indent=1)
    print("wrote", a.out)


if __name__ == "__main__":
    main()
