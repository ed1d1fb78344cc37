"""Shared test inputs: edge-case tile generators and the encoder parameter matrix."""
from __future__ import annotations

import numpy as np

# the knobs of bc7enc_compress_block_params vierkant could reach (SURVEY.md F10, App. E "stage-wise oracles")
PARAM_CASES = {
    "defaults": dict(),
    "filterbank_off": dict(mode17_partition_estimation_filterbank=0),
    "uber1": dict(uber_level=1),
    "uber2": dict(uber_level=2),
    "uber4_fb_off": dict(uber_level=4, mode17_partition_estimation_filterbank=0),
    "partitions16": dict(max_partitions=16),
    "partitions1": dict(max_partitions=1),
    "partitions0": dict(max_partitions=0),
    "no_least_squares": dict(try_least_squares=0),
    "mode6_only": dict(mode_mask=1 << 6),
    "modes_6_1": dict(mode_mask=(1 << 6) | (1 << 1)),
    "modes_5_1": dict(mode_mask=(1 << 5) | (1 << 1)),
    "modes_7_1": dict(mode_mask=(1 << 7) | (1 << 1)),
    "force_alpha": dict(force_alpha=1),
    "linear": dict(perceptual=0, weights=[1, 1, 1, 1]),
    "linear_uber2": dict(perceptual=0, weights=[1, 1, 1, 1], uber_level=2),
    "weights": dict(weights=[64, 128, 8, 77]),
    "pbit1_weight": dict(pbit1_weight=1.5),
    "mode_weights": dict(mode1_error_weight=0.9, mode6_error_weight=1.1, mode5_error_weight=1.2, mode7_error_weight=0.8),
    "bias_mode1_pbits": dict(bias_mode1_pbits=1),
    # per-texel errors above 2^28 (but below 2^31): the kernels' packed 28-bit argmin keys must not be used
    "big_weights": dict(weights=[800, 400, 100, 200]),
    "big_weights_linear": dict(perceptual=0, weights=[9000, 7000, 5000, 6000]),
}
# the three knobs bc7enc_rdo's RDO post-processor drives (bc7enc.cpp:697, :838/:885, :1819); vierkant never sets them.
# Served by the extended kernel variant (kKvExt), whose estimator keeps the reference's early-outs.
RDO_CASES = {
    "low_freq_weight": dict(low_frequency_partition_weight=0.75),
    "quant_mode6": dict(quant_mode6_endpoints=1),
    "force_selectors": dict(force_selectors=1, selectors=list(range(16)), mode_mask=1 << 6),
    "low_freq_weight_fb_off": dict(low_frequency_partition_weight=0.5, mode17_partition_estimation_filterbank=0),
    "low_freq_weight_up": dict(low_frequency_partition_weight=1.5),
    "low_freq_weight_zero": dict(low_frequency_partition_weight=0.0),
    "low_freq_weight_linear": dict(low_frequency_partition_weight=0.6, perceptual=0, weights=[1, 1, 1, 1]),
    "quant_mode6_uber2": dict(quant_mode6_endpoints=1, uber_level=2),
    "quant_mode6_linear": dict(quant_mode6_endpoints=1, perceptual=0, weights=[1, 1, 1, 1]),
    "force_selectors_all_modes": dict(force_selectors=1, selectors=[0, 1, 2, 3, 3, 2, 1, 0, 1, 3, 0, 2, 2, 0, 3, 1]),
    "force_selectors_opaque_modes": dict(force_selectors=1, selectors=[7, 6, 5, 4, 3, 2, 1, 0, 0, 2, 4, 6, 1, 3, 5, 7],
                                         mode_mask=(1 << 6) | (1 << 1)),
    "rdo_all_three": dict(force_selectors=1, selectors=[3, 3, 2, 2, 1, 1, 0, 0, 0, 1, 2, 3, 3, 2, 1, 0], quant_mode6_endpoints=1,
                          low_frequency_partition_weight=0.8, uber_level=1),
}
# parameter sets the C ABI must reject (VKT_BCN_ERR_INVALID)
INVALID_CASES = {
    "selector_beyond_mode1_palette": dict(force_selectors=1, selectors=[8] + [0] * 15, mode_mask=(1 << 6) | (1 << 1)),
    "selector_beyond_mode5_palette": dict(force_selectors=1, selectors=[4] + [0] * 15),
    "negative_low_freq_weight": dict(low_frequency_partition_weight=-0.5),
    "nan_low_freq_weight": dict(low_frequency_partition_weight=float("nan")),
    "uber_level_5": dict(uber_level=5),
    "no_opaque_mode": dict(mode_mask=1 << 5),
    # mode 1 is the only opaque mode but has no partition to try: the reference encodes an uninitialised result (bc7enc.cpp:2336)
    "mode1_without_partitions": dict(mode_mask=(1 << 1) | (1 << 5) | (1 << 7), max_partitions=0),
}


def edge_tiles(seed: int, n: int) -> np.ndarray:
    """(12 * n, 16, 4) uint8 tiles covering the encoder's special paths: noise, solid colours (single-colour tables),
    two-colour blocks (partitions), smooth gradients, grey ramps (degenerate endpoints), saturated / near-black values,
    alpha-only variation (mode 5), each with and without alpha."""
    rng = np.random.default_rng(seed)
    t = []
    t.append(rng.integers(0, 256, (n, 16, 4), dtype=np.uint8))
    o = rng.integers(0, 256, (n, 16, 4), dtype=np.uint8)
    o[..., 3] = 255
    t.append(o)
    s = np.repeat(rng.integers(0, 256, (n, 1, 4), dtype=np.uint8), 16, axis=1)
    t.append(s)
    s2 = s.copy()
    s2[..., 3] = 255
    t.append(s2)
    a = rng.integers(0, 256, (n, 1, 4))
    b = rng.integers(0, 256, (n, 1, 4))
    m = rng.integers(0, 2, (n, 16, 1))
    tc = np.where(m == 1, a, b).astype(np.uint8)
    t.append(tc)
    tc2 = tc.copy()
    tc2[..., 3] = 255
    t.append(tc2)
    base = rng.integers(0, 256, (n, 1, 4))
    dx = rng.integers(-20, 21, (n, 1, 4))
    dy = rng.integers(-20, 21, (n, 1, 4))
    xx = (np.arange(16) % 4).reshape(1, 16, 1)
    yy = (np.arange(16) // 4).reshape(1, 16, 1)
    g = np.clip(base + dx * xx + dy * yy + rng.integers(-2, 3, (n, 16, 4)), 0, 255).astype(np.uint8)
    t.append(g)
    g2 = g.copy()
    g2[..., 3] = 255
    t.append(g2)
    gr = np.clip(rng.integers(0, 256, (n, 1, 1)) + rng.integers(-3, 4, (n, 16, 1)), 0, 255).astype(np.uint8)
    gr = np.repeat(gr, 4, axis=2)
    gr[..., 3] = 255
    t.append(gr)
    t.append(rng.integers(250, 256, (n, 16, 4), dtype=np.uint8))
    t.append(rng.integers(0, 4, (n, 16, 4), dtype=np.uint8))
    av = s.copy()
    av[..., 3] = rng.integers(0, 256, (n, 16), dtype=np.uint8)
    t.append(av)
    return np.concatenate(t)


def tiles_to_image(tiles: np.ndarray, blocks_x: int) -> np.ndarray:
    """Inverse of synth.to_blocks: (n, 16, 4) tiles -> (H, W, 4) image with blocks_x tiles per row (n % blocks_x == 0)."""
    n = tiles.shape[0]
    assert n % blocks_x == 0
    by = n // blocks_x
    return np.ascontiguousarray(tiles.reshape(by, blocks_x, 4, 4, 4).transpose(0, 2, 1, 3, 4).reshape(by * 4, blocks_x * 4, 4))


def random_params(rng):
    """A random valid parameter set over every knob of bc7enc_compress_block_params (the RDO-only ones included)."""
    kw = {}
    opaque = [m for m in (1, 6) if rng.random() < 0.8] or [6]
    alpha = [m for m in (5, 6, 7) if rng.random() < 0.7] or [6]
    kw["mode_mask"] = int(sum(1 << m for m in set(opaque) | set(alpha)))
    kw["max_partitions"] = int(rng.choice([0, 1, 7, 16, 33, 35, 48, 64]))
    if 6 not in opaque and kw["max_partitions"] == 0:
        kw["max_partitions"] = 1  # (mode 1 alone with no partitions is undefined in the reference and refused by the C ABI)
    kw["uber_level"] = int(rng.choice([0, 0, 1, 2, 3, 4]))
    kw["perceptual"] = int(rng.random() < 0.6)
    if kw["perceptual"]:
        kw["weights"] = [int(x) for x in rng.integers(1, 200, 4)]
    else:
        kw["weights"] = [int(x) for x in rng.integers(1, 40, 4)] if rng.random() < 0.8 else [int(x) for x in rng.integers(1000, 9000, 4)]
    kw["try_least_squares"] = int(rng.random() < 0.8)
    kw["mode17_partition_estimation_filterbank"] = int(rng.random() < 0.5)
    kw["force_alpha"] = int(rng.random() < 0.15)
    kw["bias_mode1_pbits"] = int(rng.random() < 0.2)
    kw["pbit1_weight"] = float(rng.choice([1.0, 1.0, 0.7, 1.3, 2.0]))
    for k in ("mode1_error_weight", "mode5_error_weight", "mode6_error_weight", "mode7_error_weight"):
        kw[k] = float(rng.choice([1.0, 1.0, 0.8, 1.25]))
    if rng.random() < 0.25:
        kw["low_frequency_partition_weight"] = float(rng.choice([0.0, 0.5, 0.75, 1.5]))
    if rng.random() < 0.2:
        kw["quant_mode6_endpoints"] = 1
    if rng.random() < 0.15:
        kw["force_selectors"] = 1
        top = 3 if any(m in (5, 7) for m in alpha) else (7 if 1 in opaque else 15)
        kw["selectors"] = [int(x) for x in rng.integers(0, top + 1, 16)]
    return kw
