#!/usr/bin/env python
"""Order of the key bins in the filterbank regrouping of the opaque BC7 kernels (bc7_core.cuh, kBinOfKey).

At iteration 14 of the partition scan a CTA sorts its 256 blocks by key (the best of the first 14 partitions) and every warp
takes 32 consecutive blocks of that order; in iterations 14..34 a warp scores the UNION of the candidates its blocks still need
(bc7enc.cpp:1786-1799: candidate `it` is needed by key k iff pred[order[it]] has bit k + 1), and the CTA's write-back barrier
waits for the slowest warp.  This tool reads the predictor table from bc7_tables.cpp, simulates the regrouping for a bin order on
key maps (uniform random keys, or files of one key byte per block in launch order as written by a host build of the estimator),
and anneals an order that minimises the mean union per warp.  CPU only.

    bin_order.py                      cost of the iteration order and of the shipped order on uniform keys
    bin_order.py --anneal [files...]  anneal on a mixture of uniform keys and the given key maps, print the constant
"""
import argparse
import os
import random
import re

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def tables():
    src = open(os.path.join(ROOT, "vierkant_b200", "csrc", "bc7_tables.cpp")).read()
    order = [int(x) for x in re.search(r"k_order\[64\] = \{(.*?)\};", src, re.S).group(1).replace("\n", " ").split(",")]
    body = re.search(r"k_pred\[35\] = \{(.*?)\};", src, re.S).group(1).replace("\n", " ")
    pred = eval("[" + body + "]", {"bit": lambda x: 1 << x, "k_all": 0xFFFFFFFF})
    assert len(order) == 64 and len(pred) == 35
    return order, pred


def candidate_sets():
    """cand[k]: bit `it` set iff a block whose key is the partition of iteration k needs iteration it (14..34)."""
    order, pred = tables()
    cand = []
    for k in range(14):
        key = order[k]
        cand.append(sum(1 << it for it in range(14, 35) if pred[order[it]] & (1 << (key + 1))))
    return cand


def shipped_order():
    src = open(os.path.join(ROOT, "vierkant_b200", "csrc", "bc7_core.cuh")).read()
    v = int(re.search(r"kBinOfKey = (0x[0-9A-Fa-f]+)ull", src).group(1), 16)
    return [(v >> (4 * k)) & 15 for k in range(14)]


def cost(perm, ctas, cand):
    """(mean union per warp, mean over CTAs of the largest union) for bin order perm on ctas[n, 256] (keys 0..13, 15 = done)."""
    lut = np.full(16, 15)
    lut[:14] = perm
    srt = np.sort(lut[ctas], axis=1)
    by_bin = np.zeros(16, dtype=np.int64)
    for k in range(14):
        by_bin[perm[k]] = cand[k]
    w = srt.reshape(srt.shape[0], 8, 32)
    u = np.zeros(w.shape[:2], dtype=np.int64)
    for j in range(32):
        u |= by_bin[w[:, :, j]]
    pc = np.zeros_like(u)
    while u.any():
        pc += u & 1
        u >>= 1
    return float(pc.mean()), float(pc.max(axis=1).mean())


def uniform_keys(n=600, seed=0):
    return np.random.default_rng(seed).integers(0, 14, (n, 256)).astype(np.uint8)


def anneal(ctas, cand, seed=1, steps=3000):
    rng = random.Random(seed)
    cur = list(range(14))
    cc = cost(cur, ctas, cand)[0]
    best, bc, t = cur[:], cc, 0.05
    for _ in range(steps):
        a, b = rng.sample(range(14), 2)
        nxt = cur[:]
        nxt[a], nxt[b] = nxt[b], nxt[a]
        c = cost(nxt, ctas, cand)[0]
        if c < cc or rng.random() < np.exp((cc - c) / t):
            cur, cc = nxt, c
            if c < bc:
                best, bc = nxt[:], c
        t *= 0.999
    return best


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--anneal", action="store_true")
    ap.add_argument("files", nargs="*")
    a = ap.parse_args()
    cand = candidate_sets()
    print("candidates per key:", [bin(c).count("1") for c in cand])
    sets = {"uniform": uniform_keys()}
    for f in a.files:
        sets[os.path.basename(f)] = np.fromfile(f, dtype=np.uint8).reshape(-1, 256)
    orders = {"iteration order": list(range(14)), "shipped (kBinOfKey)": shipped_order()}
    if a.anneal:
        rng = np.random.default_rng(1)
        mix = np.concatenate([d[rng.choice(d.shape[0], min(300, d.shape[0]), replace=False)] for d in sets.values()])
        best = anneal(mix, cand, seed=3)
        orders["annealed"] = best
        print("annealed order:", best, "constant: 0x%Xull" % sum(b << (4 * k) for k, b in enumerate(best)))
    for name, p in orders.items():
        print("%-22s" % name, "  ".join("%s: %.2f / %.2f" % (n, *cost(p, d, cand)) for n, d in sets.items()))
