#!/usr/bin/env python
"""bench.py -- BC7 encode throughput on B200 (BASELINE.json metric: Mpixel/s), beside the reference CPU path.

  python bench.py [--gpus N] [--steps K] [--warmup W]            our arm (CUDA through the C ABI)
  python bench.py --impl reference [--gpus N] [--steps K] ...     the reference's own CPU path on the host cores

Workload (BASELINE.json configs[1]): one 4096x4096 RGBA8 opaque albedo-like texture with its full mip chain
(11 levels, 1 398 101 blocks, 22.37 Mpixel) per GPU and step, default bc7enc parameters.  A "step" is one pass of the
hot path over that chain.  With N > 1 every rank encodes its own chain (independent textures, no collective on the
data path; weak scaling) and `value` is the whole-job pixel rate over the max-over-ranks device time.

  value  : all mip levels already resident in HBM, only our kernels inside the timed region (CUDA events on the
           launching stream).  Four different textures are rotated so the inputs of consecutive steps (358 MB) exceed L2.
  e2e    : the reference-facing call itself, vkt_bcn_cuda_compress == vierkant::bcn::compress(): the 4096x4096 source
           image in pinned host memory in, every level's blocks in pinned host memory out; H2D of the source, the
           stbir-exact resize chain (level 0 included, as the reference does), classification, encode kernels and D2H
           of the blocks are all inside the timed region.  This is what the reference arm (--impl reference, the
           reference's compress() with stbir on the host cores) is compared with.
  roofline     : ALU/issue-slot bound: algorithmic lane-ops per launch / kernel time vs a peak
                 microbenchmarked in this run; HBM GB/s alongside (informational).
  cpu_baseline : the reference (oracle/_ref, unmodified sources) on the host cores, rank 0 at N=1: vierkant::bcn::compress of
                 the SAME chain with all host threads (value), one thread on a 1024^2 crop (n1), and bc7enc_compress_block
                 alone over the pre-filtered levels with a row partition over the threads (blocks_only).
  parity       : the reference's blocks of the whole chain compared block by block with the e2e call's (untimed).
  e2e_pageable : the e2e call once more with pageable host buffers, as the C++ drop-in passes them (informational).
  e2e_dropin   : integration/texture_block_compression_cuda.cpp's vierkant::bcn::compress() itself (crocore image in,
                 compress_result_t with freshly allocated std::vectors out), compiled against the reference's headers.
  strong       : ONE 16384^2 (C5) and ONE 8192^2 (C3) chain sharded by block rows over the N ranks -- bench_strong.py.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

from vierkant_b200 import synth  # noqa: E402

BYTES_PER_PIXEL = 5.0          # 64 B in + 16 B out per 16-pixel block
NCU_TRAFFIC_BYTES = 77.48e6    # measured DRAM bytes of the C2 level-0 launch (ncu --set full, profiles/r2_ff_opaque_ncu_summary.txt)

# BASELINE.json configs.  The default (and the only one the driver runs) is configs[1]; the others are selectable with
# --workload for the numbers quoted in DESIGN.md.  ops_per_pixel: algorithmic scalar ops (gcov
# event counts x per-event costs): opaque defaults 4.7e4 ops/block, alpha defaults 4.7e4, opaque uber 4 + filterbank off
# 2.1e5; c3 (half alpha, filterbank off: ~36 instead of 28.7 partitions scored) ~5.3e4.
WORKLOADS = {
    "c2": dict(base=4096, kind=0, rotate=4, params=dict(), ops_per_pixel=2.9e3,
               name="4096x4096 RGBA8 opaque albedo-like (synthetic kind 0) + full mip chain, default bc7enc params",
               params_name="bc7enc defaults (perceptual, 64 partitions, filterbank on, uber 0)"),
    "c1": dict(base=1024, kind=0, rotate=16, params=dict(), ops_per_pixel=2.9e3,
               name="1024x1024 RGBA8 noise+gradient (synthetic kind 0) + full mip chain, default bc7enc params",
               params_name="bc7enc defaults (perceptual, 64 partitions, filterbank on, uber 0)"),
    "c3": dict(base=8192, kind=1, rotate=1, params=dict(max_partitions=64, mode17_partition_estimation_filterbank=0),
               ops_per_pixel=3.3e3,
               name="8192x8192 RGBA8 with alpha gradients (synthetic kind 1: modes 1/5/6/7) + full mip chain, max partitions",
               params_name="perceptual, 64 partitions, filterbank off (every pattern is a candidate), uber 0"),
    "c4": dict(base=4096, kind=0, rotate=1, textures=8, params=dict(), ops_per_pixel=2.9e3,
               name="glTF-like material batch: 8 x 4096x4096 RGBA8 textures + mips per GPU (64 over 8 GPUs), every 4th with alpha "
                    "(synthetic kind 1), all encoded as BC7",
               params_name="bc7enc defaults (perceptual, 64 partitions, filterbank on, uber 0)"),
    "c5": dict(base=16384, kind=0, rotate=1,
               params=dict(uber_level=4, max_partitions=64, mode17_partition_estimation_filterbank=0), ops_per_pixel=1.3e4,
               name="16384x16384 RGBA8 opaque (synthetic kind 0) + full mip chain, highest quality",
               params_name="perceptual, uber level 4 (BC7ENC_MAX_UBER_LEVEL), 64 partitions, filterbank off"),
}
BASE = 4096
KIND = 0
WORKLOAD = WORKLOADS["c2"]["name"]
OPS_PER_PIXEL = WORKLOADS["c2"]["ops_per_pixel"]
ROTATE = 4


def chain_dims(base: int):
    """Level sizes of vierkant::bcn::compress (texture_block_compression.cpp:80-86,141-146)."""
    r4 = lambda v: (v + 3) & ~3
    w = h = r4(base)
    levels = max(0, int(np.log2(max(w, h)) - 2)) + 1
    out = []
    for _ in range(levels):
        out.append((w, h))
        w, h = r4(max(w // 2, 1)), r4(max(h // 2, 1))
    return out


def load_peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return json.load(f), "measured"
    except Exception:
        return {"hbm_gbs": 6650.0, "sm_max_mhz": 1965.0}, "fallback"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.index = index
        self.proc = None
        self.lines = []
        self.first = 0

    def start(self, wait_s: float = 3.0):
        """Start polling (every 20 ms) and wait until the first sample has arrived: nvidia-smi can take longer to come up
        than a short timed region lasts.  Call it before the warm-up and mark() right before the timed region."""
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "20",
                                          "-i", str(self.index)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
            t0 = time.time()
            while not self.lines and time.time() - t0 < wait_s and self.proc.poll() is None:
                time.sleep(0.01)
        except Exception:
            self.proc = None

    def mark(self):
        """Samples from here on belong to the timed region (earlier ones: warm-up, kept only if the region gets none)."""
        self.first = len(self.lines)

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self) -> dict:
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for line in (self.lines[self.first:] or self.lines):
            f = [x.strip() for x in line.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1]))
                mx.append(float(f[2]))
            except ValueError:
                continue
            for name, v in zip(names, f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "samples": len(sm), "reasons": sorted(reasons)}


def bind_near_gpu(local: int):
    """Pin this rank's host threads (and therefore its first-touch pinned buffers) to the CPUs NVML reports as local to
    its GPU: with one process per GPU the H2D / D2H streams then stay on the GPU's own NUMA node instead of crossing the
    socket interconnect.  Returns a short description for the JSON line; never fatal."""
    try:
        import pynvml
        import torch
        pynvml.nvmlInit()
        try:
            pr = torch.cuda.get_device_properties(local)
            h = pynvml.nvmlDeviceGetHandleByPciBusId(f"{pr.pci_domain_id:08x}:{pr.pci_bus_id:02x}:{pr.pci_device_id:02x}.0")
        except Exception:
            h = pynvml.nvmlDeviceGetHandleByIndex(local)
        ncpu = os.cpu_count() or 1
        mask = pynvml.nvmlDeviceGetCpuAffinity(h, (ncpu + 63) // 64)
        cpus = [i for i in range(ncpu) if (mask[i // 64] >> (i % 64)) & 1]
        allowed = sorted(set(cpus) & set(os.sched_getaffinity(0)))
        if allowed:
            os.sched_setaffinity(0, allowed)
            return f"{len(allowed)} CPUs local to the GPU ({allowed[0]}..{allowed[-1]})"
        return "no local CPU set reported"
    except Exception as e:  # noqa: BLE001
        return f"not bound ({type(e).__name__})"


def load_ops_per_pixel(workload: str, default: float):
    """Algorithmic scalar ops per pixel of a config (SURVEY.md 8d op model x gcov event rates of the reference on this
    config's own inputs): profiles/gcov_ops.json, written by tools/gcov_ops.py.  Falls back to the survey's figure."""
    try:
        with open(os.path.join(ROOT, "profiles", "gcov_ops.json")) as f:
            e = json.load(f)["configs"][workload]
        return float(e["ops_per_pixel"]), f"profiles/gcov_ops.json ({e['blocks']} blocks of {e['input']}: {e['ops_per_block']:.0f} ops/block)"
    except Exception:
        return default, "SURVEY.md App. D (gcov on unfiltered 1024^2 inputs)"


def reference_oracle():
    """(oracle, kind): the unmodified reference compiled in place (oracle/_ref) or, where it is absent, the C port."""
    from oracle import pyoracle
    if pyoracle.RefOracle.available():
        return pyoracle.RefOracle(), "reference"
    pyoracle.build("port")
    return pyoracle.PortOracle(), "port"


def reference_chain(oracle, kind: str, tex: np.ndarray, cores: int):
    """vierkant::bcn::compress(image, BC7, generate_mipmaps) of the reference with a ThreadPoolClassic(cores) delegate, as
    model::compress_textures calls it (src/model/model_loading.cpp:110-118); the C port mirrors it where _ref is absent."""
    if kind == "reference":
        return oracle.compress(tex, 1, True, cores)["levels"]
    return oracle.compress(tex, 1, True, threads=cores)["levels"]


def cpu_baseline_full(tex: np.ndarray, filtered_levels: list, reps: int = 2):
    """The reference's CPU path on the host cores for the bench's own chain: all threads, one thread
    (bounded: a 1024^2 crop + mips) and blocks only.  Returns (dict for the JSON line, reference blocks per level)."""
    oracle, kind = reference_oracle()
    cores = oracle.hardware_concurrency() if kind == "reference" else (os.cpu_count() or 1)
    h, w, _ = tex.shape
    npix = sum(a * b for a, b in chain_dims(w))
    best, levels = None, None
    for _ in range(reps):
        t0 = time.perf_counter()
        levels = reference_chain(oracle, kind, tex, cores)
        dt = time.perf_counter() - t0
        best = dt if best is None else min(best, dt)
    crop = np.ascontiguousarray(tex[:1024, :1024])
    t0 = time.perf_counter()
    reference_chain(oracle, kind, crop, 1)
    n1_s = time.perf_counter() - t0
    n1_pix = sum(a * b for a, b in chain_dims(1024))
    tiles = [synth.to_blocks(np.ascontiguousarray(lv)) for lv in filtered_levels]
    t0 = time.perf_counter()
    blocks_only = [oracle.encode_blocks(t, None, threads=cores) for t in tiles]
    bo_s = time.perf_counter() - t0
    same = all(np.array_equal(a, b) for a, b in zip(blocks_only, levels))
    name = "vierkant::bcn::compress (unmodified reference, oracle/_ref)" if kind == "reference" else "C port of vierkant::bcn::compress (oracle/*.c)"
    info = {"value": npix / best * 1e-6, "unit": "Mpixel/s", "cores": cores, "kind": kind,
            "sample": f"{name} of the whole {w}x{h} chain ({npix * 1e-6:.2f} Mpixel, stbir included), thread-pool delegate with {cores} threads, best of {reps}",
            "seconds": best,
            "n1": {"value": n1_pix / n1_s * 1e-6, "unit": "Mpixel/s", "cores": 1,
                   "sample": f"the same call, no delegate (one thread), on the top-left 1024x1024 crop + mips ({n1_pix * 1e-6:.2f} Mpixel)"},
            "blocks_only": {"value": npix / bo_s * 1e-6, "unit": "Mpixel/s", "cores": cores,
                            "sample": "bc7enc_compress_block over every block of the pre-filtered levels (the GPU path's own stbir-exact levels), "
                                      f"row partition over {cores} threads; blocks equal compress()'s: {same}"}}
    return info, levels


def run_reference(args):
    """--impl reference: the reference's own CPU implementation on the SAME config -- each step is vierkant::bcn::compress
    of the whole workload chain, all host threads (stbir included, as the reference does)."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    steps, warmup = args.steps, args.warmup
    wl = WORKLOADS[args.workload]
    base = args.base or wl["base"]
    if wl["params"]:
        raise SystemExit("--impl reference: vierkant::bcn::compress() has no parameter argument; configs with non-default bc7enc parameters "
                         "are covered by bench.py's strong.*.parity.reference.cpu_blocks_only")
    tex = np.ascontiguousarray(synth.make_texture(base, base, wl["kind"], seed=0xB200))
    npix = sum(w * h for w, h in chain_dims(base))
    oracle, kind = reference_oracle()
    cores = oracle.hardware_concurrency() if kind == "reference" else (os.cpu_count() or 1)
    step = lambda: reference_chain(oracle, kind, tex, cores)
    for _ in range(warmup):
        step()
    t0 = time.perf_counter()
    for _ in range(steps):
        step()
    dt = time.perf_counter() - t0
    v = npix * steps / dt * 1e-6
    sample = (f"each step = vierkant::bcn::compress of the whole workload chain ({base}x{base} + mips, {npix * 1e-6:.2f} Mpixel; stbir included "
              f"as in the reference), ThreadPoolClassic delegate, {cores} host threads" if kind == "reference" else
              f"each step = the C port of vierkant::bcn::compress over the whole workload chain, {cores} host threads")
    workload = wl["name"] if base == wl["base"] else wl["name"].replace(f"{wl['base']}x{wl['base']}", f"{base}x{base}")
    print(json.dumps({
        "impl": "reference", "metric": "bc7_encode_mpixel_per_s", "value": v, "unit": "Mpixel/s", "n_gpus": args.gpus,
        "steps": steps, "warmup": warmup, "ms_per_step": dt / steps * 1e3, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "u8/int32+f32", "data": "synthetic",
        "config": {"workload": workload, "levels": len(chain_dims(base)), "mpixel_per_step": npix * 1e-6, "sample": sample,
                   "same_config": True},
        "cpu_baseline": {"value": v, "unit": "Mpixel/s", "cores": cores, "kind": kind, "sample": sample},
        "e2e": {"value": v, "unit": "Mpixel/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="c2", choices=sorted(WORKLOADS), help="BASELINE.json config (default c2 = configs[1])")
    ap.add_argument("--base", type=int, default=None, help="override the workload's level-0 size (experiments only)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--strong", default="c5,c3", help="strong-scaling configs to run after the headline (comma list of c5, c3; '' = none)")
    ap.add_argument("--strong-base", type=int, default=None, help="override the strong configs' level-0 size (experiments only)")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
        return
    args.warmup = max(args.warmup, 3)

    import torch
    import torch.distributed as dist
    from vierkant_b200 import capi

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the encoder has no CPU path")
    torch.cuda.set_device(local)
    affinity = bind_near_gpu(local) if world > 1 else "single process: not bound"
    dev = torch.device("cuda", local)
    if world > 1:
        # NCCL announces its version on stdout when the first communicator is made: keep stdout for the one JSON line
        sys.stdout.flush()
        saved = os.dup(1)
        os.dup2(2, 1)
        try:
            dist.init_process_group("nccl", device_id=dev)
            warm = torch.zeros(1, device=dev)
            dist.all_reduce(warm)
            torch.cuda.synchronize()
        finally:
            sys.stdout.flush()
            os.dup2(saved, 1)
            os.close(saved)

    wl = WORKLOADS[args.workload]
    base = args.base or wl["base"]
    KIND, ROTATE, OPS_PER_PIXEL = wl["kind"], wl["rotate"], wl["ops_per_pixel"]
    NTEX = wl.get("textures", 1)  # textures per GPU and step (c4: a material batch, encoded one compress() call each)
    WORKLOAD = wl["name"] if base == wl["base"] else wl["name"].replace(f"{wl['base']}x{wl['base']}", f"{base}x{base}")
    ctx = capi.BcnContext([local])
    params = capi.default_params(**wl["params"])
    dims = chain_dims(base)
    npix = sum(w * h for w, h in dims) * NTEX
    nblocks = sum((w // 4) * (h // 4) for w, h in dims) * NTEX

    # Inputs.  Source textures come from the App. C generator (per-rank seeds).  The resident arm (`value`) encodes the
    # stbir-filtered levels of those chains: made here, untimed, by the library's own bit-exact resize
    # (vkt_bcn_cuda_resize_u8, the path tests/test_chain_gpu.py pins to the reference) -- level 0 is the 1:1 Mitchell pass of
    # the source, level l the 2:1 pass of level l-1, exactly as compress() derives them.
    host_src = []      # [texture] pinned source image (what the e2e call is given)
    host_levels = []   # [texture][level] filtered level images (numpy)
    dev_levels = []
    for r in range(ROTATE * NTEX):
        kind = KIND if NTEX == 1 else (1 if (r % 4 == 3) else 0)
        img = synth.make_texture(dims[0][0], dims[0][1], kind, seed=0xB200 + 16 * rank + r)
        host_src.append(torch.from_numpy(img).pin_memory())
        hl, dl, prev = [], [], img
        for (w, h) in dims:
            prev = ctx.resize_u8(prev, w, h)
            hl.append(prev)
            dl.append(torch.from_numpy(prev).to(dev))
        host_levels.append(hl)
        dev_levels.append(dl)
    dev_out = [[torch.empty(((w // 4) * (h // 4), 16), dtype=torch.uint8, device=dev) for (w, h) in dims] for _ in range(NTEX)]
    host_out = [torch.empty(((w // 4) * (h // 4), 16), dtype=torch.uint8).pin_memory() for (w, h) in dims]
    torch.cuda.synchronize()
    stream = torch.cuda.current_stream().cuda_stream

    # one device batch per rotation: the levels of its NTEX textures
    dev_batches = []
    for r in range(ROTATE):
        ims, outs = [], []
        for j in range(NTEX):
            ims += [(t, w, h, 4) for t, (w, h) in zip(dev_levels[r * NTEX + j], dims)]
            outs += dev_out[j]
        dev_batches.append(ctx.make_device_batch(ims, outs))

    # Consecutive steps are independent chains: they are submitted on two streams in turn, as vkt_bcn_cuda_compress_batch submits
    # the textures of a material on its two lanes, so the drain of one chain's last CTAs overlaps the head of the next chain
    # instead of idling the device (a launch on its own ends with ~0.2 ms of falling occupancy).  Every output buffer set
    # belongs to one stream (dev_out_s), so overlapping steps never write the same memory.
    lanes = [torch.cuda.Stream(device=dev), torch.cuda.Stream(device=dev)]
    dev_out_s = [dev_out, [[torch.empty_like(t) for t in lv] for lv in dev_out]]
    dev_batches_s = [dev_batches]
    alt = []
    for r in range(ROTATE):
        ims, outs = [], []
        for j in range(NTEX):
            ims += [(t, w, h, 4) for t, (w, h) in zip(dev_levels[r * NTEX + j], dims)]
            outs += dev_out_s[1][j]
        alt.append(ctx.make_device_batch(ims, outs))
    dev_batches_s.append(alt)

    def step_device(i):
        # all 11 levels of the chain in one call: one classify + two encode launches (vkt_bcn_cuda_encode_batch_device)
        ctx.encode_batch_device(capi.MODE_BC7, dev_batches_s[i & 1][i % ROTATE], params, 0, lanes[i & 1].cuda_stream)

    import ctypes as C
    out_ptrs = (C.c_void_p * len(dims))(*[t.data_ptr() for t in host_out])

    batch_srcs = None
    if NTEX > 1:  # a material batch goes through vkt_bcn_cuda_compress_batch: whole chains, pipelined over two lanes per device
        host_out_b = [[torch.empty(((w // 4) * (h // 4), 16), dtype=torch.uint8).pin_memory() for (w, h) in dims] for _ in range(NTEX)]
        out_ptrs_b = [(C.c_void_p * len(dims))(*[t.data_ptr() for t in ho]) for ho in host_out_b]
        batch_srcs = []
        for r in range(ROTATE):
            arr = (capi.Source * NTEX)()
            for j in range(NTEX):
                arr[j] = capi.Source(host_src[r * NTEX + j].data_ptr(), dims[0][0], dims[0][1], 4, capi.MODE_BC7, out_ptrs_b[j])
            batch_srcs.append(arr)

    def step_e2e(i):
        if batch_srcs is not None:
            ctx._check(ctx.lib.vkt_bcn_cuda_compress_batch(ctx.handle, batch_srcs[i % ROTATE], NTEX, 1, C.byref(params)))
            return
        src = host_src[i % ROTATE]
        ctx._check(ctx.lib.vkt_bcn_cuda_compress(ctx.handle, capi.MODE_BC7, src.data_ptr(), dims[0][0], dims[0][1], 4, 1,
                                                 C.byref(params), out_ptrs))

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- kernel-only ------------------------------------------------------------------------------------------
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()  # (before the warm-up: the poller is up and sampling when the timed region begins)
    for i in range(args.warmup):
        step_device(i)
    barrier()
    s0 = ctx.stats()
    sampler.mark()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    cur = torch.cuda.current_stream()
    e0.record(cur)  # the device is idle here (barrier above): both lanes start after this point
    for ln in lanes:
        ln.wait_stream(cur)
    for i in range(args.steps):
        step_device(i)
    for ln in lanes:
        cur.wait_stream(ln)
    e1.record(cur)  # after the last kernel of either lane
    barrier()
    ms = e0.elapsed_time(e1)
    clocks = sampler.stop() if rank == 0 else None
    s1 = ctx.stats()
    launches = s1["kernel_launches"] - s0["kernel_launches"]

    # dominant kernel alone: the level-0 launch, timed with its own event pair
    k0, k1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    kms = []
    for i in range(max(3, min(args.steps, 10))):
        lv = dev_levels[(i % ROTATE) * NTEX]
        torch.cuda.synchronize()
        k0.record()
        ctx.encode_bc7_device(lv[0], dims[0][0], dims[0][1], 4, dev_out[0][0], params, 0, stream)
        k1.record()
        torch.cuda.synchronize()
        kms.append(k0.elapsed_time(k1))
    kernel_ms = float(np.mean(kms))

    # ---- end to end through the host-buffer C ABI ----------------------------------------------------------------
    for i in range(args.warmup):
        step_e2e(i)
    barrier()
    b0 = ctx.stats()
    t0 = time.perf_counter()
    for i in range(args.steps):
        step_e2e(i)
    torch.cuda.synchronize()
    e2e_s = time.perf_counter() - t0
    b1 = ctx.stats()
    barrier()

    # the same call with DEVICE destinations (the blocks land in device memory -- an imported Vulkan staging
    # buffer -- and never cross to the host): only the source upload uses the link.  Informational; all ranks at once.
    devdst_ms = None
    if batch_srcs is None:
        dd_ptrs = (C.c_void_p * len(dims))(*[t.data_ptr() for t in dev_out[0]])

        def step_devdst(i):
            ctx._check(ctx.lib.vkt_bcn_cuda_compress(ctx.handle, capi.MODE_BC7, host_src[i % ROTATE].data_ptr(), dims[0][0], dims[0][1], 4, 1,
                                                     C.byref(params), dd_ptrs))
        for i in range(args.warmup):
            step_devdst(i)
        barrier()
        t0 = time.perf_counter()
        for i in range(args.steps):
            step_devdst(i)
        torch.cuda.synchronize()
        devdst_ms = (time.perf_counter() - t0) * 1e3 / args.steps
        barrier()

    # the same call with PAGEABLE host buffers -- what the C++ drop-in passes (a malloc'ed crocore image in, std::vector blocks out):
    # informational, N = 1 / single-texture workloads only, outside every other timed region
    pageable_ms = None
    if world == 1 and batch_srcs is None and dims[0][0] <= 8192:
        p_src = torch.from_numpy(host_src[0].numpy().copy())  # plain (unpinned) host memory
        p_outs = [torch.empty_like(o, pin_memory=False) for o in host_out]
        p_ptrs = (C.c_void_p * len(p_outs))(*[o.data_ptr() for o in p_outs])
        for i in range(args.warmup + args.steps):
            if i == args.warmup:
                t0 = time.perf_counter()
            ctx._check(ctx.lib.vkt_bcn_cuda_compress(ctx.handle, capi.MODE_BC7, p_src.data_ptr(), dims[0][0], dims[0][1], 4, 1,
                                                     C.byref(params), p_ptrs))
        pageable_ms = (time.perf_counter() - t0) * 1e3 / args.steps

    # the C++ drop-in itself: integration/texture_block_compression_cuda.cpp's vierkant::bcn::compress(), compiled against the
    # reference's headers -- a crocore image wrapping a malloc'ed buffer in, a compress_result_t with freshly allocated
    # std::vector<block_t> levels out (allocation, zero fill and release of the result are part of every call, as in vierkant)
    dropin = None
    dropin_so = os.path.join(ROOT, "integration", "_build", "libvkt_dropin_test.so")
    if world == 1 and batch_srcs is None and os.path.exists(dropin_so):
        try:
            os.environ["VIERKANT_BCN_CUDA_DEVICES"] = str(local)  # the drop-in's process-wide context: this GPU only
            D = C.CDLL(dropin_so)
            D.dropin_compress.argtypes = [C.c_void_p, C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint32, C.c_int]
            D.dropin_compress.restype = C.c_void_p
            D.dropin_result_free.argtypes = [C.c_void_p]
            D.dropin_result_duration_ms.argtypes, D.dropin_result_duration_ms.restype = [C.c_void_p], C.c_int64
            D.dropin_result_level_data.argtypes, D.dropin_result_level_data.restype = [C.c_void_p, C.c_uint32], C.c_void_p
            srcs = [host_src[r].numpy().copy() for r in range(ROTATE)]
            last = None
            d_warm = max(args.warmup, 6)  # the first calls also grow the process heap the result vectors come from
            for i in range(d_warm + args.steps):
                if i == d_warm:
                    t0 = time.perf_counter()
                r = D.dropin_compress(srcs[i % ROTATE].ctypes.data, dims[0][0], dims[0][1], 4, capi.MODE_BC7, 1)
                if i + 1 == d_warm + args.steps:
                    last = r
                else:
                    D.dropin_result_free(r)
            d_ms = (time.perf_counter() - t0) * 1e3 / args.steps
            n0 = (dims[0][0] // 4) * (dims[0][1] // 4)
            lvl0 = np.frombuffer((C.c_uint8 * (16 * n0)).from_address(D.dropin_result_level_data(last, 0)), dtype=np.uint8).reshape(n0, 16).copy()
            dropin = {"value": npix / (d_ms * 1e-3) * 1e-6, "unit": "Mpixel/s", "ms_per_step": d_ms, "duration_field_ms": int(D.dropin_result_duration_ms(last)),
                      "level0": lvl0, "texture": (d_warm + args.steps - 1) % ROTATE, "warmup": d_warm,
                      "api": "vierkant::bcn::compress(compress_info_t) of integration/texture_block_compression_cuda.cpp (libvkt_dropin_test.so): "
                             "pageable crocore::Image in, compress_result_t (new std::vector<block_t> per level) out, result released each step"}
            D.dropin_result_free(last)
        except Exception as e:  # noqa: BLE001
            dropin = {"value": None, "error": repr(e)}

    tms = torch.tensor([ms, e2e_s * 1e3, devdst_ms or 0.0], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(tms, op=dist.ReduceOp.MAX)
    ms_max, e2e_ms_max, devdst_ms_max = float(tms[0]), float(tms[1]), float(tms[2])

    # ---- ONE chain sharded over the ranks (north_star: 16K-class strong scaling) ----------------------------------------
    strong = {}
    for name in [x for x in args.strong.split(",") if x]:
        import bench_strong
        if name not in bench_strong.STRONG:
            continue
        # release the headline's buffers first?  They are small (0.6 GB); the C5 chain needs ~4 GB of the 180 GB per GPU.
        try:
            rep = bench_strong.run_strong(name, ctx, rank, world, dev, args, dist, base=args.strong_base, check=not args.no_cpu_baseline,
                                          clock_sampler=ClockSampler(local) if rank == 0 else None)
        except Exception as e:  # noqa: BLE001 -- the strong arm must never take the headline line down
            rep = {"error": repr(e)}
            if rank == 0:
                strong[name] = rep
            break  # the ranks are out of step now: no further collective work
        if rank == 0:
            strong[name] = rep

    if rank == 0:
        peaks, peak_src = load_peaks()
        total_pix = npix * world * args.steps
        value = total_pix / (ms_max * 1e-3) * 1e-6
        e2e_value = total_pix / (e2e_ms_max * 1e-3) * 1e-6
        props = torch.cuda.get_device_properties(local)
        sm_mhz = (clocks or {}).get("sm_mhz") or peaks.get("sm_max_mhz", 1965.0)
        # issue-slot peak: SMs x 4 schedulers x 32 lanes x clock, at the max SM clock (conservative
        # denominator) -- replaced by the microbenchmarked figure when the library provides one
        alu_peak = props.multi_processor_count * 4 * 32 * peaks.get("sm_max_mhz", 1965.0) * 1e6
        try:
            probe = ctx.measure_issue_peak(0)
        except Exception:
            probe = None
        l0_pix = dims[0][0] * dims[0][1]
        OPS_PER_PIXEL, ops_source = load_ops_per_pixel(args.workload, OPS_PER_PIXEL)
        achieved = l0_pix * OPS_PER_PIXEL / (kernel_ms * 1e-3)
        line = {
            "metric": "bc7_encode_mpixel_per_s", "value": value, "unit": "Mpixel/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms_max / args.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "u8/int32+f32", "data": "synthetic",
            "config": {"workload": WORKLOAD, "levels": len(dims), "textures_per_gpu": NTEX, "blocks_per_step_per_gpu": nblocks,
                       "mpixel_per_step_per_gpu": npix * 1e-6, "partitioning": f"{world * NTEX} independent texture chains, {NTEX} per GPU, no collective",
                       "streams": "resident arm: independent chains alternate between two streams (as compress_batch's two lanes do); timed with CUDA events that bracket both",
                       "l2": f"{ROTATE * NTEX} textures rotated: {ROTATE * npix * 4 / 1e6:.0f} MB of inputs > 126 MB L2",
                       "resident_inputs": "the stbir-filtered levels of each texture's chain (level 0 = 1:1 Mitchell pass, level l from level l-1), made untimed by vkt_bcn_cuda_resize_u8",
                       "params": wl["params_name"], "host_affinity": affinity},
            "e2e": {"value": e2e_value, "unit": "Mpixel/s", "h2d_bytes_per_step": (b1["h2d_bytes"] - b0["h2d_bytes"]) // args.steps,
                    "d2h_bytes_per_step": (b1["d2h_bytes"] - b0["d2h_bytes"]) // args.steps, "ms_per_step": e2e_ms_max / args.steps,
                    "api": ("vkt_bcn_cuda_compress == vierkant::bcn::compress(): pinned host source image in, stbir-exact resize chain + "
                            "classify + encode on the GPU, pinned host blocks of all levels out") if NTEX == 1 else
                           ("vkt_bcn_cuda_compress_batch: the textures of the batch, each as in vierkant::bcn::compress() (pinned source in, "
                            "resize chain + encode on the GPU, pinned blocks out), pipelined over two lanes per device")},
            "e2e_pageable": (None if pageable_ms is None else
                             {"value": npix / (pageable_ms * 1e-3) * 1e-6, "unit": "Mpixel/s", "ms_per_step": pageable_ms,
                              "note": "same call, pageable host buffers as the C++ drop-in passes them (staged by the library); informational"}),
            "e2e_device_destinations": (None if not devdst_ms_max else
                                        {"value": npix * world / (devdst_ms_max * 1e-3) * 1e-6, "unit": "Mpixel/s", "ms_per_step": devdst_ms_max,
                                         "note": "same call, pinned host source in, every level's blocks left in device memory (the N4 hand-off: an imported "
                                                 "Vulkan staging buffer as destination); no D2H; informational"}),
            "gpu_launches": int(launches),
            "clocks": clocks,
            "roofline": {"bound": "alu", "achieved": achieved * 1e-12, "peak": alu_peak * 1e-12, "unit": "Tlane-op/s",
                         "frac": achieved / alu_peak,
                         # dram__bytes_read.sum + dram__bytes_write.sum of the level-0 launch, one `ncu --set full` capture
                         "traffic": NCU_TRAFFIC_BYTES if (args.workload == "c2" and base == wl["base"]) else None,
                         "traffic_source": "profiles/r2_ff_opaque_ncu_summary.txt (73.28 MB read + 4.21 MB written; algorithmic 83.9 MB, part of the writes still in L2)",
                         "kernel": f"bc7_encode_kernel<perceptual> launch set on level 0 ({base}x{base})", "kernel_ms": kernel_ms,
                         "ops_per_pixel": OPS_PER_PIXEL, "ops_source": ops_source,
                         "peak_source": f"{props.multi_processor_count} SMs x 4 x 32 lanes x {peaks.get('sm_max_mhz', 1965.0):.0f} MHz ({peak_src} clock)",
                         "sm_mhz_during_run": sm_mhz,
                         "issue_probe_tlaneops": None if probe is None else probe * 1e-12,
                         "frac_of_probe": None if not probe else achieved / probe,
                         "hbm": {"achieved_gbs": l0_pix * BYTES_PER_PIXEL / (kernel_ms * 1e-3) * 1e-9, "peak_gbs": peaks.get("hbm_gbs"),
                                 "note": f"informational, {peak_src}"}},
        }
        lvl0 = None
        if dropin is not None:
            lvl0 = dropin.pop("level0", None)
            line["e2e_dropin"] = dropin
        if strong:
            line["strong"] = strong
        if world == 1 and not args.no_cpu_baseline and NTEX == 1 and not wl["params"] and dims[0][0] <= 4096:
            try:
                tex0 = host_src[0].numpy()
                info, ref_levels = cpu_baseline_full(tex0, host_levels[0])
                line["cpu_baseline"] = info
                # the reference's blocks of the WHOLE chain double as the in-run parity check (untimed): the same texture through
                # the same C-ABI call the e2e figure times, compared block by block
                _, ours = ctx.compress(tex0, capi.MODE_BC7, True, params)
                total = sum(int(a.shape[0]) for a in ref_levels)
                bad = sum(int((np.asarray(a).reshape(-1, 16) != np.asarray(b).reshape(-1, 16)).any(axis=1).sum())
                          for a, b in zip(ours, ref_levels))
                line["parity"] = {"against": ("unmodified reference (oracle/_ref)" if info["kind"] == "reference" else "C port (oracle/*.c)") +
                                             ", vierkant::bcn::compress of the whole workload chain",
                                  "levels": len(ref_levels), "blocks": total, "mismatched_blocks": bad,
                                  "bit_exact": bool(bad == 0 and len(ours) == len(ref_levels))}
                # the resident arm's outputs (rotation 0 was encoded last at steps % ROTATE == 1 ... re-encode to be sure)
                ctx.encode_batch_device(capi.MODE_BC7, dev_batches[0], params, 0, stream)
                torch.cuda.synchronize()
                rbad = sum(int((o.cpu().numpy() != np.asarray(b).reshape(-1, 16)).any(axis=1).sum()) for o, b in zip(dev_out[0], ref_levels))
                line["parity"]["resident_arm_mismatched_blocks"] = rbad
                if "e2e_dropin" in line and lvl0 is not None and line["e2e_dropin"].get("texture") == 0:
                    line["e2e_dropin"]["level0_matches_reference"] = bool(np.array_equal(lvl0, ref_levels[0]))
            except Exception as e:  # the baseline must never take the bench down
                line["cpu_baseline"] = {"value": None, "unit": "Mpixel/s", "cores": 0, "kind": "unavailable", "sample": repr(e)}
        if world == 1 and not args.no_cpu_baseline and "parity" not in line:
            # the other workloads (--workload c3 | c4 | c5): the same untimed check of the e2e call's blocks against the reference
            try:
                if NTEX > 1:  # c4: every texture of the batch through vkt_bcn_cuda_compress_batch vs the reference's compress()
                    oracle, kind = reference_oracle()
                    cores = oracle.hardware_concurrency() if kind == "reference" else (os.cpu_count() or 1)
                    ctx._check(ctx.lib.vkt_bcn_cuda_compress_batch(ctx.handle, batch_srcs[0], NTEX, 1, C.byref(params)))
                    total = bad = 0
                    for j in range(NTEX):
                        want = reference_chain(oracle, kind, host_src[j].numpy(), cores)
                        total += sum(int(a.shape[0]) for a in want)
                        bad += sum(int((o.numpy() != np.asarray(b).reshape(-1, 16)).any(axis=1).sum()) for o, b in zip(host_out_b[j], want))
                    line["parity"] = {"against": ("unmodified reference (oracle/_ref)" if kind == "reference" else "C port (oracle/*.c)") +
                                                 f", vierkant::bcn::compress of each of the {NTEX} textures of the batch",
                                      "blocks": total, "mismatched_blocks": bad, "bit_exact": bad == 0}
                elif args.workload in ("c3", "c5") and base == wl["base"]:
                    import bench_strong
                    ctx._check(ctx.lib.vkt_bcn_cuda_compress(ctx.handle, capi.MODE_BC7, host_src[0].data_ptr(), dims[0][0], dims[0][1], 4, 1,
                                                             C.byref(params), out_ptrs))
                    rep = bench_strong.reference_sample(args.workload, bench_strong.STRONG[args.workload], host_src[0].numpy(),
                                                        [o.numpy() for o in host_out], dims, len(os.sched_getaffinity(0)))
                    line["parity"] = {k: rep[k] for k in ("against", "blocks", "mismatched_blocks", "bit_exact", "sample")}
                    line["cpu_baseline"] = {**rep["cpu_blocks_only"], "kind": "reference" if rep["against"].startswith("unmodified") else "port",
                                            "sample": rep["cpu_blocks_only"]["what"] + " (the blocks of the parity sample; vierkant::bcn::compress() takes no parameters)"}
            except Exception as e:  # noqa: BLE001
                line["parity"] = {"error": repr(e)}
        print(json.dumps(line))
    ctx.close()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
