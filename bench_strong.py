"""bench_strong.py -- the north_star's strong-scaling arm of bench.py: ONE mip chain sharded over the ranks.

BASELINE.json configs[2] (C3: 8192^2 with alpha gradients, max partitions) and configs[4] (C5: 16384^2, uber level 4, all
partitions) are single textures whose block rows and mip levels shard across the GPUs of one box.  The reference shards a
level into batches of four block rows for its thread pool (src/texture_block_compression.cpp:107-139); here rank r of N
(one process per GPU, as torchrun starts them) encodes block rows [rows * r / N, rows * (r + 1) / N) of every level that is
large enough to slice through vkt_bcn_cuda_compress_shard_begin / _end (halo rows recomputed, not exchanged), rank 0
finishes the small tail levels, and every rank's copy engine writes its blocks at their final position of one shared,
page-locked result buffer (vierkant_b200/hostshare.py).  No collective touches the data path; torch.distributed only
provides the start-up barrier and the max-over-ranks reduction of the timings.

Per config the arm reports
  value   : source image and level block buffers resident in HBM (every rank holds a device copy of the source)
  e2e     : source in the shared pinned host buffer, blocks gathered into the shared pinned host level arrays
  n1      : the same two figures for ONE GPU encoding the whole chain (vkt_bcn_cuda_compress on rank 0, others idle),
            measured in the same run, and efficiency_vs_n1 = value / (N * n1.value)
  parity  : gathered blocks == the one-GPU result (all blocks), and a sample checked against the unmodified reference
            (oracle/_ref): C3 the whole chain, C5 a 2048-row slab of level 0 across the middle rank boundary, rows around
            every 1/8 boundary of levels 0..2, and levels >= 3 complete.
"""
from __future__ import annotations

import ctypes as C
import os
import time

import numpy as np

from vierkant_b200 import capi, hostshare, synth

STRONG = {
    "c5": dict(base=16384, kind=0, params=dict(uber_level=4, max_partitions=64, mode17_partition_estimation_filterbank=0),
               name="16384x16384 RGBA8 opaque (synthetic kind 0) + full mip chain (13 levels), uber level 4, 64 partitions, filterbank off",
               max_steps=5),
    "c3": dict(base=8192, kind=1, params=dict(max_partitions=64, mode17_partition_estimation_filterbank=0),
               name="8192x8192 RGBA8 with alpha gradients (synthetic kind 1) + full mip chain (12 levels), 64 partitions, filterbank off",
               max_steps=10),
}


def _align(v: int, a: int = 4096) -> int:
    return (v + a - 1) // a * a


class SharedChain:
    """The shared host buffers of one sharded chain: source image, every level's block array, hand-over rows, barrier flags."""

    def __init__(self, tag: str, width: int, height: int, world: int, rank: int, barrier_fn):
        self.w, self.h, self.world, self.rank = width, height, world, rank
        self.plan = capi.compress_plan(width, height, True)
        self.sp = capi.shard_plan(width, height, True, world)
        self.L = int(self.plan.num_levels)
        self.level_blocks = [int(self.plan.level_num_blocks[l]) for l in range(self.L)]
        self.level_dims = [(int(self.plan.level_width[l]), int(self.plan.level_height[l])) for l in range(self.L)]
        self.out_off, total = [], 0
        for n in self.level_blocks:
            self.out_off.append(total)
            total += _align(n * 16)
        sizes = {"src": width * height * 4, "out": total, "hand": max(int(self.sp.handover_bytes), 4096), "flag": 64 * world}
        self.bufs = {}
        for key, n in sizes.items():
            if rank == 0:
                self.bufs[key] = hostshare.SharedBuffer(f"{tag}_{key}", n, True)
        barrier_fn()
        for key, n in sizes.items():
            if rank != 0:
                self.bufs[key] = hostshare.SharedBuffer(f"{tag}_{key}", n, False)
        self.src = self.bufs["src"].array.reshape(height, width, 4)
        self.levels = [self.bufs["out"].array[o:o + n * 16].reshape(n, 16) for o, n in zip(self.out_off, self.level_blocks)]
        self.handover = self.bufs["hand"].array
        self.barrier = hostshare.FlagBarrier(self.bufs["flag"], rank, world)
        self.barrier.reset()

    def close(self):
        self.src = self.levels = self.handover = self.barrier = None
        for b in self.bufs.values():
            b.close()


def reference_sample(cfg_name: str, cfg: dict, src: np.ndarray, got_levels: list[np.ndarray], level_dims, threads: int) -> dict:
    """Blocks of the gathered result against the unmodified reference (oracle/_ref; the C port where it is absent):
    reference pixels of every level by the reference's own resize (row bands in parallel, byte-identical to its whole-image
    call: oracle/slab.py), reference blocks by its bc7enc_compress_block with the config's parameters."""
    from oracle import pyoracle, slab
    if pyoracle.RefOracle.available():
        oracle, against = pyoracle.RefOracle(), "unmodified reference (oracle/_ref)"
    else:
        pyoracle.build("port")
        oracle, against = pyoracle.PortOracle(), "C port of the reference (oracle/*.c)"
    params = pyoracle.default_params(**cfg["params"])
    t0 = time.perf_counter()
    px = slab.chain_levels(oracle, src, len(level_dims), threads)
    t_filter = time.perf_counter() - t0
    regions = []  # (level, first block row, end block row)
    if cfg_name == "c5":
        h0 = level_dims[0][1]
        regions.append((0, (h0 // 2 - 1024) // 4, (h0 // 2 + 1024) // 4))
        for l in range(3):
            rows = level_dims[l][1] // 4
            for k in range(1, 8):
                if l == 0 and k == 4:
                    continue  # inside the slab
                regions.append((l, rows * k // 8 - 2, rows * k // 8 + 2))
        regions += [(l, 0, level_dims[l][1] // 4) for l in range(3, len(level_dims))]
    else:
        regions = [(l, 0, level_dims[l][1] // 4) for l in range(len(level_dims))]
    checked = bad = 0
    enc_s, enc_pix = 0.0, 0
    per_level = {}
    for l, r0, r1 in regions:
        bx = level_dims[l][0] // 4
        tiles = synth.to_blocks(np.ascontiguousarray(px[l][4 * r0:4 * r1]))
        t0 = time.perf_counter()
        want = oracle.encode_blocks(tiles, params, threads=threads)
        enc_s += time.perf_counter() - t0
        enc_pix += tiles.shape[0] * 16
        got = got_levels[l][r0 * bx:r1 * bx]
        nbad = int((got != want).any(axis=1).sum())
        checked += want.shape[0]
        bad += nbad
        per_level[l] = per_level.get(l, 0) + want.shape[0]
    return {"against": against, "blocks": checked, "mismatched_blocks": bad, "bit_exact": bad == 0,
            "blocks_per_level": per_level, "sample": ("whole chain" if cfg_name != "c5" else
                                                      "2048-row slab of level 0 across the middle rank boundary, 4 block rows around every 1/8 boundary of levels 0-2, levels >= 3 complete"),
            "reference_filter_s": t_filter,
            "cpu_blocks_only": {"value": enc_pix / enc_s * 1e-6 if enc_s else None, "unit": "Mpixel/s", "cores": threads,
                                "what": "bc7enc_compress_block over the sampled blocks (pre-filtered), row partition over host threads"}}


def run_strong(cfg_name: str, ctx: capi.BcnContext, rank: int, world: int, dev, args, dist, base: int | None = None, check: bool = True,
               clock_sampler=None):
    """One strong-scaling config on `world` ranks.  Returns the report dict on rank 0, None elsewhere."""
    import torch
    cfg = STRONG[cfg_name]
    W = H = base or cfg["base"]
    params = capi.default_params(**cfg["params"])
    steps = max(1, min(args.steps, cfg["max_steps"]))
    warmup = max(3, min(args.warmup, 3))
    tag = f"vkt_strong_{os.environ.get('MASTER_PORT', 'solo')}_{os.getuid()}_{cfg_name}"

    def job_barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    sc = SharedChain(tag, W, H, world, rank, job_barrier)
    try:
        ncpu = len(os.sched_getaffinity(0))
        r0, r1 = H * rank // world, H * (rank + 1) // world
        synth.fill_shared(sc.bufs["src"].path, 0, W, H, cfg["kind"], 0xB200 + cfg["kind"], (r0, r1), procs=max(1, min(16, ncpu // world)))
        job_barrier()
        for key in ("src", "out", "hand"):
            ctx.host_register(sc.bufs[key].array)
        d_src = torch.from_numpy(sc.src).to(dev)  # whole image on every rank's device (the resident arm reads its rows in place)
        d_levels = [torch.zeros((n, 16), dtype=torch.uint8, device=dev) for n in sc.level_blocks]
        torch.cuda.synchronize()
        L = sc.L
        host_ptrs = (C.c_void_p * L)(*[l.ctypes.data for l in sc.levels])
        dev_ptrs = (C.c_void_p * L)(*[t.data_ptr() for t in d_levels])
        phase = [0]

        def step(src, ptrs):
            # begin: queue this rank's slices;  barrier (flags only): every rank's rows of the last sliced level are in the
            # hand-over buffer;  end: rank 0 queues the tail levels, everybody waits for its own work;  lockstep per chain
            ctx.compress_shard_begin(capi.MODE_BC7, src, W, H, 4, True, params, rank, world, ptrs, sc.handover)
            phase[0] += 1
            sc.barrier.arrive(phase[0])
            if rank == 0:
                sc.barrier.wait_all(phase[0])
            ctx.compress_shard_end(capi.MODE_BC7, src, W, H, 4, True, params, rank, world, ptrs, sc.handover)
            phase[0] += 1
            sc.barrier.arrive(phase[0])
            sc.barrier.wait_all(phase[0])

        def timed(fn):
            for _ in range(warmup):
                fn()
            job_barrier()
            s0 = ctx.stats()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            t0 = time.perf_counter()
            for _ in range(steps):
                fn()
            e1.record()  # every call above has waited for its own GPU work: the pair brackets all of it
            torch.cuda.synchronize()
            wall = (time.perf_counter() - t0) * 1e3
            s1 = ctx.stats()
            job_barrier()
            t = torch.tensor([max(e0.elapsed_time(e1), wall)], dtype=torch.float64, device=dev)
            if world > 1:
                dist.all_reduce(t, op=dist.ReduceOp.MAX)
            return float(t[0]) / steps, {k: (s1[k] - s0[k]) // steps for k in s0}

        if clock_sampler is not None:  # rank 0: SM clocks / throttle reasons over both timed regions (seconds long for C5)
            clock_sampler.start()
        ms_res, st_res = timed(lambda: step(d_src, dev_ptrs))
        ms_e2e, st_e2e = timed(lambda: step(sc.src, host_ptrs))
        clocks = clock_sampler.stop() if clock_sampler is not None else None
        # per-rank traffic of the gathered run, summed over the job
        tr = torch.tensor([st_e2e["h2d_bytes"], st_e2e["d2h_bytes"], st_res["kernel_launches"]], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(tr, op=dist.ReduceOp.SUM)

        # ---- one GPU alone (rank 0): the N = 1 figure of the same run, and the blocks the gathered result must equal
        n1 = None
        n1_out = None
        if rank == 0:
            n1_out = [torch.empty((n, 16), dtype=torch.uint8).pin_memory() for n in sc.level_blocks]
            n1_ptrs = (C.c_void_p * L)(*[t.data_ptr() for t in n1_out])
            n1_dev = [torch.empty((n, 16), dtype=torch.uint8, device=dev) for n in sc.level_blocks]
            n1_dev_ptrs = (C.c_void_p * L)(*[t.data_ptr() for t in n1_dev])

            def alone(src_ptr, ptrs):
                ctx._check(ctx.lib.vkt_bcn_cuda_compress(ctx.handle, capi.MODE_BC7, src_ptr, W, H, 4, 1, C.byref(params), ptrs))

            def timed_alone(fn):
                for _ in range(warmup if world > 1 else 1):
                    fn()
                torch.cuda.synchronize()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                t0 = time.perf_counter()
                for _ in range(steps):
                    fn()
                e1.record()
                torch.cuda.synchronize()
                return max(e0.elapsed_time(e1), (time.perf_counter() - t0) * 1e3) / steps

            if world > 1:
                n1 = {"ms_resident": timed_alone(lambda: alone(d_src.data_ptr(), n1_dev_ptrs)),
                      "ms_e2e": timed_alone(lambda: alone(sc.src.ctypes.data, n1_ptrs))}
            else:
                alone(sc.src.ctypes.data, n1_ptrs)  # N = 1: the sharded call IS the one-GPU call; keep the cross-check
        job_barrier()

        report = None
        if rank == 0:
            npix = sum(w * h for w, h in sc.level_dims)
            nblocks = sum(sc.level_blocks)
            value, e2e = npix / ms_res * 1e-3, npix / ms_e2e * 1e-3
            same = all(np.array_equal(a, b.numpy()) for a, b in zip(sc.levels, n1_out))
            # the resident arm leaves each rank's rows in that rank's HBM: compare rank 0's own rows with the gathered result
            dev_same = True
            for l, (a, b) in enumerate(zip(sc.levels, d_levels)):
                b0, b1 = capi.shard_rows(W, H, True, 0, world, l)
                bx = sc.level_dims[l][0] // 4
                dev_same = dev_same and np.array_equal(a[b0 * bx:b1 * bx], b[b0 * bx:b1 * bx].cpu().numpy())
            report = {
                "workload": cfg["name"] if W == cfg["base"] else cfg["name"].replace(f"{cfg['base']}x{cfg['base']}", f"{W}x{W}"),
                "scaling": "strong", "n_gpus": world, "steps": steps, "warmup": warmup, "unit": "Mpixel/s",
                "value": value, "ms_per_step": ms_res, "e2e": e2e, "e2e_ms_per_step": ms_e2e,
                "blocks": nblocks, "mpixel": npix * 1e-6, "levels": L, "sliced_levels": int(sc.sp.sliced_levels), "workers": int(sc.sp.workers),
                "h2d_bytes_per_step": int(tr[0]), "d2h_bytes_per_step": int(tr[1]), "gpu_launches_per_step": int(tr[2]),
                "partitioning": (f"block rows of levels 0..{int(sc.sp.sliced_levels) - 1} split evenly over {int(sc.sp.workers)} ranks (halo rows recomputed), "
                                 f"levels {int(sc.sp.sliced_levels)}..{L - 1} on rank 0 from a {int(sc.sp.handover_bytes)} B host hand-over; no collective"
                                 if int(sc.sp.workers) > 1 else "one rank encodes the whole chain"),
                "timing": "CUDA events around K lockstep chains (every call waits for its own GPU work), max with the host clock, max over ranks",
                "clocks": clocks,
            }
            try:  # ALU roofline of the whole job: gcov-pinned ops per pixel x pixel rate / (N x issue peak)
                import json
                with open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "profiles", "gcov_ops.json")) as f:
                    opp = float(json.load(f)["configs"][cfg_name]["ops_per_pixel"])
                props = torch.cuda.get_device_properties(dev)
                peak = props.multi_processor_count * 4 * 32 * 1965.0e6
                report["roofline"] = {"bound": "alu", "ops_per_pixel": opp, "ops_source": "profiles/gcov_ops.json",
                                      "achieved": value * 1e6 * opp * 1e-12, "peak": world * peak * 1e-12, "unit": "Tlane-op/s",
                                      "frac": value * 1e6 * opp / (world * peak),
                                      "peak_source": f"{world} x {props.multi_processor_count} SMs x 4 x 32 lanes x 1965 MHz"}
            except Exception:  # noqa: BLE001
                pass
            if n1 is not None:
                n1v, n1e = npix / n1["ms_resident"] * 1e-3, npix / n1["ms_e2e"] * 1e-3
                report["n1"] = {"value": n1v, "e2e": n1e, "ms_per_step": n1["ms_resident"], "e2e_ms_per_step": n1["ms_e2e"],
                                "what": "vkt_bcn_cuda_compress of the whole chain on rank 0's GPU alone, same run"}
                report["efficiency_vs_n1"] = value / (world * n1v)
                report["e2e_efficiency_vs_n1"] = e2e / (world * n1e)
            else:
                report["efficiency_vs_n1"] = 1.0
            parity = {"gathered_equals_one_gpu": bool(same), "resident_rows_of_rank0_equal_gathered": bool(dev_same), "blocks_compared": nblocks}
            if check:
                try:
                    parity["reference"] = reference_sample(cfg_name, cfg, sc.src, sc.levels, sc.level_dims, ncpu)
                except Exception as e:  # noqa: BLE001 -- the check must never take the bench line down
                    parity["reference"] = {"error": repr(e)}
            report["parity"] = parity
        job_barrier()
        for key in ("src", "out", "hand"):
            ctx.host_unregister(sc.bufs[key].array)
        return report
    finally:
        sc.close()
