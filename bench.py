#!/usr/bin/env python
"""Benchmark of the atmosphere-LUT build (BASELINE.json: "atmosphere-lut build time (s) and ray-scatter
texels/s at 1/2/4/8 B200").

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]

One *step* = one full `generate-atmosphere-luts` build (atmosphere_lut.clj:43-105) at shipped resolution:
transmittance, surface-radiance-base, first-order ray scatter (Rayleigh + Mie strength), 5 iterations of
point-scatter / surface-radiance / ray-scatter / accumulation, and the final resampling into file layout.
Inputs are parameters only (deterministic, no data files), so "inputs resident in HBM" means the builder
with its quadrature tables and table buffers exists before the timed region.

metric (`value`) = 4-D texels x (1 + iterations) scattering orders / build time: the build rate, strong scaling.
The per-kernel rates SURVEY.md 8(d) names are printed beside it: `ray_scatter_pass_texels_per_s` = N4 / (one
ray-scatter pass) and `first_order_texels_per_s` = N4 / (the first-order pass).

With N > 1 (torchrun, one process per GPU) the SAME build is sharded over the ranks.  Default exchange `p2p`: whole
height rows are dealt to the ranks, the kernels store every finished texel into all GPUs' tables over NVLink (CUDA
IPC) and a flag barrier closes each table -- no collective moves table data.  `--mode nccl`: contiguous slabs and
one NCCL all-gather per table (12 per build); its build time is also measured and printed as `nccl_mode` in the
default run.  Outside the timed region rank 0 builds the same tables on one GPU and compares the sharded result byte
for byte (`sharded_identical`, both modes); a difference makes the run fail.

`--impl reference` times the reference algorithm on the host cores: the double-precision CPU oracle (oracle/, a
restatement of the Clojure code -- there is no JVM on the box) on a bounded random sample of texels of every stage,
extrapolated per texel to the full build, on ALL host cores (also under torchrun, which exports OMP_NUM_THREADS=1).
"""
import argparse
import ctypes
import json
import os
import statistics
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "ray_scatter_texels_per_s"
UNIT = "texels/s"
WORKLOAD = "full atmosphere-lut build, shipped resolution (4-D 32x127x32x8, T 64x255, E 16x63, ray-steps 100, " \
           "sphere-steps 15, 5 iterations), Earth defaults"
STRESS_WORKLOAD = "stress: 4-D 64x253x64x16 (2x shipped per axis), T 64x255, E 16x63, ray-steps 100, sphere-steps 15, " \
                  "10 iterations, Earth defaults"


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=30)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--cpu-sample", type=int, default=24576, help="texels per stage in the CPU baseline sample")
    ap.add_argument("--mode", default="p2p", choices=["p2p", "nccl"],
                    help="multi-GPU exchange: p2p = kernels store into all GPUs' tables over NVLink + flag barrier; "
                         "nccl = contiguous slabs + one NCCL all-gather per table")
    ap.add_argument("--workload", default="shipped", choices=["shipped", "stress"],
                    help="shipped = BASELINE.json configs[2] (the headline); stress = configs[4]: 4-D table at 2x "
                         "resolution per axis (64x253x64x16), 10 iterations")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-extras", action="store_true",
                    help="skip the stress build, the other exchange mode and the identity checks (profiling runs)")
    return ap.parse_args()


# ------------------------------------------------------------------ clocks

class ClockSampler:
    """Samples SM clock and throttle reasons through NVML while the timed region runs."""

    def __init__(self, index):
        self.samples, self.reasons = [], set()
        self.max_mhz = None
        self._stop = threading.Event()
        self._thread = None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
        except Exception:
            self.nv = None

    def _loop(self):
        nv = self.nv
        names = {"hw_slowdown": nv.nvmlClocksThrottleReasonHwSlowdown,
                 "hw_thermal_slowdown": nv.nvmlClocksThrottleReasonHwThermalSlowdown,
                 "sw_thermal_slowdown": nv.nvmlClocksThrottleReasonSwThermalSlowdown,
                 "sw_power_cap": nv.nvmlClocksThrottleReasonSwPowerCap}
        while not self._stop.is_set():
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                mask = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for k, bit in names.items():
                    if mask & bit:
                        self.reasons.add(k)
            except Exception:
                pass
            self._stop.wait(0.01)

    def start(self):
        if self.nv:
            self._thread = threading.Thread(target=self._loop, daemon=True)
            self._thread.start()

    def stop(self):
        if self._thread:
            self._stop.set()
            self._thread.join()
        return {"sm_mhz": statistics.median(self.samples) if self.samples else None, "sm_max_mhz": self.max_mhz,
                "reasons": sorted(self.reasons), "samples": len(self.samples)}


# ------------------------------------------------------------------ CPU baseline (oracle; checker only)

def cpu_baseline(sample, threads=None):
    """Times the CPU oracle on `sample` random texels of every stage and extrapolates to the full build.
    All host cores by default: torchrun exports OMP_NUM_THREADS=1, so the thread count is set explicitly."""
    import numpy as np
    from oracle import oracle as orc
    orc.set_num_threads(threads or os.cpu_count() or 1)
    cores = orc.num_threads()
    pl = orc.planet(**orc.EARTH)
    mie, ray = orc.scatter(**orc.MIE), orc.scatter(**orc.RAYLEIGH)
    shape4, shape_t, shape_e = (32, 127, 32, 8), (64, 255), (16, 63)
    cfg = orc.config(shape4, shape_t, shape_e, 100, 15)
    n4, nt, ne = int(np.prod(shape4)), int(np.prod(shape_t)), int(np.prod(shape_e))
    iterations = 5
    rng = np.random.default_rng(1)
    idx4 = np.sort(rng.choice(n4, size=min(sample, n4), replace=False))
    idxe = np.sort(rng.choice(ne, size=min(max(16, sample // 64), ne), replace=False))
    # lookup cost does not depend on table values: synthetic previous-order tables
    tab4 = rng.random(shape4 + (3,)) * 1e-3
    tabe = rng.random(shape_e + (3,)) * 1e-3
    stages = {}

    def timed(name, reps, total, n, fn):
        t0 = time.perf_counter()
        fn()
        dt = time.perf_counter() - t0
        stages[name] = {"sample_s": dt, "sample_texels": n, "full_s": dt / n * total * reps}

    orc.counters_reset()
    timed("transmittance", 1, nt, nt, lambda: orc.table_transmittance(pl, [mie, ray], cfg))
    timed("surface_radiance_base", 1, ne, ne, lambda: orc.table_surface_radiance_base(pl, [mie, ray], cfg))
    # the reference integrates Rayleigh and Mie strength in two separate passes (atmosphere_lut.clj:71-72)
    timed("first_order_rayleigh", 1, n4, len(idx4), lambda: orc.table_first_order(pl, [mie, ray], cfg, ray, 0, idx4))
    timed("first_order_mie_strength", 1, n4, len(idx4), lambda: orc.table_first_order(pl, [mie, ray], cfg, mie, 1, idx4))
    ds1 = orc.SSourceSpec(tab4, tab4, mie)
    ds = orc.SSourceSpec(tab4)
    timed("point_scatter_iter1", 1, n4, len(idx4), lambda: orc.table_point_scatter(pl, [mie, ray], cfg, ds1, tabe, idx4))
    timed("point_scatter_iter2+", iterations - 1, n4, len(idx4),
          lambda: orc.table_point_scatter(pl, [mie, ray], cfg, ds, tabe, idx4))
    timed("surface_radiance_iter1", 1, ne, len(idxe), lambda: orc.table_surface_radiance(pl, cfg, ds1, idxe))
    timed("surface_radiance_iter2+", iterations - 1, ne, len(idxe), lambda: orc.table_surface_radiance(pl, cfg, ds, idxe))
    timed("ray_scatter", iterations, n4, len(idx4), lambda: orc.table_ray_scatter(pl, [mie, ray], cfg, tab4, idx4))
    timed("accumulate+final", iterations + 2, n4, len(idx4),
          lambda: orc.table_resample_sum_4d(pl, cfg, [tab4, tab4], idx4))
    counters = orc.counters_get()
    full_s = sum(s["full_s"] for s in stages.values())
    sample_s = sum(s["sample_s"] for s in stages.values())
    texels = n4 * (1 + iterations)
    return {"value": texels / full_s, "unit": UNIT, "cores": cores, "kind": "port",
            "sample": "%d random 4-D texels per stage (%d surface-radiance texels), full 2-D tables; previous-order "
                      "tables synthetic; extrapolated per texel to the full build" % (len(idx4), len(idxe)),
            "sample_cpu_s": sample_s, "build_time_s_extrapolated": full_s, "extrapolated": True,
            "esamples_in_sample": counters["esamples"],
            "note": "CPU restatement in C (double, OpenMP), not JVM Clojure: a lower bound on the reference's time; "
                    "the same oracle ran the WHOLE shipped build in 755 s on 8 cores (tests/golden/make_shipped_golden.py)",
            "stages": {k: round(v["full_s"], 3) for k, v in stages.items()}}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    t0 = time.perf_counter()
    per_step = []
    base = None
    for i in range(args.warmup + args.steps):
        base = cpu_baseline(args.cpu_sample)
        if i >= args.warmup:
            per_step.append(base["build_time_s_extrapolated"])
        if time.perf_counter() - t0 > 240:   # bounded: the whole arm ends within a few minutes
            break
    if not per_step:
        per_step = [base["build_time_s_extrapolated"]]
    build_s = statistics.mean(per_step)
    texels = 32 * 127 * 32 * 8 * 6
    value = texels / build_s
    base["value"] = value
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
            "steps": len(per_step), "warmup": args.warmup, "ms_per_step": build_s * 1e3, "higher_is_better": True,
            "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "build_time_s": build_s,
            "config": {"workload": WORKLOAD, "note": "each step = bounded sample extrapolated to the full build"},
            "cpu_baseline": base,
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------ B200 arm

def stage_sums(stage_times):
    """[(name, ms)] -> {key: ms} with the per-iteration stages summed ("iter3_ray_scatter" -> "ray_scatter")."""
    out = {}
    for name, ms in stage_times:
        key = name.split("_", 1)[1] if name.startswith("iter") else name
        out[key] = out.get(key, 0.0) + ms
    return out


class Runner:
    """One configuration on this rank's GPU: device-timed steps of builder.run()."""

    def __init__(self, torch, dist, world, flush):
        self.torch, self.dist, self.world, self.flush = torch, dist, world, flush

    def barrier(self):
        if self.world > 1:
            self.dist.barrier()
        self.torch.cuda.synchronize()

    def max_over_ranks(self, value):
        if self.world == 1:
            return float(value)
        t = self.torch.tensor([float(value)], dtype=self.torch.float64, device="cuda")
        self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
        return float(t.item())

    def time_steps(self, builder, steps, warmup):
        """(ms per step, max over ranks): CUDA events on the builder's stream around every build, a 512 MiB
        write between steps so that no table survives in L2."""
        torch = self.torch
        stream = torch.cuda.ExternalStream(builder.stream)
        for _ in range(max(warmup, 3)):
            builder.run()
        self.barrier()
        starts = [torch.cuda.Event(enable_timing=True) for _ in range(steps)]
        ends = [torch.cuda.Event(enable_timing=True) for _ in range(steps)]
        self.barrier()
        t0 = time.perf_counter()
        for i in range(steps):
            with torch.cuda.stream(stream):
                self.flush.zero_()                      # evict the previous step's tables from L2
            starts[i].record(stream)
            builder.run()
            ends[i].record(stream)
        self.barrier()
        wall = time.perf_counter() - t0
        builder.sync()                                  # surfaces a barrier timeout as an error
        total = torch.tensor([sum(s.elapsed_time(e) for s, e in zip(starts, ends))], dtype=torch.float64, device="cuda")
        if self.world > 1:
            self.dist.all_reduce(total, op=self.dist.ReduceOp.MAX)
        return float(total.item()) / steps, wall


def cubemap_bench(rank, world, runner, with_cpu, out_level=4):
    """SURVEY.md section 8(f) row 4, `clj -T:build cube-maps` (build.clj:294-310): every tile of one output level of the
    cube-map pyramid with the shipped constants (675-pixel map tiles, 65 / 129-pixel cube tiles) from synthetic rasters
    resident in device memory.  Tiles are independent: rank r takes every world-th tile, no exchange (weak scaling is
    the natural mode; here the level is a fixed job, so this is strong scaling of the 1536-tile level)."""
    import numpy as np
    from sfsim_b200 import cubemap
    from sfsim_b200.synthetic import synthetic_world
    in_level = out_level - 3                           # build.clj:300-310
    cfg = cubemap.make_config(in_level, out_level)
    ls, lc, lw = max(0, min(4, in_level)), max(0, min(5, in_level + 1)), max(0, min(4, in_level + 1))
    elev, day, night = synthetic_world(675, sorted({ls, lw}), [lc], seed=1)
    w = cubemap.World(675)
    for level, a in elev.items():
        w.set_elevation(level, a)
    w.set_color(False, lc, day[lc])
    w.set_color(True, lc, night[lc])
    tiles = cubemap.tile_shard(out_level, rank, world)
    total_tiles = 6 * 4 ** out_level
    pixels = total_tiles * cfg.color_tilesize ** 2
    for _ in range(3):
        w.time_cube_map_tiles(cfg, tiles)
    ms = []
    for _ in range(5):
        runner.barrier()
        ms.append(runner.max_over_ranks(w.time_cube_map_tiles(cfg, tiles)))
    dev_ms = float(np.median(ms))
    # end to end: the streamed level call -- tile list in, every tile's five arrays delivered to a host callback in
    # page-locked memory (691 MB per 1536 tiles cross PCIe), copies overlapped with the kernels of the next batch
    per_tile = 2 * cfg.color_tilesize ** 2 * 4 + cfg.color_tilesize * ((cfg.color_tilesize + 3) & ~3) + \
        cfg.surface_tilesize ** 2 * 12 + cfg.color_tilesize ** 2 * 15
    d2h = per_tile * len(tiles)
    w.time_cube_map_level(cfg, rank, world)
    runner.barrier()
    t0 = time.perf_counter()
    delivered = 0
    for _ in range(3):
        delivered = w.time_cube_map_level(cfg, rank, world)[1]
    runner.barrier()
    e2e_s = (time.perf_counter() - t0) / 3
    assert delivered == len(tiles)
    # the same call without the float normals (a host that encodes the PNG from normal_bytes): 44 % fewer bytes on the bus
    lean = ("day", "night", "water", "surface", "normal_bytes")
    w.time_cube_map_level(cfg, rank, world, outputs=lean)
    runner.barrier()
    t0 = time.perf_counter()
    for _ in range(3):
        w.time_cube_map_level(cfg, rank, world, outputs=lean)
    runner.barrier()
    lean_s = (time.perf_counter() - t0) / 3
    res = {"workload": "cube-map tiles, output level %d (6 x %d x %d tiles, in-level %d), map tiles 675 px, cube tiles 65 / 129 "
                       "px, synthetic rasters (elevation levels %s, colour level %d)" % (
                           out_level, 1 << out_level, 1 << out_level, in_level, sorted({ls, lw}), lc),
           "tiles": total_tiles, "ms_per_level": dev_ms, "colour_pixels_per_s": pixels / (dev_ms * 1e-3),
           "tiles_per_s": total_tiles / (dev_ms * 1e-3),
           "e2e": {"tiles_per_s": total_tiles / e2e_s, "seconds_per_level": e2e_s, "d2h_bytes_per_level_this_rank": d2h,
                   "h2d_bytes_per_level_this_rank": int(tiles.nbytes), "api": "sfsim_cubemap_level (batches of 256 tiles, counting "
                   "callback)", "host_buffers": "library-owned page-locked staging, two sets",
                   "without_float_normals": {"tiles_per_s": total_tiles / lean_s, "seconds_per_level": lean_s,
                                             "d2h_bytes_per_level_this_rank": (per_tile - cfg.color_tilesize ** 2 * 12) * len(tiles)}},
           "bound": "FP64 pipe / issue slots (ncu: profiles/r2/); DRAM traffic 1.2 GB per level = the outputs"}
    if with_cpu and rank == 0:
        from oracle import cubemap as ocm             # the cpu_baseline leg: the C restatement on the host cores
        ow = ocm.OracleWorld(675, elev, day, night)
        sample = [tuple(t) for t in tiles[:: max(1, len(tiles) // 32)][:32]]
        ow.make_cube_map_tile(*sample[0][:1], in_level, out_level, *sample[0][1:])      # warm-up: threads, page faults
        t0 = time.perf_counter()
        for face, b, a in sample:
            ow.make_cube_map_tile(face, in_level, out_level, b, a)
        cpu_s = (time.perf_counter() - t0) / len(sample)
        res["cpu_baseline"] = {"value": 1.0 / cpu_s, "unit": "tiles/s", "cores": os.cpu_count(), "kind": "port",
                               "sample": "%d tiles of the same level with the C restatement (OpenMP over pixel rows)" % len(sample)}
    w.close()
    return res


def run_b200(args):
    import numpy as np
    import torch
    import torch.distributed as dist
    from sfsim_b200 import _lib, atmosphere_lut

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if world != args.gpus and world > 1:
        raise SystemExit("--gpus %d does not match WORLD_SIZE %d" % (args.gpus, world))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: there is no CPU fallback for the product path")
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    lib = _lib.load()
    _lib.check(lib.atmlut_init(local_rank))
    shipped_cfg = _lib.default_config()
    stress_cfg = _lib.make_config(ray_scatter_shape=(64, 253, 64, 16), iterations=10)
    cfg, workload = (shipped_cfg, WORKLOAD) if args.workload == "shipped" else (stress_cfg, STRESS_WORKLOAD)
    flush = torch.empty(512 << 20, dtype=torch.uint8, device="cuda")   # > 126 MB L2
    runner = Runner(torch, dist, world, flush)

    def texels_of(c):
        return c.height_size * c.elevation_size * c.light_elevation_size * c.heading_size

    # ---------------- the headline: device-timed builds
    builder = atmosphere_lut.AtmosphereLutBuilder(cfg=cfg, rank=rank, world=world, mode=args.mode)
    sampler = ClockSampler(local_rank)
    runner.time_steps(builder, 2, args.warmup)          # warm-up (graph capture, clocks)
    sampler.start()
    ms_per_step, t_wall = runner.time_steps(builder, args.steps, 0)
    clocks = sampler.stop()
    work = builder.work()

    # ---------------- per-stage times: the same build with event records between the stages
    builder.run_timed()
    builder.sync()
    runner.barrier()
    builder.run_timed()
    builder.sync()
    stage_times = builder.stage_times()
    my_stages = stage_sums(stage_times)
    all_stages = [my_stages]
    if world > 1:
        all_stages = [None] * world
        dist.all_gather_object(all_stages, my_stages)

    n4 = texels_of(cfg)
    texels = n4 * (1 + cfg.iterations)
    d2h = sum(int(np.prod(s)) * 4 for s in atmosphere_lut.output_shapes(cfg))
    h2d = ctypes.sizeof(_lib.Planet) + 2 * ctypes.sizeof(_lib.Scatter) + ctypes.sizeof(_lib.Config)

    # ---------------- end to end through the public call: host parameters in, host tables out
    def time_e2e(fn, n):
        runner.barrier()
        t0 = time.perf_counter()
        for _ in range(n):
            fn()
        runner.barrier()
        return (time.perf_counter() - t0) / n

    n_e2e = max(5, args.steps // 3)
    e2e_pageable = None
    if world == 1:
        out = atmosphere_lut.allocate_outputs(cfg, pinned=True)
        for _ in range(3):
            atmosphere_lut.generate_tables(cfg=cfg, out=out)
        e2e_s = time_e2e(lambda: atmosphere_lut.generate_tables(cfg=cfg, out=out), n_e2e)
        # what a host without pinned buffers gets (the JVM's FFM arenas): pageable destinations, staged by the library
        out_p = atmosphere_lut.allocate_outputs(cfg, pinned=False)
        for o in out_p:
            o.fill(0)                                   # touch the pages once, like a reused arena
        atmosphere_lut.generate_tables(cfg=cfg, out=out_p)
        e2e_pageable_s = time_e2e(lambda: atmosphere_lut.generate_tables(cfg=cfg, out=out_p), n_e2e)
        e2e_pageable = {"value": texels / e2e_pageable_s, "unit": UNIT, "build_time_s": e2e_pageable_s,
                        "note": "atmlut_generate into pageable host memory (malloc / JVM arena): the download is "
                                "pipelined through two library-owned pinned buffers"}
        api = "atmlut_generate"
    else:
        out = atmosphere_lut.allocate_outputs(cfg, pinned=True)

        def sharded_e2e():
            builder.run()
            if rank == 0:
                builder.download(out)
            else:
                builder.sync()
        sharded_e2e()
        e2e_s = time_e2e(sharded_e2e, n_e2e)
        api = "atmlut_builder_run + atmlut_builder_download on rank 0"
    e2e = {"value": texels / e2e_s, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
           "build_time_s": e2e_s, "api": api, "host_buffers": "pinned"}

    # ---------------- identity of the sharded build, the other exchange mode, the stress configuration
    extras = {}
    identical = None
    if not args.no_extras:
        def single_gpu_tables(c):
            if rank != 0:
                return None
            single = atmosphere_lut.AtmosphereLutBuilder(cfg=c)
            single.run()
            single.sync()
            t = single.download()
            single.close()
            return t

        def same_bytes(b, want):
            flag = torch.zeros(1, device="cuda")
            if rank == 0:
                got = b.download()
                flag[0] = 1.0 if all(np.array_equal(g, w) for g, w in zip(got, want)) else 0.0
            if world > 1:
                dist.broadcast(flag, 0)
            return bool(flag.item() > 0.5)

        if world > 1:
            want = single_gpu_tables(cfg)
            builder.run()
            builder.sync()
            runner.barrier()
            identical = {args.mode: same_bytes(builder, want)}
            other = "nccl" if args.mode == "p2p" else "p2p"
            b2 = atmosphere_lut.AtmosphereLutBuilder(cfg=cfg, rank=rank, world=world, mode=other)
            ms2, _ = runner.time_steps(b2, max(3, args.steps // 2), 3)
            identical[other] = same_bytes(b2, want)
            extras["%s_mode" % other] = {"ms_per_step": ms2, "value": texels / (ms2 * 1e-3), "unit": UNIT,
                                         "note": "the same sharded build with the other exchange"}
            if other == "nccl":
                extras["nccl_mode"]["all_gathers_per_build"] = b2.gathers // (max(3, args.steps // 2) + 3)
            b2.close()
        if args.workload == "shipped":
            # BASELINE.json configs[4] beside the headline
            sb = atmosphere_lut.AtmosphereLutBuilder(cfg=stress_cfg, rank=rank, world=world, mode=args.mode)
            s_ms, _ = runner.time_steps(sb, 3, 2)
            sb.run_timed()
            sb.sync()
            runner.barrier()
            sb.run_timed()
            sb.sync()
            s_stages = stage_sums(sb.stage_times())
            s_n4 = texels_of(stress_cfg)
            extras["stress"] = {"workload": STRESS_WORKLOAD, "ms_per_step": s_ms,
                                "value": s_n4 * (1 + stress_cfg.iterations) / (s_ms * 1e-3), "unit": UNIT,
                                "stage_ms_rank0": {k: round(v, 3) for k, v in s_stages.items()}, "steps": 3}
            if world > 1:
                s_want = single_gpu_tables(stress_cfg)
                sb.run()
                sb.sync()
                runner.barrier()
                extras["stress"]["sharded_identical"] = same_bytes(sb, s_want)
                if identical is not None:
                    identical["stress_" + args.mode] = extras["stress"]["sharded_identical"]
            sb.close()
        if args.workload == "shipped":
            extras["cubemap"] = cubemap_bench(rank, world, runner, with_cpu=not args.no_cpu_baseline and world == 1)

    # ---------------- rooflines
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "profiles", "pipe_peaks_b200.json")))
    except Exception:
        pass
    first_ms = my_stages.get("first_order")
    ray_ms = my_stages.get("ray_scatter")
    point_ms = my_stages.get("point_scatter")
    if rank == 0:
        launches_per_step = int(work["kernel_launches"])
        mufu_peak = peaks.get("mufu_ex2_per_s")
        sm_count = torch.cuda.get_device_properties(local_rank).multi_processor_count
        clock_hz = (clocks.get("sm_mhz") or clocks.get("sm_max_mhz") or 1965.0) * 1e6
        hbm_peak = None
        try:
            hbm_peak = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"]
        except Exception:
            pass
        # DRAM traffic of one first-order launch from the committed `ncu --set full` capture (single GPU only)
        traffic, traffic_source = None, None
        if world == 1 and args.workload == "shipped":
            for name in ("r2/r2_final_ncu_summary.txt", "r1_final_ncu_summary.txt"):
                try:
                    blocks = open(os.path.join(ROOT, "profiles", name)).read().split("\n== ")
                    text = [b for b in blocks if "k_first_order" in b.splitlines()[0]][0]
                    units = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
                    traffic = 0.0
                    for ln in text.splitlines():
                        if "dram__bytes_read.sum" in ln or "dram__bytes_write.sum" in ln:
                            parts = ln.split()
                            traffic += float(parts[1]) * units[parts[2]]
                    traffic_source = "from_profile: profiles/%s (one ncu --set full capture of this kernel, not this run)" % name
                    break
                except Exception:
                    traffic = None
        roofline = None
        if first_ms and work.get("esamples_first_order") and mufu_peak:
            achieved = 2.0 * work["esamples_first_order"] / (first_ms * 1e-3)
            issued = work.get("mufu_ex2_per_esample") or 2.0
            roofline = {"bound": "sfu", "kernel": "k_first_order", "achieved": achieved / 1e9, "peak": mufu_peak / 1e9,
                        "unit": "G exp/s (MUFU.EX2 peak)", "frac": achieved / mufu_peak, "traffic": traffic,
                        "traffic_source": traffic_source,
                        "mufu_ex2_issued_per_esample": issued,
                        "mufu_pipe_frac": achieved * issued / 2.0 / mufu_peak,
                        "hbm": {"algorithmic_bytes": 2 * n4 // world * 16,
                                "achieved_GBs": 2 * n4 / world * 16 / (first_ms * 1e-3) / 1e9, "peak_GBs": hbm_peak,
                                "note": "two float4 output tables per launch; the kernel is SFU/FMA-bound, not HBM-bound"},
                        "peak_source": "measured on this pool's B200 by tools/pipe_peaks.cu (profiles/pipe_peaks_b200.json)",
                        "algorithmic_units": "2 exponentials per overall-extinction sample (one per scatter component, "
                                             "SURVEY.md section 8d); samples counted on the device.  `frac` is that "
                                             "algorithmic rate against the MUFU.EX2 peak; the kernel ISSUES "
                                             "mufu_ex2_issued_per_esample exponentials per sample (Earth's scale heights "
                                             "are 24000 m / 20 and / 3, so every second pair of samples takes both "
                                             "densities from one exponential as t^20 and t^3 on the FMA pipe): "
                                             "`mufu_pipe_frac` is the MUFU pipe's own utilisation by these exponentials",
                        "esamples_per_launch": work["esamples_first_order"], "launch_ms": first_ms}
        # ray-scatter from the dJ table: bound by shared-memory wavefronts.  Algorithmic bytes per 4-D lookup: four
        # RGB corners of the blended tile (48 B read) and the thread's share of the tile (12 B written) = 60 B, i.e. 15
        # wavefronts of 128 B per warp-lookup (the kernel keeps red/green and blue in separate planes, so no padding
        # lane is moved; with padded float4 texels -- the model of round 1 -- it would be 20).
        roofline_k6 = None
        if ray_ms and cfg.iterations:
            lookups = (n4 / world) * 100.0 * cfg.iterations
            per_s = lookups / (ray_ms * 1e-3)
            peak = sm_count * clock_hz * 32.0 / 15.0
            roofline_k6 = {"bound": "shared-memory wavefronts", "kernel": "k_ray_scatter", "achieved": per_s / 1e9,
                           "peak": peak / 1e9, "unit": "G lookups/s", "frac": per_s / peak,
                           "frac_padded_float4_model": per_s / (sm_count * clock_hz * 32.0 / 20.0),
                           "model": "1 wavefront (128 B) / clk / SM; 60 algorithmic bytes per lookup = 15 wavefronts per "
                                    "warp-lookup, %d SMs at the measured %.0f MHz" % (sm_count, clock_hz / 1e6),
                           "lookups_per_pass": lookups / cfg.iterations, "pass_ms": ray_ms / cfg.iterations,
                           "measured_counters": "profiles/r2/r2_final_ncu_summary.txt: 28.5 shared wavefronts and 121 "
                                                "thread-instructions per warp-lookup"}
        line = {"metric": METRIC, "value": texels / (ms_per_step * 1e-3), "unit": UNIT, "n_gpus": world,
                "steps": args.steps, "warmup": max(args.warmup, 3) + 2, "ms_per_step": ms_per_step,
                "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32+f64",
                "data": "synthetic", "build_time_s": ms_per_step * 1e-3,
                "config": {"workload": workload, "l2": "512 MiB flush write between timed steps",
                           "timing": "CUDA events per step on the builder's stream, max over ranks; one CUDA graph "
                                     "launch per build",
                           "exchange": "none (1 GPU)" if world == 1 else
                                       ("p2p: peer stores over NVLink + flag barriers, no collective" if args.mode == "p2p"
                                        else "nccl: one all-gather per table"),
                           "parallelism": ("single" if world == 1 else "%s%d" % (builder.mode, world)),
                           "wall_s_timed_region": t_wall},
                "clocks": clocks, "e2e": e2e, "gpu_launches": launches_per_step * args.steps,
                "kernel_launches_per_build": launches_per_step,
                "ray_scatter_pass_texels_per_s": (n4 * cfg.iterations / (ray_ms * 1e-3)) if ray_ms else None,
                "first_order_texels_per_s": (n4 / (first_ms * 1e-3)) if first_ms else None,
                "point_scatter_pass_texels_per_s": (n4 * cfg.iterations / (point_ms * 1e-3)) if point_ms else None,
                "stage_ms": {k: round(v, 4) for k, v in my_stages.items()},
                "stage_ms_max_over_ranks": {k: round(max(r[k] for r in all_stages), 4) for k in my_stages},
                "stage_ms_min_over_ranks": {k: round(min(r[k] for r in all_stages), 4) for k in my_stages},
                "stage_note": "stages from event-record nodes in a second CUDA graph of the same build, this rank; the "
                              "exchange stages absorb the waiting for the slowest rank",
                "work_per_step": work, "roofline": roofline, "roofline_k6": roofline_k6}
        if e2e_pageable:
            line["e2e_pageable"] = e2e_pageable
        if identical is not None:
            line["sharded_identical"] = all(identical.values())
            line["sharded_identical_detail"] = identical
        line.update(extras)
        if not args.no_cpu_baseline and world == 1 and args.workload == "shipped":
            line["cpu_baseline"] = cpu_baseline(args.cpu_sample)
        print(json.dumps(line), flush=True)
    builder.close()
    failed = identical is not None and not all(identical.values())
    if world > 1:
        dist.destroy_process_group()
    if failed:
        raise SystemExit("the sharded build differs from the single-GPU build: %s" % identical)


def main():
    args = parse()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_b200(args)


if __name__ == "__main__":
    main()
