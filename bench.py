#!/usr/bin/env python
"""Benchmark of the atmosphere-LUT build (BASELINE.json: "atmosphere-lut build time (s) and ray-scatter
texels/s at 1/2/4/8 B200").

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]

One *step* = one full `generate-atmosphere-luts` build (atmosphere_lut.clj:43-105) at shipped resolution:
transmittance, surface-radiance-base, first-order ray scatter (Rayleigh + Mie strength), 5 iterations of
point-scatter / surface-radiance / ray-scatter / accumulation, and the final resampling into file layout.
Inputs are parameters only (deterministic, no data files), so "inputs resident in HBM" means the builder
with its quadrature tables and table buffers exists before the timed region.

metric = ray-scatter texels per second = 4-D texels x (1 + iterations) scattering orders / build time.
With N > 1 (torchrun, one process per GPU) the SAME build is sharded over the ranks (strong scaling): each
rank integrates a contiguous slab of (height, elevation) pairs of every 4-D table and the tables are
reassembled by one NCCL all-gather each (12 per build).

`--impl reference` times the reference algorithm on the host cores: the double-precision CPU oracle
(oracle/, a restatement of the Clojure code -- there is no JVM on the box) on a bounded random sample of
texels of every stage, extrapolated per texel to the full build.
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
    """Times the CPU oracle on `sample` random texels of every stage and extrapolates to the full build."""
    import numpy as np
    from oracle import oracle as orc
    if threads:
        orc.set_num_threads(threads)
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
            "note": "CPU restatement in C (double, OpenMP), not JVM Clojure: a lower bound on the reference's time",
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
    cfg = _lib.default_config()
    workload = WORKLOAD
    if args.workload == "stress":
        cfg = _lib.make_config(ray_scatter_shape=(64, 253, 64, 16), iterations=10)
        workload = "stress: 4-D 64x253x64x16 (2x shipped per axis), T 64x255, E 16x63, ray-steps 100, sphere-steps 15, " \
                   "10 iterations, Earth defaults"
    builder = atmosphere_lut.AtmosphereLutBuilder(cfg=cfg, rank=rank, world=world, mode=args.mode)
    stream = torch.cuda.ExternalStream(lib.atmlut_stream())
    flush = torch.empty(512 << 20, dtype=torch.uint8, device="cuda")   # > 126 MB L2

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def step_device():
        builder.run()

    for _ in range(max(args.warmup, 3)):
        step_device()
    barrier()
    sampler = ClockSampler(local_rank)
    sampler.start()
    starts = [torch.cuda.Event(enable_timing=True) for _ in range(args.steps)]
    ends = [torch.cuda.Event(enable_timing=True) for _ in range(args.steps)]
    barrier()
    t_wall0 = time.perf_counter()
    for i in range(args.steps):
        with torch.cuda.stream(stream):
            flush.zero_()                      # evict the previous step's tables from L2
        starts[i].record(stream)
        step_device()
        ends[i].record(stream)
    barrier()
    t_wall = time.perf_counter() - t_wall0
    clocks = sampler.stop()
    step_ms = [s.elapsed_time(e) for s, e in zip(starts, ends)]
    total_ms = torch.tensor([sum(step_ms)], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(total_ms, op=dist.ReduceOp.MAX)
    total_ms = float(total_ms.item())
    ms_per_step = total_ms / args.steps
    stage_times = builder.stage_times()
    work = builder.work()
    # per-rank stage times (the barrier / all-gather stages absorb the load imbalance between ranks)
    my_stages = {}
    for name, ms in stage_times:
        key = name.split("_", 1)[1] if name.startswith("iter") else name
        my_stages[key] = my_stages.get(key, 0.0) + ms
    all_stages = [my_stages]
    if world > 1:
        all_stages = [None] * world
        dist.all_gather_object(all_stages, my_stages)

    # end to end through the public one-shot call: host parameters in, host (pinned) tables out
    e2e = None
    if world == 1:
        out = atmosphere_lut.allocate_outputs(cfg, pinned=True)
        for _ in range(3):
            atmosphere_lut.generate_tables(cfg=cfg, out=out)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        n_e2e = max(5, args.steps // 3)
        for _ in range(n_e2e):
            atmosphere_lut.generate_tables(cfg=cfg, out=out)
        e2e_s = (time.perf_counter() - t0) / n_e2e
    else:
        # sharded: device build + download of the file-layout tables on rank 0
        out = atmosphere_lut.allocate_outputs(cfg, pinned=True)
        barrier()
        t0 = time.perf_counter()
        n_e2e = max(5, args.steps // 3)
        for _ in range(n_e2e):
            builder.run()
            if rank == 0:
                builder.download(out)
            else:
                builder.sync()
            if world > 1:
                dist.barrier()
        e2e_s = (time.perf_counter() - t0) / n_e2e
    n4 = cfg.height_size * cfg.elevation_size * cfg.light_elevation_size * cfg.heading_size
    texels = n4 * (1 + cfg.iterations)
    d2h = sum(int(np.prod(s)) * 4 for s in atmosphere_lut.output_shapes(cfg))
    h2d = ctypes.sizeof(_lib.Planet) + 2 * ctypes.sizeof(_lib.Scatter) + ctypes.sizeof(_lib.Config)
    e2e = {"value": texels / e2e_s, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
           "build_time_s": e2e_s, "api": "atmlut_generate" if world == 1 else "atmlut_builder_run+download"}

    # roofline of the dominant E-sample kernel (first-order ray scatter): MUFU.EX2 pipe
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "profiles", "pipe_peaks_b200.json")))
    except Exception:
        pass
    first_ms = dict(stage_times).get("first_order")
    esamples_first = None
    roofline = None
    if first_ms:
        # E-samples of this rank's first-order launch: counter 0 of the builder (exact, counted on the device)
        esamples_first = work.get("esamples_first_order")
    if rank == 0:
        stage_dict = {}
        for name, ms in stage_times:
            key = name.split("_", 1)[1] if name.startswith("iter") else name
            stage_dict[key] = stage_dict.get(key, 0.0) + ms
        launches_per_step = int(work["kernel_launches"])
        mufu_peak = peaks.get("mufu_ex2_per_s")
        # DRAM traffic of one K3 launch from the committed `ncu --set full` capture, and the driver-measured HBM peak
        traffic = None
        try:
            text = open(os.path.join(ROOT, "profiles", "r1_final_ncu_summary.txt")).read().split("== k_point_scatter")[0]
            units = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
            traffic = 0.0
            for line in text.splitlines():
                if "dram__bytes_read.sum" in line or "dram__bytes_write.sum" in line:
                    parts = line.split()
                    traffic += float(parts[1]) * units[parts[2]]
        except Exception:
            traffic = None
        hbm_peak = None
        try:
            hbm_peak = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"]
        except Exception:
            pass
        if first_ms and work.get("esamples_first_order") and mufu_peak:
            achieved = 2.0 * work["esamples_first_order"] / (first_ms * 1e-3)
            roofline = {"bound": "sfu", "kernel": "k_first_order", "achieved": achieved / 1e9, "peak": mufu_peak / 1e9,
                        "unit": "Gop/s (MUFU.EX2)", "frac": achieved / mufu_peak, "traffic": traffic,
                        "traffic_note": "dram read+write bytes of one launch, profiles/r1_final_ncu_summary.txt",
                        "hbm": {"algorithmic_bytes": 2 * n4 // world * 16,
                                "achieved_GBs": 2 * n4 / world * 16 / (first_ms * 1e-3) / 1e9, "peak_GBs": hbm_peak,
                                "note": "two float4 output tables per launch; the kernel is SFU-bound, not HBM-bound"},
                        "peak_source": "measured on this pool's B200 by tools/pipe_peaks.cu (profiles/pipe_peaks_b200.json)",
                        "algorithmic_units": "2 MUFU.EX2 per overall-extinction sample; samples counted on the device",
                        "esamples_per_launch": work["esamples_first_order"], "launch_ms": first_ms}
        line = {"metric": METRIC, "value": texels / (ms_per_step * 1e-3), "unit": UNIT, "n_gpus": world,
                "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": ms_per_step,
                "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32+f64",
                "data": "synthetic", "build_time_s": ms_per_step * 1e-3,
                "config": {"workload": workload, "l2": "512 MiB flush write between timed steps",
                           "timing": "CUDA events per step on the library stream, max over ranks",
                           "parallelism": ("single" if world == 1 else "%s%d" % (builder.mode, world)), "wall_s_timed_region": t_wall},
                "clocks": clocks, "e2e": e2e, "gpu_launches": launches_per_step * args.steps,
                "stage_ms": {k: round(v, 4) for k, v in stage_dict.items()},
                "stage_ms_max_over_ranks": {k: round(max(r[k] for r in all_stages), 4) for k in my_stages},
                "stage_ms_min_over_ranks": {k: round(min(r[k] for r in all_stages), 4) for k in my_stages},
                "work_per_step": work, "roofline": roofline,
                "lookup_kernels": {
                    "note": "issue-slot / FP64 bound gather kernels; no single-pipe roofline applies (ncu: profiles/)",
                    "ray_scatter_lookups_per_s": (n4 / world) * 100 * cfg.iterations / max(stage_dict.get("ray_scatter", 0.0) * 1e-3, 1e-12),
                    "point_scatter_direction_evals_per_s": (n4 / world) * 71 * cfg.iterations / max(stage_dict.get("point_scatter", 0.0) * 1e-3, 1e-12)}}
        if not args.no_cpu_baseline and world == 1 and args.workload == "shipped":
            line["cpu_baseline"] = cpu_baseline(args.cpu_sample)
        print(json.dumps(line), flush=True)
    builder.close()
    if world > 1:
        dist.destroy_process_group()


def main():
    args = parse()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_b200(args)


if __name__ == "__main__":
    main()
