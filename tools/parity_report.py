"""Stage-by-stage parity report: CUDA library vs the CPU oracle on one configuration.

Every stage of generate-atmosphere-luts is run through the library's per-table entry points with the
ORACLE's previous-stage tables as input (so errors do not compound) and through the whole build.
Prints one JSON line per stage with the max relative error per LUT entry.

    python tools/parity_report.py [--shape4 8 31 8 2] [--shape-t 16 63] [--shape-e 4 15] [--iterations 2]
"""
import argparse
import ctypes as C
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

from oracle import oracle as orc  # noqa: E402  (checker only)
from sfsim_b200 import _lib, atmosphere_lut  # noqa: E402


def rel_err(gpu, ref, floor):
    gpu = np.asarray(gpu, dtype=np.float64).reshape(-1)
    ref = np.asarray(ref, dtype=np.float64).reshape(-1)
    err = np.abs(gpu - ref) / np.maximum(np.abs(ref), floor)
    i = int(np.argmax(err))
    return float(err[i]), i, float(ref[i]), float(gpu[i])


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--shape4", type=int, nargs=4, default=[8, 31, 8, 2])
    ap.add_argument("--shape-t", type=int, nargs=2, default=[16, 63])
    ap.add_argument("--shape-e", type=int, nargs=2, default=[4, 15])
    ap.add_argument("--ray-steps", type=int, default=100)
    ap.add_argument("--sphere-steps", type=int, default=15)
    ap.add_argument("--iterations", type=int, default=2)
    ap.add_argument("--height", type=float, default=35000.0)
    ap.add_argument("--floor", type=float, default=1e-30)
    args = ap.parse_args()

    lib = _lib.load()
    planet = dict(atmosphere_lut.earth, height=args.height)
    cfg = _lib.make_config(ray_scatter_shape=args.shape4, transmittance_shape=args.shape_t,
                           surface_radiance_shape=args.shape_e, ray_steps=args.ray_steps,
                           sphere_steps=args.sphere_steps, iterations=args.iterations)
    pl = _lib.make_planet(planet)
    sc = _lib.make_scatter_array([atmosphere_lut.mie, atmosphere_lut.rayleigh])

    opl = orc.planet(planet["radius"], planet["height"], planet["brightness"])
    omie, oray = orc.scatter(**orc.MIE), orc.scatter(**orc.RAYLEIGH)
    ocfg = orc.config(args.shape4, args.shape_t, args.shape_e, args.ray_steps, args.sphere_steps)
    rec = {}
    t0 = time.time()
    ofiles = orc.generate_atmosphere_luts(opl, omie, oray, ocfg, iterations=args.iterations, record=rec)
    t_oracle = time.time() - t0

    def report(stage, gpu, ref):
        e, i, r, g = rel_err(gpu, ref, args.floor)
        print(json.dumps({"stage": stage, "max_rel_err": e, "at": i, "ref": r, "gpu": g}), flush=True)

    s4, st, se = tuple(args.shape4) + (3,), tuple(args.shape_t) + (3,), tuple(args.shape_e) + (3,)
    out = np.zeros(st, np.float32)
    _lib.check(lib.atmlut_transmittance_table(C.byref(pl), sc, 2, C.byref(cfg), _lib.ptr(out)))
    report("transmittance", out, rec["T"])
    out = np.zeros(se, np.float32)
    _lib.check(lib.atmlut_surface_radiance_base_table(C.byref(pl), sc, 2, C.byref(cfg), _lib.ptr(out)))
    report("surface_radiance_base", out, rec["Ebase"])
    r1, m1 = np.zeros(s4, np.float32), np.zeros(s4, np.float32)
    _lib.check(lib.atmlut_first_order_tables(C.byref(pl), sc, 2, C.byref(cfg), 1, 0, _lib.ptr(r1), 0, 1, _lib.ptr(m1)))
    report("first_order_rayleigh", r1, rec["R1"])
    report("first_order_mie_strength", m1, rec["M1"])
    ds_a, ds_b, de = _lib.f32(rec["R1"]), _lib.f32(rec["M1"]), _lib.f32(rec["Ebase"])
    for it in range(args.iterations):
        out = np.zeros(s4, np.float32)
        _lib.check(lib.atmlut_point_scatter_table(C.byref(pl), sc, 2, C.byref(cfg), _lib.ptr(ds_a),
                                                  _lib.ptr(ds_b) if ds_b is not None else None, 0, _lib.ptr(de),
                                                  _lib.ptr(out)))
        report("point_scatter_%d" % it, out, rec["dJ%d" % it])
        out = np.zeros(se, np.float32)
        _lib.check(lib.atmlut_surface_radiance_table(C.byref(pl), sc, 2, C.byref(cfg), _lib.ptr(ds_a),
                                                     _lib.ptr(ds_b) if ds_b is not None else None, 0, _lib.ptr(out)))
        report("surface_radiance_%d" % it, out, rec["dE%d" % it])
        out = np.zeros(s4, np.float32)
        dj = _lib.f32(rec["dJ%d" % it])
        _lib.check(lib.atmlut_ray_scatter_table(C.byref(pl), sc, 2, C.byref(cfg), _lib.ptr(dj), _lib.ptr(out)))
        report("ray_scatter_%d" % it, out, rec["dS%d" % it])
        out = np.zeros(s4, np.float32)
        s_prev = _lib.f32(rec["R1"] if it == 0 else rec["S%d" % (it - 1)])
        ds_new = _lib.f32(rec["dS%d" % it])
        _lib.check(lib.atmlut_resample_table(C.byref(pl), C.byref(cfg), 0, _lib.ptr(s_prev), _lib.ptr(ds_new),
                                             _lib.ptr(out)))
        report("accumulate_s_%d" % it, out, rec["S%d" % it])
        ds_a, ds_b, de = ds_new, None, _lib.f32(rec["dE%d" % it])
    t0 = time.time()
    files = atmosphere_lut.generate_tables(planet, (atmosphere_lut.mie, atmosphere_lut.rayleigh), cfg)
    t_gpu = time.time() - t0
    for name, g, o in zip(atmosphere_lut.FILE_NAMES, files, ofiles):
        report("file:" + name, g, o)
    print(json.dumps({"oracle_s": t_oracle, "gpu_generate_s": t_gpu, "oracle_threads": orc.num_threads(),
                      "oracle_counters": None}))


if __name__ == "__main__":
    main()
