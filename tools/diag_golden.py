"""Diagnose deviations of the GPU chain from tests/golden/shipped_build.npz: which texels, which kernel."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import oracle as orc  # noqa: E402
from sfsim_b200 import _lib  # noqa: E402
from tests.test_gpu_tables import Lib  # noqa: E402


def rel(gpu, ref):
    gpu = np.asarray(gpu, dtype=np.float64)
    ref = np.asarray(ref, dtype=np.float64)
    return np.abs(gpu - ref) / np.maximum(np.abs(ref), 1e-20)


def report(name, gpu, ref, idx, shape4, top=8):
    r = rel(gpu, ref).max(axis=1)
    order = np.argsort(-r)[:top]
    print("%s: max %.3g, texels above 1e-4: %d of %d" % (name, r.max(), int((r > 1e-4).sum()), len(r)))
    H, E, S, A = shape4
    for o in order:
        i = int(idx[o])
        a, s, e, h = i % A, (i // A) % S, (i // (A * S)) % E, i // (A * S * E)
        print("   texel %8d (h %2d e %3d s %2d a %d)  rel %.3g  gpu %s  ref %s" %
              (i, h, e, s, a, r[o], np.array2string(np.asarray(gpu[o], dtype=np.float64), precision=6),
               np.array2string(np.asarray(ref[o]), precision=6)))


def main():
    g = np.load(os.path.join(os.path.dirname(__file__), "..", "tests", "golden", "shipped_build.npz"))
    cfg = _lib.default_config()
    shape4 = cfg.ray_scatter_shape
    lib = Lib(cfg)
    idx = g["idx"]
    pl = orc.planet(**orc.EARTH)
    ocfg = orc.config(cfg.ray_scatter_shape, cfg.transmittance_shape, cfg.surface_radiance_shape)
    mie, ray = orc.scatter(**orc.MIE), orc.scatter(**orc.RAYLEIGH)
    r1, m1 = lib.first_order()
    report("R1 vs golden", r1.reshape(-1, 3)[idx], g["R1"], idx, shape4, 3)
    report("M1 vs golden", m1.reshape(-1, 3)[idx], g["M1"], idx, shape4, 3)
    de = lib.surface_radiance_base()
    dj = lib.point_scatter(r1, m1, de)
    report("dJ0 vs golden", dj.reshape(-1, 3)[idx], g["dJ0"], idx, shape4, 5)
    ds = lib.ray_scatter(dj)
    report("dS0 vs golden (chain)", ds.reshape(-1, 3)[idx], g["dS0"], idx, shape4, 12)
    r = rel(ds.reshape(-1, 3)[idx], g["dS0"]).max(axis=1)
    bad = idx[np.argsort(-r)[:64]]
    bad = np.sort(bad)
    same_in = orc.table_ray_scatter(pl, [mie, ray], ocfg, dj, bad)
    report("dS0 vs oracle on the SAME dJ (kernel only), 64 worst texels", ds.reshape(-1, 3)[bad], same_in, bad, shape4, 12)
    pos = np.searchsorted(idx, bad)
    report("oracle(dJ gpu) vs golden dS0 (input sensitivity)", same_in, g["dS0"][pos], bad, shape4, 12)


if __name__ == "__main__":
    main()
