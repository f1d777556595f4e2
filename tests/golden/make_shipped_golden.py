"""Generates tests/golden/shipped_build.npz: a sampled golden of the FULL shipped-resolution build.

    python tests/golden/make_shipped_golden.py            (about 10 minutes on 8 cores)

Runs the CPU oracle (oracle/, double precision, the reference's algorithm sample for sample) through the whole
of generate-atmosphere-luts (atmosphere_lut.clj:43-105) at the shipped configuration -- 4-D 32x127x32x8,
T 64x255, E 16x63, ray-steps 100, sphere-steps 15, 5 iterations, Earth defaults -- and keeps

* the complete transmittance and surface-radiance files (16 320 and 1 008 texels),
* N4_SAMPLES seeded random texels of ray-scatter.scatter and mie-strength.scatter (flat 4-D texel index
  h,e,s,a row-major + value), i.e. the result of all 5 scattering orders and every re-tabulation,
* the same texels of the intermediate tables (dJ, dS, S) and the whole dE / E tables of iterations 1, 2 and 5, so a
  deviation can be traced to the order it first appears in.

Values are stored as float64 (before pack-matrices' cast).  The reference repository holds no shipped-resolution
file or checksum, so this fixture is pinned through the oracle, which is itself pinned to the
reference's known answers (tests/test_oracle_known_answers.py).
"""
import os
import sys
import time

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))

from oracle import oracle as orc  # noqa: E402

SHIPPED = dict(shape4=(32, 127, 32, 8), shape_t=(64, 255), shape_e=(16, 63), ray_steps=100, sphere_steps=15)
N4_SAMPLES = 8192
SEED = 20261017


def sample_indices(n4, count=N4_SAMPLES, seed=SEED):
    return np.sort(np.random.default_rng(seed).choice(n4, size=count, replace=False)).astype(np.int64)


def main():
    pl = orc.planet(**orc.EARTH)
    mie, ray = orc.scatter(**orc.MIE), orc.scatter(**orc.RAYLEIGH)
    cfg = orc.config(SHIPPED["shape4"], SHIPPED["shape_t"], SHIPPED["shape_e"], SHIPPED["ray_steps"],
                     SHIPPED["sphere_steps"])
    n4 = int(np.prod(SHIPPED["shape4"]))
    idx = sample_indices(n4)
    rec = {}
    t0 = time.time()
    orc.generate_atmosphere_luts(pl, mie, ray, cfg, iterations=5, record=rec,
                                 log=lambda m: print("%7.1f s  %s" % (time.time() - t0, m), flush=True))
    print("oracle build: %.1f s on %d threads" % (time.time() - t0, orc.num_threads()), flush=True)
    keep = {"idx": idx, "LT": rec["LT"], "LE": rec["LE"], "T": rec["T"], "Ebase": rec["Ebase"]}
    for name in ("LS", "LM", "R1", "M1"):
        keep[name] = rec[name].reshape(n4, 3)[idx]
    for it in (0, 1, 4):                                  # the first two orders step by step, and the last one
        for name in ("dJ", "dS", "S"):
            keep["%s%d" % (name, it)] = rec["%s%d" % (name, it)].reshape(n4, 3)[idx]
        keep["dE%d" % it] = rec["dE%d" % it]
        keep["E%d" % it] = rec["E%d" % it]
    np.savez_compressed(os.path.join(HERE, "shipped_build.npz"), **keep)
    print("wrote shipped_build.npz")


if __name__ == "__main__":
    main()
