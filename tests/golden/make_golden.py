"""Regenerates the golden fixtures in this directory from the CPU oracle.

    python tests/golden/make_golden.py

* floats.raw: the 16 bytes of the reference's test/clj/sfsim/fixtures/util/floats.raw (t_util.clj:54).
* reduced_*.scatter: the four output files of generate-atmosphere-luts for BASELINE.json configs[0]
  (4-D [8,31,8,2], T [16,63], E [4,15], ray-steps 100, sphere-steps 15, 5 iterations, Earth defaults),
  headerless little-endian float32 in file layout, produced by oracle/ (double precision, cast at the end).
* reduced_first_order.npz: the un-resampled first-order tables of the same configuration.
* small_shader_luts.npz: the size-12 / ray-steps-10 / 100 km tables of t_atmosphere.clj:569-578.
"""
import os
import struct
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))

from oracle import oracle as orc  # noqa: E402

REDUCED = dict(shape4=(8, 31, 8, 2), shape_t=(16, 63), shape_e=(4, 15), ray_steps=100, sphere_steps=15)


def main():
    open(os.path.join(HERE, "floats.raw"), "wb").write(struct.pack("<4f", 2.0, 3.0, 5.0, 7.0))
    pl = orc.planet(**orc.EARTH)
    mie, ray = orc.scatter(**orc.MIE), orc.scatter(**orc.RAYLEIGH)
    cfg = orc.config(REDUCED["shape4"], REDUCED["shape_t"], REDUCED["shape_e"], REDUCED["ray_steps"],
                     REDUCED["sphere_steps"])
    rec = {}
    files = orc.generate_atmosphere_luts(pl, mie, ray, cfg, iterations=5, record=rec)
    for name, data in zip(("transmittance", "surface-radiance", "ray-scatter", "mie-strength"), files):
        orc.spit_floats(os.path.join(HERE, "reduced_%s.scatter" % name), data)
    np.savez_compressed(os.path.join(HERE, "reduced_first_order.npz"), R1=rec["R1"].astype(np.float32),
                        M1=rec["M1"].astype(np.float32), Ebase=rec["Ebase"].astype(np.float32),
                        T=rec["T"].astype(np.float32))
    size = 12
    earth = orc.planet(6378000.0, 100000.0)
    scfg = orc.config((size,) * 4, (size, size), (size, size), ray_steps=10)
    T = orc.table_transmittance(earth, [mie, ray], scfg)
    S = orc.table_first_order(earth, [mie, ray], scfg, ray, 0)
    np.savez_compressed(os.path.join(HERE, "small_shader_luts.npz"), T=T.astype(np.float32), S=S.astype(np.float32))


if __name__ == "__main__":
    main()
