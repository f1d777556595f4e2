"""Small builds for compute-sanitizer (memcheck / racecheck): reduced tables with and without graph replay, the
kparts > 1 path, wide tiles, plus the batch evaluators."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from sfsim_b200 import _lib, atmosphere, atmosphere_lut, interpolate  # noqa: E402

which = sys.argv[1:] or ["reduced", "small", "tiny", "wide", "batch"]
if "reduced" in which:   # BASELINE.json configs[0] shapes, fewer steps
    cfg = _lib.make_config(ray_scatter_shape=(8, 31, 8, 2), transmittance_shape=(16, 63), surface_radiance_shape=(4, 15),
                           ray_steps=20, sphere_steps=15, iterations=3)
    print("reduced", [float(t.max()) for t in atmosphere_lut.generate_tables(cfg=cfg)], flush=True)
if "small" in which:
    cfg = _lib.make_config(ray_scatter_shape=(3, 7, 8, 4), transmittance_shape=(6, 11), surface_radiance_shape=(3, 5),
                           ray_steps=16, sphere_steps=6, iterations=2)
    print("small", [float(t.max()) for t in atmosphere_lut.generate_tables(cfg=cfg)], flush=True)
if "tiny" in which:      # kparts > 1 path
    cfg2 = _lib.make_config(ray_scatter_shape=(2, 3, 2, 2), transmittance_shape=(4, 5), surface_radiance_shape=(2, 3),
                            ray_steps=8, sphere_steps=4, iterations=1)
    print("tiny", [float(t.max()) for t in atmosphere_lut.generate_tables(cfg=cfg2)], flush=True)
if "wide" in which:      # several texel chunks per pair in the point-scatter kernel, generic ray-scatter kernel
    cfg3 = _lib.make_config(ray_scatter_shape=(2, 3, 40, 32), transmittance_shape=(3, 5), surface_radiance_shape=(2, 3),
                            ray_steps=8, sphere_steps=4, iterations=2)
    print("wide", [float(t.max()) for t in atmosphere_lut.generate_tables(cfg=cfg3)], flush=True)
if "batch" in which:
    earth = dict(atmosphere_lut.earth)
    sc = [atmosphere_lut.mie, atmosphere_lut.rayleigh]
    print(atmosphere.transmittance(earth, sc, 10, (0, 6378000.0, 0), (0, 1, 0), True))
    space = atmosphere.ray_scatter_space(earth, (3, 7, 8, 4))
    src = atmosphere.FirstOrder(atmosphere.FirstOrder.COMPONENT, earth, sc, atmosphere_lut.rayleigh, 8, (1, 1, 1))
    tab = interpolate.interpolate_function(interpolate.RayScatter(earth, sc, 8, src), space)
    e = interpolate.interpolation_table(np.zeros((3, 5, 3), np.float32), atmosphere.surface_radiance_space(earth, (3, 5)))
    print(atmosphere.point_scatter(earth, sc, tab, e, (1, 1, 1), 6, 8, (6379000.0, 0, 0), (0, 1, 0), (0.6, 0.8, 0)))
    print(atmosphere.surface_radiance(earth, tab, 8, (6379000.0, 0, 0), (0.6, 0.8, 0)))
