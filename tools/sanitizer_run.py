"""Small builds for compute-sanitizer (memcheck / racecheck): reduced tables with and without graph replay, the
kparts > 1 path, wide tiles, the batch evaluators, and cube-map tiles (batch call and the streamed level)."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from sfsim_b200 import _lib, atmosphere, atmosphere_lut, interpolate  # noqa: E402

which = sys.argv[1:] or ["reduced", "small", "tiny", "wide", "batch", "cubemap"]
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
if "cubemap" in which:
    from sfsim_b200 import synthetic as ocm
    from sfsim_b200 import cubemap
    width = 16
    elev, day, night = ocm.synthetic_world(width, [0, 1], [1], seed=3)
    w = cubemap.World(width)
    for level in (0, 1):
        w.set_elevation(level, elev[level])
    w.set_color(False, 1, day[1])
    w.set_color(True, 1, night[1])
    cfg = cubemap.make_config(0, 1, width=width, surface_tilesize=9)
    out = w.make_cube_map_tiles(cfg, cubemap.tile_shard(1))
    seen = []
    w.make_cube_map(0, 1, lambda key, tile: seen.append(int(tile["day"][0, 0, 0])), batch=5, surface_tilesize=9)
    print("cubemap", out["day"].shape, len(seen), float(np.abs(out["surface"]).max()),
          w.project_onto_globe([(1.0, 2.0, 3.0)], 1).tolist(), flush=True)
    w.close()
