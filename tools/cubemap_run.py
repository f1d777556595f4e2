"""Cube-map tile generation on one GPU with the shipped constants and synthetic rasters: device time per batch and
pixels/s (also the command the profilers run).  `python tools/cubemap_run.py [out_level] [batches]`"""
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402

from sfsim_b200 import synthetic as ocm  # noqa: E402
from sfsim_b200 import cubemap  # noqa: E402

out_level = int(sys.argv[1]) if len(sys.argv) > 1 else 4
batches = int(sys.argv[2]) if len(sys.argv) > 2 else 3
in_level = out_level - 3                                  # build.clj:300-310
cfg = cubemap.make_config(in_level, out_level)
ls = max(0, min(4, in_level))
lc = max(0, min(5, in_level + 1))
lw = max(0, min(4, in_level + 1))
t0 = time.time()
elev, day, night = ocm.synthetic_world(675, sorted({ls, lw}), [lc], seed=1)
w = cubemap.World(675)
for level, a in elev.items():
    w.set_elevation(level, a)
w.set_color(False, lc, day[lc])
w.set_color(True, lc, night[lc])
t_setup = time.time() - t0
tiles = cubemap.tile_shard(out_level)[:1536]
times = [w.time_cube_map_tiles(cfg, tiles) for _ in range(batches)]
ms = float(np.median(times[1:] if len(times) > 1 else times))
pixels = len(tiles) * cfg.color_tilesize ** 2
print(json.dumps({"out_level": out_level, "in_level": in_level, "tiles": len(tiles), "ms_per_batch": ms, "all_ms": times,
                  "colour_pixels_per_s": pixels / ms * 1e3, "tiles_per_s": len(tiles) / ms * 1e3,
                  "raster_setup_s": round(t_setup, 1)}))
w.close()
