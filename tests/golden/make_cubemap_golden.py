"""Regenerates tests/golden/cubemap_tiles.npz from the CPU oracle (oracle/cubemap_oracle.c, pinned to t_cubemap.clj).

    python tests/golden/make_cubemap_golden.py

Two cube-map tiles of globe.clj:29-80 from the deterministic synthetic rasters of oracle/cubemap.py (map tile width 48,
surface tile 33 / colour tile 65 pixels): face 2 at (in-level 0, out-level 2) and face 5 at (in-level -2, out-level 1),
with the un-truncated colour / water values (`raw`) the byte comparison needs at truncation boundaries.
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))

from oracle import cubemap as ocm  # noqa: E402

WIDTH, SURFACE_TILESIZE, SEED = 48, 33, 11
TILES = [dict(face=2, in_level=0, out_level=2, b=1, a=3), dict(face=5, in_level=-2, out_level=1, b=0, a=1)]


def world():
    elev, day, night = ocm.synthetic_world(WIDTH, [0, 1], [0, 1], seed=SEED)
    return elev, day, night


def main():
    elev, day, night = world()
    ow = ocm.OracleWorld(WIDTH, elev, day, night)
    out = {}
    for n, t in enumerate(TILES):
        tile = ow.make_cube_map_tile(t["face"], t["in_level"], t["out_level"], t["b"], t["a"], surface_tilesize=SURFACE_TILESIZE)
        for k, v in tile.items():
            out["%d_%s" % (n, k)] = v
    np.savez_compressed(os.path.join(HERE, "cubemap_tiles.npz"), **out)
    print({k: v.shape for k, v in out.items()})


if __name__ == "__main__":
    main()
