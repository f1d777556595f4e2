"""Host-side mirror of sfsim.worley (src/clj/sfsim/worley.clj) over libsfsim_atmosphere.so (include/sfsim_noise.h).

`worley_noise(divisions, size)` is the drop-in for the reference function of the same name (build.clj:34-38): the
random point grid is drawn on the host exactly like `random-point-grid` (worley.clj:27-44, any `random(cellsize)`
source may be passed, as in the reference), the size^3 closest-distance samples and the normalisation run on the GPU.
"""
import ctypes as C
import random as _random

import numpy as np

from . import _lib
from ._lib import check

worley_size = 16    # worley.clj:24


def random_point_grid(divisions, size, random=None):
    """worley.clj:27-44: grid[k][j][i] = (i, j, k) * cellsize + (random(cellsize), random(cellsize), random(cellsize));
    `random(n)` returns a number in [0, n) (clojure.core/rand)."""
    random = random or (lambda n: _random.random() * n)
    cellsize = size / divisions if size % divisions else size // divisions
    grid = np.zeros((divisions, divisions, divisions, 3))
    for k in range(divisions):
        for j in range(divisions):
            for i in range(divisions):
                grid[k, j, i] = (i * cellsize + random(cellsize), j * cellsize + random(cellsize),
                                 k * cellsize + random(cellsize))
    return grid


def _run(fn, grid, size, dtype):
    grid = _lib.f64(grid)
    if grid.ndim != 4 or grid.shape[3] != 3 or not (grid.shape[0] == grid.shape[1] == grid.shape[2]):
        raise TypeError("the grid must have shape [divisions][divisions][divisions][3]")
    out = np.zeros(int(size) ** 3, dtype=dtype)
    check(fn(_lib.ptr(grid), int(grid.shape[0]), int(size), _lib.ptr(out)))
    return out


def closest_distances(grid, size):
    """closest-distance-to-point-in-grid (worley.clj:69-80) at every sample point of worley-noise, in double"""
    return _run(_lib.load().sfsim_worley_distances, grid, size, np.float64)


def worley_noise(divisions, size, grid=None, random=None):
    """worley.clj:95-112; returns float32[size^3] in (k, j, i) order, ready for spit-floats.  `grid` replaces the
    random point grid (the reference's tests rebind random-point-grid, t_worley.clj:70-74)."""
    if grid is None:
        grid = random_point_grid(divisions, size, random)
    return _run(_lib.load().sfsim_worley_noise, grid, size, np.float32)
