"""Host-side mirror of sfsim.perlin (src/clj/sfsim/perlin.clj) over libsfsim_atmosphere.so (include/sfsim_noise.h).

`perlin_noise(divisions, size)` is the drop-in for the reference function (build.clj:40-43): the gradient grid is
drawn on the host like `random-gradient-grid` (perlin.clj:38-50), sampling and normalisation run on the GPU.
"""
import random as _random

import numpy as np

from . import _lib
from .worley import _run

# perlin.clj:31-35
GRADIENTS = [(1, 1, 0), (-1, 1, 0), (1, -1, 0), (-1, -1, 0), (1, 0, 1), (-1, 0, 1), (1, 0, -1), (-1, 0, -1),
             (0, 1, 1), (0, -1, 1), (0, 1, -1), (0, -1, -1)]


def random_gradient(selector=None):
    """perlin.clj:26-35: `selector` picks one of the twelve gradient vectors (default rand-nth)"""
    selector = selector or _random.choice
    return tuple(float(c) for c in selector(GRADIENTS))


def random_gradient_grid(divisions, random_gradient_fn=None):
    """perlin.clj:38-50: grid[z][y][x]"""
    fn = random_gradient_fn or random_gradient
    grid = np.zeros((divisions, divisions, divisions, 3))
    for z in range(divisions):
        for y in range(divisions):
            for x in range(divisions):
                grid[z, y, x] = fn()
    return grid


def perlin_samples(gradients, size):
    """perlin-noise-sample (perlin.clj:111-119) at every cell of perlin-noise, un-normalised, in double"""
    return _run(_lib.load().sfsim_perlin_samples, gradients, size, np.float64)


def perlin_noise(divisions, size, gradients=None):
    """perlin.clj:122-137; returns float32[size^3] in (k, j, i) order normalised to [0, 1]"""
    if gradients is None:
        gradients = random_gradient_grid(divisions)
    return _run(_lib.load().sfsim_perlin_noise, gradients, size, np.float32)
