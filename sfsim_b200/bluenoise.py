"""Host-side mirror of sfsim.bluenoise (src/clj/sfsim/bluenoise.clj) over libsfsim_atmosphere.so (include/sfsim_noise.h).

`blue_noise(m, n, sigma)` is the drop-in for the reference function (build.clj:45-51): the seed picks are drawn on the
host like `pick-n` (bluenoise.clj:35-40), the void-and-cluster phases run on the GPU.
"""
import ctypes as C
import math
import random as _random

import numpy as np

from . import _lib
from ._lib import check

noise_size = 64    # bluenoise.clj:20


def indices_2d(m):
    """bluenoise.clj:25-28"""
    return list(range(m * m))


def pick_n(arr, n, order=None):
    """bluenoise.clj:31-35: the first n of `order(arr)` (default: a shuffle)"""
    if order is None:
        arr = list(arr)
        _random.shuffle(arr)
        return arr[:n]
    return list(order(arr))[:n]


def scatter_mask(arr, m):
    """bluenoise.clj:41-44"""
    mask = [False] * (m * m)
    for i in arr:
        mask[i] = True
    return mask


def density_function(sigma):
    """bluenoise.clj:50-53"""
    return lambda dx, dy: math.exp(-((dx * dx + dy * dy) / (2.0 * sigma * sigma)))


def density_table(m, f):
    """f over the offsets `wrap` (bluenoise.clj:73-78) can return, in the layout the library takes"""
    off = m // 2
    return np.array([[float(f(dx - off, dy - off)) for dx in range(m)] for dy in range(m)], dtype=np.float64)


def blue_noise(m, n, sigma, picks=None, f=None):
    """bluenoise.clj:175-185; returns the dither array int32[m * m].  `picks` replaces the random seed indices, `f` the
    density function."""
    lib = _lib.load()
    if picks is None:
        picks = pick_n(indices_2d(m), n)
    picks = _lib.i32(picks)
    table = np.ascontiguousarray(density_table(m, f or density_function(sigma)))
    dither = np.zeros(m * m, dtype=np.int32)
    check(lib.sfsim_blue_noise(_lib.ptr(picks), len(picks), int(m), _lib.ptr(table), _lib.ptr(dither)))
    return dither


def blue_noise_texture(m=noise_size, n=None, sigma=1.5, picks=None):
    """The float array `clj -T:build bluenoise` writes to data/bluenoise.raw (build.clj:45-51): n = m^2 / 10 seed
    samples, sigma 1.5, values dither / m / m."""
    lib = _lib.load()
    n = (m * m) // 10 if n is None else n
    if picks is None:
        picks = pick_n(indices_2d(m), n)
    picks = _lib.i32(picks)
    out = np.zeros(m * m, dtype=np.float32)
    check(lib.sfsim_blue_noise_texture(_lib.ptr(picks), len(picks), int(m), C.c_double(sigma), _lib.ptr(out)))
    return out
