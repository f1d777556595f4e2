"""ctypes binding of oracle/noise_oracle.c (TEST INFRASTRUCTURE ONLY): CPU restatement of sfsim.worley / sfsim.perlin."""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "liboracle_noise.so")
_lib = None
c_double_p = C.POINTER(C.c_double)


def lib():
    global _lib
    if _lib is None:
        src = os.path.join(_HERE, "noise_oracle.c")
        if os.path.exists(src) and (not os.path.exists(_LIB_PATH) or os.path.getmtime(_LIB_PATH) < os.path.getmtime(src)):
            subprocess.check_call(["make", "-C", _HERE, "-B", "liboracle_noise.so"], stdout=subprocess.DEVNULL)
        _lib = C.CDLL(_LIB_PATH)
        _lib.orc_closest_distance_to_point_in_grid.restype = C.c_double
        _lib.orc_ease_curve.restype = C.c_double
        _lib.orc_ease_curve.argtypes = [C.c_double]
        _lib.orc_perlin_noise_sample.restype = C.c_double
    return _lib


def _grid(grid):
    g = np.ascontiguousarray(grid, dtype=np.float64)
    assert g.ndim == 4 and g.shape[3] == 3
    return g


def _dp(a):
    return a.ctypes.data_as(c_double_p)


def _v3(v):
    return (C.c_double * 3)(*[float(x) for x in v])


def _dims(g):
    return (C.c_long * 3)(*g.shape[:3])


def extract_point_from_grid(grid, size, k, j, i):
    """worley.clj:58-66; grid[k][j][i] = point, any (dk, dj, di) shape"""
    g = _grid(grid)
    out = (C.c_double * 3)()
    lib().orc_extract_point_from_grid(_dp(g), _dims(g), C.c_long(size), C.c_long(k), C.c_long(j), C.c_long(i), out)
    return np.array(out[:])


def closest_distance_to_point_in_grid(grid, divisions, size, point):
    g = _grid(grid)
    return lib().orc_closest_distance_to_point_in_grid(_dp(g), _dims(g), C.c_long(divisions), C.c_long(size), _v3(point))


def worley_noise(grid, size):
    g = _grid(grid)
    out = np.zeros(size ** 3)
    lib().orc_worley_noise(_dp(g), C.c_long(g.shape[0]), C.c_long(size), _dp(out))
    return out


def ease_curve(t):
    return lib().orc_ease_curve(float(t))


def perlin_noise_sample(gradients, divisions, size, cell):
    g = _grid(gradients)
    return lib().orc_perlin_noise_sample(_dp(g), C.c_long(divisions), C.c_long(size), _v3(cell))


def perlin_noise(gradients, size):
    g = _grid(gradients)
    out = np.zeros(size ** 3)
    lib().orc_perlin_noise(_dp(g), C.c_long(g.shape[0]), C.c_long(size), _dp(out))
    return out
