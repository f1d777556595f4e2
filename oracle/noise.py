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


# ------------------------------------------------------------------ blue noise (bluenoise.clj)

c_long_p = C.POINTER(C.c_long)
c_ubyte_p = C.POINTER(C.c_ubyte)


def density_table(m, f):
    """ftab[(dy + m//2) * m + (dx + m//2)] = f(dx, dy) over the offsets `wrap` can return"""
    off = m // 2
    return np.array([[float(f(dx - off, dy - off)) for dx in range(m)] for dy in range(m)], dtype=np.float64)


def density_function(sigma):
    """bluenoise.clj:53-56"""
    import math
    return lambda dx, dy: math.exp(-((dx * dx + dy * dy) / (2.0 * sigma * sigma)))


def wrap(x, m):
    lib().orc_wrap.restype = C.c_long
    return lib().orc_wrap(C.c_long(x), C.c_long(m))


def _mask(mask):
    return np.ascontiguousarray(np.asarray(mask, dtype=bool).astype(np.uint8))


def argmax_with_mask(arr, mask):
    a, k = np.ascontiguousarray(arr, dtype=np.float64), _mask(mask)
    lib().orc_argmax_with_mask.restype = C.c_long
    return lib().orc_argmax_with_mask(_dp(a), k.ctypes.data_as(c_ubyte_p), C.c_long(len(a)))


def argmin_with_mask(arr, mask):
    a, k = np.ascontiguousarray(arr, dtype=np.float64), _mask(mask)
    lib().orc_argmin_with_mask.restype = C.c_long
    return lib().orc_argmin_with_mask(_dp(a), k.ctypes.data_as(c_ubyte_p), C.c_long(len(a)))


def density_sample(mask, m, f, cx, cy):
    k, t = _mask(mask), density_table(m, f)
    lib().orc_density_sample.restype = C.c_double
    return lib().orc_density_sample(k.ctypes.data_as(c_ubyte_p), C.c_long(m), _dp(t), C.c_long(cx), C.c_long(cy))


def density_array(mask, m, f):
    k, t = _mask(mask), density_table(m, f)
    out = np.zeros(m * m)
    lib().orc_density_array(k.ctypes.data_as(c_ubyte_p), C.c_long(m), _dp(t), _dp(out))
    return out


def density_change(density, m, sign, f, index):
    d, t = np.array(density, dtype=np.float64), density_table(m, f)
    lib().orc_density_change(_dp(d), C.c_long(m), C.c_int(sign), _dp(t), C.c_long(index))
    return d


def seed_pattern(mask, m, f):
    k, t = _mask(mask).copy(), density_table(m, f)
    lib().orc_seed_pattern(k.ctypes.data_as(c_ubyte_p), C.c_long(m), _dp(t))
    return k.astype(bool)


def dither_phase1(mask, m, n, f):
    k, t = _mask(mask), density_table(m, f)
    dither = np.zeros(m * m, dtype=np.int64)
    lib().orc_dither_phase1(k.ctypes.data_as(c_ubyte_p), C.c_long(m), C.c_long(n), _dp(t), dither.ctypes.data_as(c_long_p))
    return dither


def dither_phase2(mask, m, n, dither, f):
    k, t = _mask(mask).copy(), density_table(m, f)
    d = np.array(dither, dtype=np.int64)
    lib().orc_dither_phase2(k.ctypes.data_as(c_ubyte_p), C.c_long(m), C.c_long(n), d.ctypes.data_as(c_long_p), _dp(t))
    return d, k.astype(bool)


def dither_phase3(mask, m, n, dither, f):
    k, t = _mask(mask), density_table(m, f)
    d = np.array(dither, dtype=np.int64)
    lib().orc_dither_phase3(k.ctypes.data_as(c_ubyte_p), C.c_long(m), C.c_long(n), d.ctypes.data_as(c_long_p), _dp(t))
    return d


def blue_noise(m, picks, sigma=None, table=None):
    """bluenoise.clj:175-185 with the seed picks given; returns the dither array (int64[m*m])"""
    t = np.ascontiguousarray(table, dtype=np.float64) if table is not None else density_table(m, density_function(sigma))
    p = np.ascontiguousarray(picks, dtype=np.int64)
    dither = np.zeros(m * m, dtype=np.int64)
    lib().orc_blue_noise(C.c_long(m), C.c_long(len(p)), p.ctypes.data_as(c_long_p), _dp(t), dither.ctypes.data_as(c_long_p))
    return dither
