"""ctypes binding of oracle/cubemap_oracle.c (TEST INFRASTRUCTURE ONLY): CPU restatement of sfsim.cubemap and of
one tile of sfsim.globe/make-cube-map.  Faces are the integers 0..5 (::face0 .. ::face5, util.clj index->face)."""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "liboracle_cubemap.so")
_lib = None
PI = 3.141592653589793


class World(C.Structure):
    _fields_ = [("elevation", C.c_void_p * 8), ("day", C.c_void_p * 8), ("night", C.c_void_p * 8), ("width", C.c_long)]


def lib():
    global _lib
    if _lib is None:
        src = os.path.join(_HERE, "cubemap_oracle.c")
        if os.path.exists(src) and (not os.path.exists(_LIB_PATH) or os.path.getmtime(_LIB_PATH) < os.path.getmtime(src)):
            subprocess.check_call(["make", "-C", _HERE, "-B", "liboracle_cubemap.so"], stdout=subprocess.DEVNULL)
        _lib = C.CDLL(_LIB_PATH)
        for name in ("cube_map_x", "cube_map_y", "cube_map_z"):
            f = getattr(_lib, "orc_cm_" + name)
            f.restype = C.c_double
            f.argtypes = [C.c_int, C.c_double, C.c_double]
        for name in ("cube_i", "cube_j", "longitude", "latitude"):
            getattr(_lib, "orc_cm_" + name).restype = C.c_double
        _lib.orc_cm_cube_coordinate.restype = C.c_double
        _lib.orc_cm_cube_coordinate.argtypes = [C.c_long, C.c_long, C.c_long, C.c_double]
        _lib.orc_cm_map_x.restype = C.c_double
        _lib.orc_cm_map_x.argtypes = [C.c_double, C.c_long, C.c_long]
        _lib.orc_cm_map_y.restype = C.c_double
        _lib.orc_cm_map_y.argtypes = [C.c_double, C.c_long, C.c_long]
        _lib.orc_cm_interpolate4.restype = C.c_double
        _lib.orc_cm_interpolate4.argtypes = [C.c_double] * 4 + [C.c_void_p] * 2
        _lib.orc_cm_elevation_geodetic.restype = C.c_double
        _lib.orc_cm_elevation_geodetic.argtypes = [C.c_void_p, C.c_long, C.c_double, C.c_double]
        _lib.orc_cm_elevation_pixel.restype = C.c_long
        _lib.orc_cm_water_from_height.restype = C.c_long
        _lib.orc_cm_water_from_height.argtypes = [C.c_double]
        _lib.orc_cm_water_geodetic.restype = C.c_long
        _lib.orc_cm_water_geodetic.argtypes = [C.c_void_p, C.c_long, C.c_double, C.c_double]
        _lib.orc_cm_normal_byte.restype = C.c_int8
        _lib.orc_cm_normal_byte.argtypes = [C.c_float]
    return _lib


def _v3(v):
    return (C.c_double * 3)(*[float(x) for x in v])


def _out3():
    return (C.c_double * 3)()


def cube_map_x(face, j, i):
    return lib().orc_cm_cube_map_x(face, j, i)


def cube_map_y(face, j, i):
    return lib().orc_cm_cube_map_y(face, j, i)


def cube_map_z(face, j, i):
    return lib().orc_cm_cube_map_z(face, j, i)


def cube_map(face, j, i):
    o = _out3()
    lib().orc_cm_cube_map(C.c_int(face), C.c_double(j), C.c_double(i), o)
    return np.array(o[:])


def determine_face(p):
    return lib().orc_cm_determine_face(_v3(p))


def cube_i(face, p):
    return lib().orc_cm_cube_i(C.c_int(face), _v3(p))


def cube_j(face, p):
    return lib().orc_cm_cube_j(C.c_int(face), _v3(p))


def cube_coordinate(level, tilesize, tile, pixel):
    return lib().orc_cm_cube_coordinate(level, tilesize, tile, pixel)


def cube_map_corners(face, level, row, column):
    o = ((C.c_double * 3) * 4)()
    lib().orc_cm_cube_map_corners(C.c_int(face), C.c_long(level), C.c_long(row), C.c_long(column), o)
    return np.array([list(r) for r in o])


def longitude(p):
    return lib().orc_cm_longitude(_v3(p))


def latitude(p):
    return lib().orc_cm_latitude(_v3(p))


def geodetic_to_cartesian(lon, lat, height, radius):
    o = _out3()
    lib().orc_cm_geodetic_to_cartesian(C.c_double(lon), C.c_double(lat), C.c_double(height), C.c_double(radius), o)
    return np.array(o[:])


def cartesian_to_geodetic(p, radius):
    o = _out3()
    lib().orc_cm_cartesian_to_geodetic(_v3(p), C.c_double(radius), o)
    return np.array(o[:])


def project_onto_sphere(p, radius):
    o = _out3()
    lib().orc_cm_project_onto_sphere(_v3(p), C.c_double(radius), o)
    return np.array(o[:])


def project_onto_cube(p):
    o = _out3()
    lib().orc_cm_project_onto_cube(_v3(p), o)
    return np.array(o[:])


def map_x(lon, tilesize, level):
    return lib().orc_cm_map_x(lon, tilesize, level)


def map_y(lat, tilesize, level):
    return lib().orc_cm_map_y(lat, tilesize, level)


def _map_pixels(fn, angle, tilesize, level):
    idx = (C.c_long * 2)()
    frac = (C.c_double * 2)()
    fn(C.c_double(angle), C.c_long(tilesize), C.c_long(level), idx, frac)
    return [idx[0], idx[1], frac[0], frac[1]]


def map_pixels_x(lon, tilesize, level):
    return _map_pixels(lib().orc_cm_map_pixels_x, lon, tilesize, level)


def map_pixels_y(lat, tilesize, level):
    return _map_pixels(lib().orc_cm_map_pixels_y, lat, tilesize, level)


def offset_longitude(p, level, tilesize):
    o = _out3()
    lib().orc_cm_offset_longitude(_v3(p), C.c_long(level), C.c_long(tilesize), o)
    return np.array(o[:])


def offset_latitude(p, level, tilesize):
    o = _out3()
    lib().orc_cm_offset_latitude(_v3(p), C.c_long(level), C.c_long(tilesize), o)
    return np.array(o[:])


def interpolate4(v, xfrac, yfrac):
    xf = (C.c_double * 2)(*xfrac)
    yf = (C.c_double * 2)(*yfrac)
    return lib().orc_cm_interpolate4(v[0], v[1], v[2], v[3], C.addressof(xf), C.addressof(yf))


def tile_center(face, level, row, column, radius):
    o = _out3()
    lib().orc_cm_tile_center(C.c_int(face), C.c_long(level), C.c_long(row), C.c_long(column), C.c_double(radius), o)
    return np.array(o[:])


def water_from_height(h):
    return lib().orc_cm_water_from_height(h)


def normal_byte(x):
    return lib().orc_cm_normal_byte(x)


def surrounding_offsets(p, d1, d2):
    o = ((C.c_double * 3) * 9)()
    lib().orc_cm_surrounding_offsets(_v3(p), _v3(d1), _v3(d2), o)
    return np.array([list(r) for r in o])


def normal_from_points(pc):
    pts = np.ascontiguousarray(pc, dtype=np.float64).reshape(9, 3)
    o = _out3()
    lib().orc_cm_normal_from_points(pts.ctypes.data_as(C.c_void_p), o)
    return np.array(o[:])


class OracleWorld:
    """The world rasters, tile-major per level: elevation {level: int16[2n][4n][w][w]}, day/night {level: uint8[..][4]}."""

    def __init__(self, width, elevation=None, day=None, night=None):
        self.width = int(width)
        self._keep = []
        self.c = World()
        self.c.width = self.width
        for field, tiles, dtype, tail in (("elevation", elevation, np.int16, ()), ("day", day, np.uint8, (4,)),
                                          ("night", night, np.uint8, (4,))):
            for level, arr in (tiles or {}).items():
                n = 1 << level
                a = np.ascontiguousarray(arr, dtype=dtype)
                assert a.shape == (2 * n, 4 * n, self.width, self.width) + tail, a.shape
                self._keep.append(a)
                getattr(self.c, field)[level] = a.ctypes.data
        self.ref = C.byref(self.c)

    def elevation_pixel(self, dy, dx, level):
        return lib().orc_cm_elevation_pixel(self.ref, C.c_long(dy), C.c_long(dx), C.c_long(level))

    def world_map_pixel(self, night, dy, dx, level):
        o = _out3()
        lib().orc_cm_world_map_pixel(self.ref, C.c_int(night), C.c_long(dy), C.c_long(dx), C.c_long(level), o)
        return np.array(o[:])

    def color_geodetic(self, night, level, lon, lat):
        o = _out3()
        lib().orc_cm_color_geodetic(self.ref, C.c_int(night), C.c_long(level), C.c_double(lon), C.c_double(lat), o)
        return np.array(o[:])

    def elevation_geodetic(self, level, lon, lat):
        return lib().orc_cm_elevation_geodetic(C.cast(self.ref, C.c_void_p), level, lon, lat)

    def water_geodetic(self, level, lon, lat):
        return lib().orc_cm_water_geodetic(C.cast(self.ref, C.c_void_p), level, lon, lat)

    def project_onto_globe(self, p, level, radius):
        o = _out3()
        lib().orc_cm_project_onto_globe(self.ref, _v3(p), C.c_long(level), C.c_double(radius), o)
        return np.array(o[:])

    def surrounding_points(self, p, in_level, out_level, tilesize, radius):
        o = ((C.c_double * 3) * 9)()
        lib().orc_cm_surrounding_points(self.ref, _v3(p), C.c_long(in_level), C.c_long(out_level), C.c_long(tilesize),
                                        C.c_double(radius), o)
        return np.array([list(r) for r in o])

    def normal_for_point(self, p, in_level, out_level, tilesize, radius):
        o = _out3()
        lib().orc_cm_normal_for_point(self.ref, _v3(p), C.c_long(in_level), C.c_long(out_level), C.c_long(tilesize),
                                      C.c_double(radius), o)
        return np.array(o[:])

    def make_cube_map_tile(self, face, in_level, out_level, b, a, surface_tilesize=65, radius=6378000.0,
                           max_surface_level=4, max_color_level=5):
        """globe.clj:29-80 for one tile: dict of day, night [ct][ct][4] uint8, water [ct][align4(ct)] uint8,
        surface [st][st][3] float32, normals [ct][ct][3] float32, raw [ct][ct][7] float64 (the
        colours and the water value before truncation)"""
        st = int(surface_tilesize)
        ct = 2 * (st - 1) + 1
        pitch = (ct + 3) & ~3
        out = {"day": np.zeros((ct, ct, 4), np.uint8), "night": np.zeros((ct, ct, 4), np.uint8),
               "water": np.zeros((ct, pitch), np.uint8), "surface": np.zeros((st, st, 3), np.float32),
               "normals": np.zeros((ct, ct, 3), np.float32), "raw": np.zeros((ct, ct, 7), np.float64)}
        lib().orc_cm_make_cube_map_tile(self.ref, C.c_int(face), C.c_long(in_level), C.c_long(out_level), C.c_long(b),
                                        C.c_long(a), C.c_long(st), C.c_long(max_surface_level),
                                        C.c_long(max_color_level), C.c_double(radius),
                                        *[out[k].ctypes.data_as(C.c_void_p) for k in ("day", "night", "water", "surface",
                                                                                      "normals", "raw")])
        return out


from sfsim_b200.synthetic import synthetic_world  # noqa: E402,F401  (numpy-only raster generator shared with bench.py)
