"""Host-side mirror of sfsim.cubemap / sfsim.globe (src/clj/sfsim/cubemap.clj, globe.clj) over libsfsim_atmosphere.so
(include/sfsim_cubemap.h).

The coordinate functions that need no raster (`cube_map`, `cube_coordinate`, `map_pixels_x`, ...) are plain double
arithmetic on the host, written operation by operation like the reference.  Everything that reads the world rasters --
`project_onto_globe`, `normal_for_point`, `elevation_geodetic`, `water_geodetic`, `color_geodetic_day/night` and the tile
loop of `make_cube_map` (globe.clj:29-80) -- runs on the GPU against rasters kept in device memory (`World`).
There is no CPU fallback.  Faces are 0..5 (::face0 .. ::face5).
"""
import ctypes as C
import math

import numpy as np

from . import _lib
from ._lib import check

PI = math.pi


class CubemapConfig(C.Structure):
    _fields_ = [("in_level", C.c_int), ("out_level", C.c_int), ("width", C.c_int), ("surface_tilesize", C.c_int),
                ("sublevel", C.c_int), ("max_surface_level", C.c_int), ("max_color_level", C.c_int),
                ("radius", C.c_double)]

    @property
    def color_tilesize(self):      # globe.clj:38
        return (1 << self.sublevel) * (self.surface_tilesize - 1) + 1


ALL_OUTPUTS = ("day", "night", "water", "surface", "normals", "normal_bytes")     # bit k of the C mask = entry k


def _mask(outputs):
    unknown = set(outputs) - set(ALL_OUTPUTS)
    if unknown:
        raise TypeError("unknown outputs %s" % sorted(unknown))
    return sum(1 << ALL_OUTPUTS.index(name) for name in set(outputs))


TILE_FN = C.CFUNCTYPE(C.c_int, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                      C.c_void_p)


def make_config(in_level, out_level, **kw):
    """The constants of make-cube-map (globe.clj:32-40) overridden by keyword."""
    cfg = CubemapConfig()
    _lib.load().sfsim_cubemap_default_config(C.byref(cfg))
    cfg.in_level, cfg.out_level = int(in_level), int(out_level)
    for k, v in kw.items():
        if not hasattr(cfg, k):
            raise TypeError("unknown config field %r" % k)
        setattr(cfg, k, float(v) if k == "radius" else int(v))
    return cfg


# ---------------------------------------------------------------- cubemap.clj:29-168 (host arithmetic)

def cube_map_x(face, j, i):
    return (-1.0 + 2.0 * i, -1.0 + 2.0 * i, 1.0, 1.0 - 2.0 * i, -1.0, -1.0 + 2.0 * i)[face]


def cube_map_y(face, j, i):
    return (1.0 - 2.0 * j, -1.0, -1.0 + 2.0 * i, 1.0, 1.0 - 2.0 * i, -1.0 + 2.0 * j)[face]


def cube_map_z(face, j, i):
    return (1.0, 1.0 - 2.0 * j, 1.0 - 2.0 * j, 1.0 - 2.0 * j, 1.0 - 2.0 * j, -1.0)[face]


def cube_map(face, j, i):
    """cubemap.clj:62-67"""
    return np.array([cube_map_x(face, j, i), cube_map_y(face, j, i), cube_map_z(face, j, i)])


def determine_face(point):
    """cubemap.clj:70-78"""
    x, y, z = [float(c) for c in point]
    if abs(x) >= max(abs(y), abs(z)):
        return 2 if x >= 0 else 4
    if abs(y) >= max(abs(x), abs(z)):
        return 3 if y >= 0 else 1
    return 0 if z >= 0 else 5


def cube_i(face, p):
    """cubemap.clj:81-91"""
    return (0.5 * (p[0] + 1.0), 0.5 * (p[0] + 1.0), 0.5 * (p[1] + 1.0), 0.5 * (1.0 - p[0]), 0.5 * (1.0 - p[1]),
            0.5 * (p[0] + 1.0))[face]


def cube_j(face, p):
    """cubemap.clj:94-104"""
    return (0.5 * (1.0 - p[1]), 0.5 * (1.0 - p[2]), 0.5 * (1.0 - p[2]), 0.5 * (1.0 - p[2]), 0.5 * (1.0 - p[2]),
            0.5 * (p[1] + 1.0))[face]


def cube_coordinate(level, tilesize, tile, pixel):
    """cubemap.clj:107-111"""
    return (tile + float(pixel) / (tilesize - 1)) / (1 << level)


def cube_map_corners(face, level, row, column):
    """cubemap.clj:114-120"""
    return [cube_map(face, cube_coordinate(level, 2, row, float(dj)), cube_coordinate(level, 2, column, float(di)))
            for dj in (0, 1) for di in (0, 1)]


def longitude(p):
    return math.atan2(p[1], p[0])


def latitude(p):
    return math.atan2(p[2], math.sqrt(p[0] * p[0] + p[1] * p[1]))


def geodetic_to_cartesian(lon, lat, height, radius):
    """cubemap.clj:137-143"""
    distance = height + radius
    cos_lat, sin_lat = math.cos(lat), math.sin(lat)
    return np.array([distance * cos_lat * math.cos(lon), distance * cos_lat * math.sin(lon), distance * sin_lat])


def project_onto_sphere(p, radius):
    """cubemap.clj:146-149"""
    m = math.sqrt(p[0] * p[0] + p[1] * p[1] + p[2] * p[2])
    return np.array([p[0] / m * radius, p[1] / m * radius, p[2] / m * radius])


def project_onto_cube(p):
    """cubemap.clj:152-160"""
    ax, ay, az = abs(p[0]), abs(p[1]), abs(p[2])
    d = ax if ax >= max(ay, az) else (ay if ay >= max(ax, az) else az)
    return np.array([p[0] / d, p[1] / d, p[2] / d])


def cartesian_to_geodetic(p, radius):
    """cubemap.clj:163-168: [longitude latitude height]"""
    height = math.sqrt(p[0] * p[0] + p[1] * p[1] + p[2] * p[2]) - radius
    return [math.atan2(p[1], p[0]), math.atan2(p[2], math.sqrt(p[0] * p[0] + p[1] * p[1])), height]


def map_x(lon, tilesize, level):
    """cubemap.clj:171-175"""
    return (PI + lon) * ((4 * (1 << level) * tilesize) / (2 * PI))


def map_y(lat, tilesize, level):
    """cubemap.clj:178-182"""
    return (PI / 2 - lat) * ((2 * (1 << level) * tilesize) / PI)


def map_pixels_x(lon, tilesize, level):
    """cubemap.clj:185-195"""
    size = 4 * (1 << level) * tilesize
    x = map_x(lon, tilesize, level)
    x0 = int(math.floor(x))
    frac1 = x - x0
    return [x0 % size, (x0 + 1) % size, 1 - frac1, frac1]


def map_pixels_y(lat, tilesize, level):
    """cubemap.clj:198-207"""
    size = 2 * (1 << level) * tilesize
    y = map_y(lat, tilesize, level)
    y0 = int(math.floor(y))
    frac1 = y - y0
    return [min(y0, size - 1), min(y0 + 1, size - 1), 1 - frac1, frac1]


def tile_center(face, level, row, column, radius):
    """cubemap.clj:302-308"""
    return project_onto_sphere(cube_map(face, cube_coordinate(level, 3, row, 1.0), cube_coordinate(level, 3, column, 1.0)),
                               radius)


def level_shape(width, level):
    """(tile rows, tile columns, width, width) of a map level: 2n x 4n tiles (cubemap.clj:185-207)"""
    n = 1 << level
    return (2 * n, 4 * n, width, width)


def tile_shard(out_level, rank=0, world_size=1):
    """The (face, b, a) triples rank `rank` of `world_size` generates: every world_size-th tile in the order of
    globe.clj:41.  Tiles are independent, so several GPUs need no exchange."""
    n = 1 << out_level
    count = len(range(rank, 6 * n * n, world_size))
    tiles = np.zeros((max(count, 1), 3), dtype=np.int32)
    got = C.c_int(0)
    check(_lib.load().sfsim_cubemap_tile_shard(int(out_level), int(rank), int(world_size), count, _lib.ptr(tiles),
                                               C.byref(got)))
    return tiles[:got.value]


# ---------------------------------------------------------------- rasters on the device

class World:
    """The Mercator map tiles of one or more levels in device memory (world-map-tile / elevation-tile,
    cubemap.clj:232-248, without the LRU cache: every level stays resident)."""

    def __init__(self, width=675):
        self.width = int(width)
        self._h = C.c_void_p()
        check(_lib.load().sfsim_cubemap_world_create(self.width, C.byref(self._h)))

    def close(self):
        if self._h:
            _lib.load().sfsim_cubemap_world_destroy(self._h)
            self._h = C.c_void_p()

    __del__ = close

    def _level(self, arr, level, dtype, tail):
        a = np.ascontiguousarray(arr, dtype=dtype)
        want = level_shape(self.width, level) + tail
        if a.shape != want:
            raise TypeError("level %d needs shape %s, got %s" % (level, want, a.shape))
        return a

    def set_elevation(self, level, tiles):
        """tiles: int16 [2n][4n][width][width], tile (ty, tx) = tmp/elevation/<level>/<tx>/<ty>.raw"""
        a = self._level(tiles, level, np.int16, ())
        check(_lib.load().sfsim_cubemap_world_set_elevation(self._h, int(level), a.ctypes.data_as(C.c_void_p)))

    def set_color(self, night, level, tiles):
        """tiles: uint8 [2n][4n][width][width][4] (RGBA as slurp-image returns it)"""
        a = self._level(tiles, level, np.uint8, (4,))
        check(_lib.load().sfsim_cubemap_world_set_color(self._h, int(bool(night)), int(level), a.ctypes.data_as(C.c_void_p)))

    def set_elevation_tile(self, level, ty, tx, tile):
        a = np.ascontiguousarray(tile, dtype=np.int16)
        if a.shape != (self.width, self.width):
            raise TypeError("an elevation tile has shape (%d, %d)" % (self.width, self.width))
        check(_lib.load().sfsim_cubemap_world_set_elevation_tile(self._h, int(level), int(ty), int(tx),
                                                                 a.ctypes.data_as(C.c_void_p)))

    def set_color_tile(self, night, level, ty, tx, rgba):
        a = np.ascontiguousarray(rgba, dtype=np.uint8)
        if a.shape != (self.width, self.width, 4):
            raise TypeError("a colour tile has shape (%d, %d, 4)" % (self.width, self.width))
        check(_lib.load().sfsim_cubemap_world_set_color_tile(self._h, int(bool(night)), int(level), int(ty), int(tx),
                                                             a.ctypes.data_as(C.c_void_p)))

    # ------------------------------------------------------------ point-wise functions (batches)

    def _points(self, p):
        a = _lib.f64(p).reshape(-1, 3)
        return a, np.zeros_like(a)

    def project_onto_globe(self, p, in_level, radius=6378000.0):
        """cubemap.clj:336-342 for points p[n][3]"""
        a, out = self._points(p)
        check(_lib.load().sfsim_cubemap_project_onto_globe_batch(self._h, int(in_level), C.c_double(radius), len(a),
                                                                 _lib.ptr(a), _lib.ptr(out)))
        return out

    def normal_for_point(self, p, in_level, out_level, tilesize, radius=6378000.0):
        """cubemap.clj:357-366 for points p[n][3]"""
        a, out = self._points(p)
        check(_lib.load().sfsim_cubemap_normal_for_point_batch(self._h, int(in_level), int(out_level), int(tilesize),
                                                               C.c_double(radius), len(a), _lib.ptr(a), _lib.ptr(out)))
        return out

    def _geodetic(self, kind, in_level, lon, lat):
        lon, lat = _lib.f64(lon).reshape(-1), _lib.f64(lat).reshape(-1)
        if lon.shape != lat.shape:
            raise TypeError("lon and lat must have the same length")
        out = np.zeros((len(lon), 3) if kind >= 2 else (len(lon),))
        check(_lib.load().sfsim_cubemap_geodetic_batch(self._h, kind, int(in_level), len(lon), _lib.ptr(lon), _lib.ptr(lat),
                                                       _lib.ptr(out)))
        return out

    def elevation_geodetic(self, in_level, lon, lat):
        """cubemap.clj:323-326"""
        return self._geodetic(0, in_level, lon, lat)

    def water_geodetic(self, in_level, lon, lat):
        """cubemap.clj:329-333"""
        return self._geodetic(1, in_level, lon, lat).astype(np.int64)

    def color_geodetic_day(self, in_level, lon, lat):
        """cubemap.clj:311-314"""
        return self._geodetic(2, in_level, lon, lat)

    def color_geodetic_night(self, in_level, lon, lat):
        """cubemap.clj:317-320"""
        return self._geodetic(3, in_level, lon, lat)

    # ------------------------------------------------------------ tiles

    def make_cube_map_tiles(self, cfg, tiles, outputs=("day", "night", "water", "surface", "normals", "normal_bytes")):
        """globe.clj:41-72 for the (face, b, a) triples in `tiles`: dict of arrays with a leading tile axis --
        day, night uint8 [ct][ct][4]; water uint8 [ct][align4(ct)]; surface float32 [st][st][3]; normals float32
        [ct][ct][3]; normal_bytes int8 [ct][ct][3] (what spit-normals encodes)."""
        tiles = _lib.i32(tiles).reshape(-1, 3)
        n, st, ct = len(tiles), cfg.surface_tilesize, cfg.color_tilesize
        pitch = (ct + 3) & ~3
        shapes = {"day": ((n, ct, ct, 4), np.uint8), "night": ((n, ct, ct, 4), np.uint8), "water": ((n, ct, pitch), np.uint8),
                  "surface": ((n, st, st, 3), np.float32), "normals": ((n, ct, ct, 3), np.float32),
                  "normal_bytes": ((n, ct, ct, 3), np.int8)}
        out = {k: np.zeros(*shapes[k]) for k in outputs}
        args = [out[k].ctypes.data_as(C.c_void_p) if k in out else None
                for k in ("day", "night", "water", "surface", "normals", "normal_bytes")]
        check(_lib.load().sfsim_cubemap_tiles(self._h, C.byref(cfg), n, _lib.ptr(tiles), *args))
        return out

    def time_cube_map_tiles(self, cfg, tiles):
        """device time in ms of one batch with all six outputs (no download)"""
        tiles = _lib.i32(tiles).reshape(-1, 3)
        ms = C.c_float(0)
        check(_lib.load().sfsim_cubemap_tiles_timed(self._h, C.byref(cfg), len(tiles), _lib.ptr(tiles), C.byref(ms)))
        return ms.value

    def time_cube_map_level(self, cfg, rank=0, world_size=1, batch=256, outputs=ALL_OUTPUTS):
        """seconds for sfsim_cubemap_level with the library's counting callback (no consumer): (seconds, tiles)"""
        import time
        acc = (C.c_longlong * 2)(0, 0)
        lib = _lib.load()
        fn = C.cast(lib.sfsim_cubemap_tile_counter, TILE_FN)
        t0 = time.perf_counter()
        check(lib.sfsim_cubemap_level(self._h, C.byref(cfg), int(rank), int(world_size), int(batch), _mask(outputs), fn,
                                      C.byref(acc)))
        return time.perf_counter() - t0, int(acc[0])

    def make_cube_map(self, in_level, out_level, on_tile, rank=0, world_size=1, batch=256, outputs=ALL_OUTPUTS, **kw):
        """make-cube-map (globe.clj:29-80) for this rank's tiles of the output level, streamed: `on_tile((face, b, a),
        tile)` is called for every tile with a dict of array views (day, night, water, surface, normals, normal_bytes)
        into page-locked memory that is valid during the call only -- the place to encode and write the tile's five files
        (spit-jpg, spit-bytes-gz, spit-floats-gz, spit-normals).  While the callback works on one batch the next one
        crosses PCIe and the one after is being computed.  `outputs`: the arrays to compute and bring back (the bus is
        the bound; leave "normals" out if `normal_bytes` is all the PNG encoder needs).  Returns the number of tiles
        delivered."""
        cfg = make_config(in_level, out_level, width=self.width, **kw)
        st, ct = cfg.surface_tilesize, cfg.color_tilesize
        pitch = (ct + 3) & ~3
        shapes = (("day", C.c_ubyte, (ct, ct, 4)), ("night", C.c_ubyte, (ct, ct, 4)), ("water", C.c_ubyte, (ct, pitch)),
                  ("surface", C.c_float, (st, st, 3)), ("normals", C.c_float, (ct, ct, 3)), ("normal_bytes", C.c_byte, (ct, ct, 3)))
        delivered = [0]
        failure = []

        def trampoline(_user, face, b, a, *ptrs):
            try:
                tile = {name: np.ctypeslib.as_array(C.cast(p, C.POINTER(ctype)), shape=shape)
                        for (name, ctype, shape), p in zip(shapes, ptrs) if p}
                on_tile((face, b, a), tile)
                delivered[0] += 1
                return 0
            except BaseException as e:      # an exception must not unwind through the C frames
                failure.append(e)
                return 1

        fn = TILE_FN(trampoline)
        status = _lib.load().sfsim_cubemap_level(self._h, C.byref(cfg), int(rank), int(world_size), int(batch),
                                                 _mask(outputs), fn, None)
        if failure:
            raise failure[0]
        check(status)
        return delivered[0]
