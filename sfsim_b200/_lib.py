"""ctypes binding of libsfsim_atmosphere.so (include/sfsim_atmosphere.h).

This is the same set of symbols the Clojure shim binds with coffi.  There is no
CPU fallback: if the library is missing or no CUDA device is present, calls raise.
"""
import ctypes as C
import os

import numpy as np

_ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB_NAME = "libsfsim_atmosphere.so"
LIB_PATH = os.environ.get("SFSIM_ATMOSPHERE_LIB", os.path.join(_ROOT, LIB_NAME))

c_double_p = C.POINTER(C.c_double)
c_float_p = C.POINTER(C.c_float)
c_int_p = C.POINTER(C.c_int)


class AtmlutError(RuntimeError):
    """Raised when a library entry point returns a non-zero status (the Clojure shim throws likewise)."""


class Planet(C.Structure):
    _fields_ = [("centre", C.c_double * 3), ("radius", C.c_double), ("height", C.c_double),
                ("brightness", C.c_double * 3)]


class Scatter(C.Structure):
    _fields_ = [("base", C.c_double * 3), ("scale", C.c_double), ("g", C.c_double), ("quotient", C.c_double)]


class Config(C.Structure):
    _fields_ = [("height_size", C.c_int), ("elevation_size", C.c_int), ("light_elevation_size", C.c_int),
                ("heading_size", C.c_int), ("transmittance_height_size", C.c_int),
                ("transmittance_elevation_size", C.c_int), ("surface_height_size", C.c_int),
                ("surface_sun_elevation_size", C.c_int), ("ray_steps", C.c_int), ("sphere_steps", C.c_int),
                ("iterations", C.c_int), ("intensity", C.c_double * 3)]

    @property
    def ray_scatter_shape(self):
        return (self.height_size, self.elevation_size, self.light_elevation_size, self.heading_size)

    @property
    def transmittance_shape(self):
        return (self.transmittance_height_size, self.transmittance_elevation_size)

    @property
    def surface_radiance_shape(self):
        return (self.surface_height_size, self.surface_sun_elevation_size)


OPT_GRAPH, OPT_BARRIER_TIMEOUT_MS = 1, 2

ALLGATHER_FN = C.CFUNCTYPE(C.c_int, C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p)

# every symbol include/sfsim_atmosphere.h, include/sfsim_noise.h and include/sfsim_cubemap.h declare
EXPORTS = [
    "atmlut_init", "atmlut_destroy", "atmlut_stream", "atmlut_last_error", "atmlut_device_count", "atmlut_default_config",
    "atmlut_generate", "atmlut_generate_multi",
    "atmlut_builder_create", "atmlut_sphere_directions", "atmlut_slab", "atmlut_builder_ipc_handle_bytes",
    "atmlut_builder_ipc_export", "atmlut_builder_ipc_import",
    "atmlut_builder_set_allgather", "atmlut_builder_set_option", "atmlut_builder_run", "atmlut_builder_run_timed",
    "atmlut_builder_stream", "atmlut_builder_sync", "atmlut_host_alloc", "atmlut_host_free",
    "atmlut_builder_download", "atmlut_builder_stage_count", "atmlut_builder_stage_name", "atmlut_builder_stage_ms",
    "atmlut_builder_work", "atmlut_builder_counter", "atmlut_builder_destroy",
    "atmlut_transmittance_table", "atmlut_surface_radiance_base_table", "atmlut_first_order_tables",
    "atmlut_point_scatter_table", "atmlut_surface_radiance_table", "atmlut_ray_scatter_table",
    "atmlut_resample_table",
    "atmlut_transmittance_batch", "atmlut_transmittance_dir_batch", "atmlut_surface_radiance_base_batch",
    "atmlut_point_scatter_first_order_batch", "atmlut_ray_scatter_first_order_batch",
    "atmlut_ray_scatter_table_batch", "atmlut_point_scatter_batch", "atmlut_surface_radiance_batch",
    "atmlut_index_forward_batch", "atmlut_index_backward_batch", "atmlut_index_map_batch",
    "atmlut_interpolate_batch", "atmlut_medium_batch",
    "atmlut_convert_4d_to_2d", "atmlut_write_floats", "atmlut_read_floats",
    # include/sfsim_noise.h
    "sfsim_worley_noise", "sfsim_perlin_noise", "sfsim_worley_distances", "sfsim_perlin_samples",
    "sfsim_blue_noise", "sfsim_blue_noise_texture",
    # include/sfsim_cubemap.h
    "sfsim_cubemap_default_config", "sfsim_cubemap_world_create", "sfsim_cubemap_world_destroy",
    "sfsim_cubemap_world_set_elevation", "sfsim_cubemap_world_set_color", "sfsim_cubemap_world_set_elevation_tile",
    "sfsim_cubemap_world_set_color_tile", "sfsim_cubemap_tiles", "sfsim_cubemap_tiles_timed", "sfsim_cubemap_tile_shard",
    "sfsim_cubemap_level", "sfsim_cubemap_tile_counter",
    "sfsim_cubemap_project_onto_globe_batch", "sfsim_cubemap_normal_for_point_batch", "sfsim_cubemap_geodetic_batch",
]

_lib = None


def load():
    """Load the shared library; raises if it has not been built (no fallback)."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise AtmlutError("%s not found: build it with `make` in %s (there is no CPU fallback)" %
                              (LIB_NAME, _ROOT))
        lib = C.CDLL(LIB_PATH)
        lib.atmlut_last_error.restype = C.c_char_p
        lib.atmlut_builder_stage_name.restype = C.c_char_p
        lib.atmlut_builder_stage_name.argtypes = [C.c_void_p, C.c_int]
        lib.atmlut_read_floats.restype = C.c_long
        lib.atmlut_stream.restype = C.c_void_p
        lib.atmlut_builder_create.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_int,
                                              C.POINTER(C.c_void_p)]
        for name in ("atmlut_builder_set_allgather",):
            getattr(lib, name).argtypes = [C.c_void_p, ALLGATHER_FN, C.c_void_p]
        for name in ("atmlut_builder_run", "atmlut_builder_run_timed", "atmlut_builder_sync", "atmlut_builder_destroy",
                     "atmlut_builder_stage_count", "atmlut_builder_stream"):
            getattr(lib, name).argtypes = [C.c_void_p]
        lib.atmlut_builder_stream.restype = C.c_void_p
        lib.atmlut_builder_set_option.argtypes = [C.c_void_p, C.c_int, C.c_int]
        lib.atmlut_host_alloc.restype = C.c_void_p
        lib.atmlut_host_alloc.argtypes = [C.c_size_t]
        lib.atmlut_host_free.argtypes = [C.c_void_p]
        lib.atmlut_builder_download.argtypes = [C.c_void_p] * 5
        lib.atmlut_builder_ipc_export.argtypes = [C.c_void_p, C.c_char_p, C.c_int]
        lib.atmlut_builder_ipc_import.argtypes = [C.c_void_p, C.c_char_p, C.c_int]
        lib.atmlut_builder_stage_ms.argtypes = [C.c_void_p, C.c_int, c_float_p]
        lib.atmlut_builder_work.argtypes = [C.c_void_p, c_double_p, c_double_p, c_double_p]
        lib.atmlut_builder_counter.argtypes = [C.c_void_p, C.c_int, c_double_p]
        lib.sfsim_cubemap_default_config.restype = None
        lib.sfsim_cubemap_world_destroy.restype = None
        lib.sfsim_cubemap_world_destroy.argtypes = [C.c_void_p]
        _lib = lib
    return _lib


def check(status):
    if status != 0:
        raise AtmlutError(load().atmlut_last_error().decode())


def vec3(v):
    return (C.c_double * 3)(float(v[0]), float(v[1]), float(v[2]))


def make_planet(planet):
    """planet: dict with radius, height and optional centre, brightness (the reference's planet map)."""
    if isinstance(planet, Planet):
        return planet
    return Planet(vec3(planet.get("centre", (0.0, 0.0, 0.0))), float(planet["radius"]), float(planet["height"]),
                  vec3(planet.get("brightness", (0.0, 0.0, 0.0))))


def make_scatter(s):
    """s: dict with base, scale and optional g, quotient (the reference's scatter map)."""
    if isinstance(s, Scatter):
        return s
    return Scatter(vec3(s["base"]), float(s["scale"]), float(s.get("g", 0.0)), float(s.get("quotient", 1.0)))


def make_scatter_array(scatters):
    arr = (Scatter * max(1, len(scatters)))()
    for i, s in enumerate(scatters):
        arr[i] = make_scatter(s)
    return arr


def default_config():
    cfg = Config()
    load().atmlut_default_config(C.byref(cfg))
    return cfg


def make_config(**kw):
    """Shipped constants (atmosphere_lut.clj:47-63) overridden by keyword."""
    cfg = default_config()
    for k, v in kw.items():
        if k == "intensity":
            cfg.intensity = vec3(v)
        elif k == "ray_scatter_shape":
            cfg.height_size, cfg.elevation_size, cfg.light_elevation_size, cfg.heading_size = [int(x) for x in v]
        elif k == "transmittance_shape":
            cfg.transmittance_height_size, cfg.transmittance_elevation_size = [int(x) for x in v]
        elif k == "surface_radiance_shape":
            cfg.surface_height_size, cfg.surface_sun_elevation_size = [int(x) for x in v]
        else:
            if not hasattr(cfg, k):
                raise TypeError("unknown config field %r" % k)
            setattr(cfg, k, int(v))
    return cfg


def f32(a):
    return np.ascontiguousarray(a, dtype=np.float32)


def f64(a):
    return np.ascontiguousarray(a, dtype=np.float64)


def i32(a):
    return np.ascontiguousarray(a, dtype=np.int32)


def ptr(a, ctype=None):
    if a is None:
        return None
    if ctype is None:
        ctype = {np.dtype(np.float32): C.c_float, np.dtype(np.float64): C.c_double,
                 np.dtype(np.int32): C.c_int}[a.dtype]
    return a.ctypes.data_as(C.POINTER(ctype))
