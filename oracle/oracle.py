"""TEST INFRASTRUCTURE ONLY -- ctypes front end of the CPU oracle (oracle/atmosphere_oracle.c).

The oracle is a double-precision restatement of the reference's Clojure algorithm
(wedesoft/sfsim src/clj/sfsim/atmosphere.clj, atmosphere_lut.clj, interpolate.clj, ray.clj,
sphere.clj).  Only tests/, __graft_entry__.smoke() and bench.py's CPU-baseline legs may import
this module; the product (sfsim_b200/) never does.

Parity pin: tests/test_oracle_known_answers.py checks it against the reference's own midje
facts (test/clj/sfsim/t_atmosphere.clj, t_sphere.clj, t_ray.clj, t_interpolate.clj, t_util.clj,
t_image.clj) including the LUT-through-GLSL goldens.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "liboracle_atmosphere.so")

c_double_p = C.POINTER(C.c_double)
c_long_p = C.POINTER(C.c_long)


class Planet(C.Structure):
    _fields_ = [("centre", C.c_double * 3), ("radius", C.c_double), ("height", C.c_double),
                ("brightness", C.c_double * 3)]


class Scatter(C.Structure):
    _fields_ = [("base", C.c_double * 3), ("scale", C.c_double), ("g", C.c_double), ("quotient", C.c_double)]


class Config(C.Structure):
    _fields_ = [("shape4", C.c_long * 4), ("shape_t", C.c_long * 2), ("shape_e", C.c_long * 2),
                ("ray_steps", C.c_long), ("sphere_steps", C.c_long), ("intensity", C.c_double * 3)]


class SSource(C.Structure):
    _fields_ = [("kind", C.c_int), ("tab_a", c_double_p), ("tab_b", c_double_p), ("phase_component", Scatter)]


POINT_FN = C.CFUNCTYPE(None, C.c_void_p, c_double_p, c_double_p, c_double_p, C.c_int, c_double_p)
SURFACE_FN = C.CFUNCTYPE(None, C.c_void_p, c_double_p, c_double_p, c_double_p)
VEC_FN = C.CFUNCTYPE(None, C.c_void_p, c_double_p, c_double_p)
ANGLE_FN = C.CFUNCTYPE(None, C.c_void_p, C.c_double, c_double_p)


def build(force=False):
    """Compile the oracle's C restatement (building the checker is not using it)."""
    src = os.path.join(_HERE, "atmosphere_oracle.c")
    hdr = os.path.join(_HERE, "atmosphere_oracle.h")
    if not force and os.path.exists(_LIB_PATH) and os.path.exists(src):
        if os.path.getmtime(_LIB_PATH) >= max(os.path.getmtime(src), os.path.getmtime(hdr)):
            return _LIB_PATH
    if os.path.exists(src):
        subprocess.check_call(["make", "-C", _HERE, "-B", "liboracle_atmosphere.so"], stdout=subprocess.DEVNULL)
    return _LIB_PATH


_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        _lib = C.CDLL(_LIB_PATH)
        _lib.orc_limit_quot.restype = C.c_double
        _lib.orc_limit_quot.argtypes = [C.c_double] * 4
        _lib.orc_limit_quot3.restype = C.c_double
        _lib.orc_limit_quot3.argtypes = [C.c_double] * 3
        _lib.orc_height.restype = C.c_double
        _lib.orc_phase.restype = C.c_double
        _lib.orc_phase.argtypes = [C.c_void_p, C.c_double]
        _lib.orc_horizon_distance.restype = C.c_double
        _lib.orc_horizon_distance.argtypes = [C.c_void_p, C.c_double]
        _lib.orc_elevation_to_index.restype = C.c_double
        _lib.orc_height_to_index.restype = C.c_double
        _lib.orc_sun_elevation_to_index.restype = C.c_double
        _lib.orc_index_to_sin_sun_elevation.restype = C.c_double
        _lib.orc_index_to_sin_sun_elevation.argtypes = [C.c_long, C.c_double]
        _lib.orc_sun_angle_to_index.restype = C.c_double
        _lib.orc_sphere_directions.restype = C.c_long
        _lib.orc_slurp_floats.restype = C.c_long
    return _lib


# ------------------------------------------------------------------ conversions

def vec3(*a):
    if len(a) == 1:
        a = a[0]
    return (C.c_double * 3)(*[float(x) for x in a])


def planet(radius, height, brightness=(0.3, 0.3, 0.3), centre=(0.0, 0.0, 0.0)):
    return Planet(vec3(centre), float(radius), float(height), vec3(brightness))


def scatter(base, scale, g=0.0, quotient=1.0):
    return Scatter(vec3(base), float(scale), float(g), float(quotient))


def scatter_array(scatters):
    arr = (Scatter * len(scatters))()
    for i, s in enumerate(scatters):
        arr[i] = s
    return arr


def config(shape4, shape_t, shape_e, ray_steps=100, sphere_steps=15, intensity=(1.0, 1.0, 1.0)):
    return Config((C.c_long * 4)(*shape4), (C.c_long * 2)(*shape_t), (C.c_long * 2)(*shape_e), int(ray_steps),
                  int(sphere_steps), vec3(intensity))


def _out3():
    return (C.c_double * 3)()


def _np3(buf):
    return np.array([buf[0], buf[1], buf[2]], dtype=np.float64)


def _shape(shape):
    return (C.c_long * len(shape))(*[int(s) for s in shape])


# ------------------------------------------------------------------ scalar/vector functions

def limit_quot(a, b, lower, upper=None):
    if upper is None:
        return lib().orc_limit_quot3(a, b, lower)
    return lib().orc_limit_quot(a, b, lower, upper)


def height(pl, p):
    return lib().orc_height(C.byref(pl), vec3(p))


def ray_sphere_intersection(centre, radius, origin, direction):
    d, l = C.c_double(), C.c_double()
    lib().orc_ray_sphere_intersection(vec3(centre), C.c_double(radius), vec3(origin), vec3(direction), C.byref(d),
                                      C.byref(l))
    return d.value, l.value


def integral_ray(origin, direction, steps, distance, fun, dims=3):
    out = (C.c_double * dims)()

    def cb(_ctx, p, o):
        val = fun(np.array([p[0], p[1], p[2]]))
        for i in range(dims):
            o[i] = float(val[i])

    lib().orc_integral_ray(vec3(origin), vec3(direction), C.c_long(steps), C.c_double(distance), VEC_FN(cb), None,
                           C.c_int(dims), out)
    return np.array(list(out))


def integrate_circle(steps, fun, dims=3):
    out = (C.c_double * dims)()

    def cb(_ctx, phi, o):
        val = fun(phi)
        for i in range(dims):
            o[i] = float(val[i])

    lib().orc_integrate_circle(C.c_long(steps), ANGLE_FN(cb), None, C.c_int(dims), out)
    return np.array(list(out))


def _sphere_integral(name, steps, normal, fun, dims):
    out = (C.c_double * dims)()

    def cb(_ctx, p, o):
        val = fun(np.array([p[0], p[1], p[2]]))
        for i in range(dims):
            o[i] = float(val[i])

    getattr(lib(), name)(C.c_long(steps), vec3(normal), VEC_FN(cb), None, C.c_int(dims), out)
    return np.array(list(out))


def integral_half_sphere(steps, normal, fun, dims=3):
    return _sphere_integral("orc_integral_half_sphere", steps, normal, fun, dims)


def integral_sphere(steps, normal, fun, dims=3):
    return _sphere_integral("orc_integral_sphere", steps, normal, fun, dims)


def sphere_directions(theta_steps, phi_steps, theta_range, normal):
    n = lib().orc_sphere_directions(C.c_long(theta_steps), C.c_long(phi_steps), C.c_double(theta_range), vec3(normal),
                                    None, None)
    dirs = np.zeros((n, 3))
    weights = np.zeros(n)
    lib().orc_sphere_directions(C.c_long(theta_steps), C.c_long(phi_steps), C.c_double(theta_range), vec3(normal),
                                dirs.ctypes.data_as(c_double_p), weights.ctypes.data_as(c_double_p))
    return dirs, weights


def orthogonal(n):
    out = _out3()
    lib().orc_orthogonal(vec3(n), out)
    return _np3(out)


def oriented_matrix(n):
    out = (C.c_double * 9)()
    lib().orc_oriented_matrix(vec3(n), out)
    return np.array(list(out)).reshape(3, 3)


def scattering(s, h):
    out = _out3()
    lib().orc_scattering(C.byref(s), C.c_double(h), out)
    return _np3(out)


def extinction(s, h):
    out = _out3()
    lib().orc_extinction(C.byref(s), C.c_double(h), out)
    return _np3(out)


def phase(s, mu):
    return lib().orc_phase(C.byref(s) if s is not None else None, mu)


def atmosphere_intersection(pl, origin, direction):
    out = _out3()
    lib().orc_atmosphere_intersection(C.byref(pl), vec3(origin), vec3(direction), out)
    return _np3(out)


def surface_intersection(pl, origin, direction):
    out = _out3()
    lib().orc_surface_intersection(C.byref(pl), vec3(origin), vec3(direction), out)
    return _np3(out)


def surface_point(pl, p):
    return bool(lib().orc_surface_point(C.byref(pl), vec3(p)))


def is_above_horizon(pl, p, d):
    return bool(lib().orc_is_above_horizon(C.byref(pl), vec3(p), vec3(d)))


def ray_extremity(pl, origin, direction):
    out = _out3()
    lib().orc_ray_extremity(C.byref(pl), vec3(origin), vec3(direction), out)
    return _np3(out)


def transmittance(pl, scatters, steps, x, x0_or_v, above=None):
    out = _out3()
    arr = scatter_array(scatters)
    if above is None:
        lib().orc_transmittance(C.byref(pl), arr, len(scatters), C.c_long(steps), vec3(x), vec3(x0_or_v), out)
    else:
        lib().orc_transmittance_dir(C.byref(pl), arr, len(scatters), C.c_long(steps), vec3(x), vec3(x0_or_v),
                                    C.c_int(bool(above)), out)
    return _np3(out)


def surface_radiance_base(pl, scatters, steps, intensity, x, l):
    out = _out3()
    lib().orc_surface_radiance_base(C.byref(pl), scatter_array(scatters), len(scatters), C.c_long(steps),
                                    vec3(intensity), vec3(x), vec3(l), out)
    return _np3(out)


def point_scatter_component(pl, scatters, component, steps, intensity, x, v, l, above):
    out = _out3()
    lib().orc_point_scatter_component(C.byref(pl), scatter_array(scatters), len(scatters), C.byref(component),
                                      C.c_long(steps), vec3(intensity), vec3(x), vec3(v), vec3(l),
                                      C.c_int(bool(above)), out)
    return _np3(out)


def strength_component(pl, scatters, component, steps, intensity, x, v, l, above):
    out = _out3()
    lib().orc_strength_component(C.byref(pl), scatter_array(scatters), len(scatters), C.byref(component),
                                 C.c_long(steps), vec3(intensity), vec3(x), vec3(v), vec3(l), C.c_int(bool(above)),
                                 out)
    return _np3(out)


def point_scatter_base(pl, scatters, steps, intensity, x, v, l, above):
    out = _out3()
    lib().orc_point_scatter_base(C.byref(pl), scatter_array(scatters), len(scatters), C.c_long(steps),
                                 vec3(intensity), vec3(x), vec3(v), vec3(l), C.c_int(bool(above)), out)
    return _np3(out)


def _point_cb(fun):
    def cb(_ctx, p, v, l, above, o):
        val = fun(np.array([p[0], p[1], p[2]]), np.array([v[0], v[1], v[2]]), np.array([l[0], l[1], l[2]]),
                  bool(above))
        for i in range(3):
            o[i] = float(val[i])

    return POINT_FN(cb)


def _surface_cb(fun):
    def cb(_ctx, p, l, o):
        val = fun(np.array([p[0], p[1], p[2]]), np.array([l[0], l[1], l[2]]))
        for i in range(3):
            o[i] = float(val[i])

    return SURFACE_FN(cb)


def ray_scatter(pl, scatters, steps, point_scatter_fn, x, v, l, above):
    out = _out3()
    lib().orc_ray_scatter(C.byref(pl), scatter_array(scatters), len(scatters), C.c_long(steps),
                          _point_cb(point_scatter_fn), None, vec3(x), vec3(v), vec3(l), C.c_int(bool(above)), out)
    return _np3(out)


def point_scatter(pl, scatters, ray_scatter_fn, surface_radiance_fn, intensity, sphere_steps, ray_steps, x, v, l,
                  above):
    out = _out3()
    lib().orc_point_scatter(C.byref(pl), scatter_array(scatters), len(scatters), _point_cb(ray_scatter_fn), None,
                            _surface_cb(surface_radiance_fn), None, vec3(intensity), C.c_long(sphere_steps),
                            C.c_long(ray_steps), vec3(x), vec3(v), vec3(l), C.c_int(bool(above)), out)
    return _np3(out)


PHASE_HOOK = C.CFUNCTYPE(C.c_double, C.c_void_p, C.c_double)
EXTREMITY_HOOK = C.CFUNCTYPE(None, C.c_void_p, c_double_p, c_double_p, c_double_p)
TRANSMITTANCE_HOOK = C.CFUNCTYPE(None, C.c_void_p, C.c_void_p, C.c_int, C.c_long, c_double_p, c_double_p, c_double_p)


class TestHooks(C.Structure):
    _fields_ = [("phase", PHASE_HOOK), ("ray_extremity", EXTREMITY_HOOK), ("transmittance", TRANSMITTANCE_HOOK)]


class redefs:
    """Context manager: the oracle's analogue of the reference tests' with-redefs around point-scatter
    (t_atmosphere.clj:330-361).  phase(mu) -> float, ray_extremity(origin, direction) -> vec3,
    transmittance(steps, x, x0) -> vec3; None keeps the real function."""

    def __init__(self, phase=None, ray_extremity=None, transmittance=None):
        def ph(_s, mu):
            return float(phase(mu))

        def ext(_pl, o, d, out):
            val = ray_extremity(np.array([o[0], o[1], o[2]]), np.array([d[0], d[1], d[2]]))
            for i in range(3):
                out[i] = float(val[i])

        def tr(_pl, _sc, _n, steps, x, x0, out):
            val = transmittance(int(steps), np.array([x[0], x[1], x[2]]), np.array([x0[0], x0[1], x0[2]]))
            for i in range(3):
                out[i] = float(val[i])

        self.hooks = TestHooks(PHASE_HOOK(ph) if phase else PHASE_HOOK(),
                               EXTREMITY_HOOK(ext) if ray_extremity else EXTREMITY_HOOK(),
                               TRANSMITTANCE_HOOK(tr) if transmittance else TRANSMITTANCE_HOOK())

    def __enter__(self):
        lib().orc_set_test_hooks(C.byref(self.hooks))
        return self

    def __exit__(self, *exc):
        lib().orc_set_test_hooks(None)
        return False


def in_scatter_from_direction(pl, scatters, ray_scatter_fn, surface_radiance_fn, ray_steps, x, v, l, omega):
    """The integrand point-scatter hands to integral-sphere, for one direction (atmosphere.clj:208-222)."""
    out = _out3()
    lib().orc_in_scatter_from_direction(C.byref(pl), scatter_array(scatters), len(scatters), _point_cb(ray_scatter_fn),
                                        None, _surface_cb(surface_radiance_fn), None, C.c_long(ray_steps), vec3(x),
                                        vec3(v), vec3(l), vec3(omega), out)
    return _np3(out)


def surface_radiance(pl, ray_scatter_fn, steps, x, l):
    out = _out3()
    lib().orc_surface_radiance(C.byref(pl), _point_cb(ray_scatter_fn), None, C.c_long(steps), vec3(x), vec3(l), out)
    return _np3(out)


# ------------------------------------------------------------------ index maps

def horizon_distance(pl, radius):
    return lib().orc_horizon_distance(C.byref(pl), radius)


def elevation_to_index(pl, size, point, direction, above):
    return lib().orc_elevation_to_index(C.byref(pl), C.c_long(size), vec3(point), vec3(direction),
                                        C.c_int(bool(above)))


def index_to_elevation(pl, size, radius, index):
    d = _out3()
    above = C.c_int()
    lib().orc_index_to_elevation(C.byref(pl), C.c_long(size), C.c_double(radius), C.c_double(index), d,
                                 C.byref(above))
    return _np3(d), bool(above.value)


def height_to_index(pl, size, point):
    return lib().orc_height_to_index(C.byref(pl), C.c_long(size), vec3(point))


def index_to_height(pl, size, index):
    out = _out3()
    lib().orc_index_to_height(C.byref(pl), C.c_long(size), C.c_double(index), out)
    return _np3(out)


def sun_elevation_to_index(size, point, l):
    return lib().orc_sun_elevation_to_index(C.c_long(size), vec3(point), vec3(l))


def index_to_sin_sun_elevation(size, index):
    return lib().orc_index_to_sin_sun_elevation(size, index)


def sun_angle_to_index(size, direction, l):
    return lib().orc_sun_angle_to_index(C.c_long(size), vec3(direction), vec3(l))


def index_to_sun_direction(size, direction, sin_sun_elevation, index):
    out = _out3()
    lib().orc_index_to_sun_direction(C.c_long(size), vec3(direction), C.c_double(sin_sun_elevation),
                                     C.c_double(index), out)
    return _np3(out)


def transmittance_forward(pl, shape, point, direction, above):
    idx = (C.c_double * 2)()
    lib().orc_transmittance_forward(C.byref(pl), _shape(shape), vec3(point), vec3(direction), C.c_int(bool(above)),
                                    idx)
    return np.array(list(idx))


def transmittance_backward(pl, shape, hi, ei):
    p, d, a = _out3(), _out3(), C.c_int()
    lib().orc_transmittance_backward(C.byref(pl), _shape(shape), C.c_double(hi), C.c_double(ei), p, d, C.byref(a))
    return _np3(p), _np3(d), bool(a.value)


def surface_radiance_forward(pl, shape, point, l):
    idx = (C.c_double * 2)()
    lib().orc_surface_radiance_forward(C.byref(pl), _shape(shape), vec3(point), vec3(l), idx)
    return np.array(list(idx))


def surface_radiance_backward(pl, shape, hi, si):
    p, l = _out3(), _out3()
    lib().orc_surface_radiance_backward(C.byref(pl), _shape(shape), C.c_double(hi), C.c_double(si), p, l)
    return _np3(p), _np3(l)


def ray_scatter_forward(pl, shape, point, direction, l, above):
    idx = (C.c_double * 4)()
    lib().orc_ray_scatter_forward(C.byref(pl), _shape(shape), vec3(point), vec3(direction), vec3(l),
                                  C.c_int(bool(above)), idx)
    return np.array(list(idx))


def ray_scatter_backward(pl, shape, hi, ei, si, ai):
    p, d, l, a = _out3(), _out3(), _out3(), C.c_int()
    lib().orc_ray_scatter_backward(C.byref(pl), _shape(shape), C.c_double(hi), C.c_double(ei), C.c_double(si),
                                   C.c_double(ai), p, d, l, C.byref(a))
    return _np3(p), _np3(d), _np3(l), bool(a.value)


# ------------------------------------------------------------------ interpolation, packing, files

def interpolate(table, coords):
    """interpolate-value on a numpy table; trailing axis of length ncomp if table.ndim == len(coords) + 1."""
    table = np.ascontiguousarray(table, dtype=np.float64)
    dims = len(coords)
    ncomp = 1 if table.ndim == dims else table.shape[-1]
    out = (C.c_double * ncomp)()
    lib().orc_interpolate(table.ctypes.data_as(c_double_p), _shape(table.shape[:dims]), C.c_int(dims), C.c_int(ncomp),
                          (C.c_double * dims)(*[float(c) for c in coords]), out)
    return np.array(list(out)) if table.ndim != dims else out[0]


def pack_floats(table):
    table = np.ascontiguousarray(table, dtype=np.float64)
    out = np.zeros(table.size, dtype=np.float32)
    lib().orc_pack_floats(table.ctypes.data_as(c_double_p), C.c_long(table.size),
                          out.ctypes.data_as(C.POINTER(C.c_float)))
    return out


def convert_4d_to_2d(table):
    """table: [d][c][b][a](ncomp) -> [d*b][c*a](ncomp)"""
    table = np.ascontiguousarray(table, dtype=np.float64)
    ncomp = 1 if table.ndim == 4 else table.shape[4]
    d, c, b, a = table.shape[:4]
    out = np.zeros((d * b, c * a) + (() if table.ndim == 4 else (ncomp,)))
    lib().orc_convert_4d_to_2d(table.ctypes.data_as(c_double_p), _shape((d, c, b, a)), C.c_int(ncomp),
                               out.ctypes.data_as(c_double_p))
    return out


def spit_floats(path, data):
    data = np.ascontiguousarray(data, dtype=np.float32)
    rc = lib().orc_spit_floats(path.encode(), data.ctypes.data_as(C.POINTER(C.c_float)), C.c_long(data.size))
    if rc != 0:
        raise IOError(path)


def slurp_floats(path):
    n = os.path.getsize(path) // 4
    out = np.zeros(n, dtype=np.float32)
    got = lib().orc_slurp_floats(path.encode(), out.ctypes.data_as(C.POINTER(C.c_float)), C.c_long(n))
    if got != n:
        raise IOError(path)
    return out


# ------------------------------------------------------------------ table builders

def _indices(indices):
    if indices is None:
        return None, 0, None
    arr = np.ascontiguousarray(indices, dtype=np.int64)
    return arr.ctypes.data_as(c_long_p), arr.size, arr


def _tab_ptr(t):
    return t.ctypes.data_as(c_double_p)


def _c64(t):
    return np.ascontiguousarray(t, dtype=np.float64)


def table_transmittance(pl, scatters, cfg, indices=None):
    ptr, n, keep = _indices(indices)
    shape = tuple(cfg.shape_t) + (3,) if indices is None else (n, 3)
    out = np.zeros(shape)
    lib().orc_table_transmittance(C.byref(pl), scatter_array(scatters), len(scatters), C.byref(cfg), ptr, C.c_long(n),
                                  _tab_ptr(out))
    return out


def table_surface_radiance_base(pl, scatters, cfg, indices=None):
    ptr, n, keep = _indices(indices)
    shape = tuple(cfg.shape_e) + (3,) if indices is None else (n, 3)
    out = np.zeros(shape)
    lib().orc_table_surface_radiance_base(C.byref(pl), scatter_array(scatters), len(scatters), C.byref(cfg), ptr,
                                          C.c_long(n), _tab_ptr(out))
    return out


def table_first_order(pl, scatters, cfg, component, strength, indices=None):
    ptr, n, keep = _indices(indices)
    shape = tuple(cfg.shape4) + (3,) if indices is None else (n, 3)
    out = np.zeros(shape)
    lib().orc_table_first_order(C.byref(pl), scatter_array(scatters), len(scatters), C.byref(cfg), C.byref(component),
                                C.c_int(int(strength)), ptr, C.c_long(n), _tab_ptr(out))
    return out


class SSourceSpec:
    """dS closure: single table (kind 0) or rayleigh + mie_strength * phase(mie) (kind 1, atmosphere_lut.clj:79-84)."""

    def __init__(self, tab_a, tab_b=None, phase_component=None):
        self.tab_a = _c64(tab_a)
        self.tab_b = _c64(tab_b) if tab_b is not None else None
        self.phase_component = phase_component
        self.c = SSource(1 if tab_b is not None else 0, _tab_ptr(self.tab_a),
                         _tab_ptr(self.tab_b) if tab_b is not None else None,
                         phase_component if phase_component is not None else scatter((0, 0, 0), 1.0))


def table_point_scatter(pl, scatters, cfg, ds, de, indices=None):
    ptr, n, keep = _indices(indices)
    shape = tuple(cfg.shape4) + (3,) if indices is None else (n, 3)
    out = np.zeros(shape)
    de = _c64(de)
    lib().orc_table_point_scatter(C.byref(pl), scatter_array(scatters), len(scatters), C.byref(cfg), C.byref(ds.c),
                                  _tab_ptr(de), ptr, C.c_long(n), _tab_ptr(out))
    return out


def table_surface_radiance(pl, cfg, ds, indices=None):
    ptr, n, keep = _indices(indices)
    shape = tuple(cfg.shape_e) + (3,) if indices is None else (n, 3)
    out = np.zeros(shape)
    lib().orc_table_surface_radiance(C.byref(pl), C.byref(cfg), C.byref(ds.c), ptr, C.c_long(n), _tab_ptr(out))
    return out


def table_ray_scatter(pl, scatters, cfg, dj, indices=None):
    ptr, n, keep = _indices(indices)
    shape = tuple(cfg.shape4) + (3,) if indices is None else (n, 3)
    out = np.zeros(shape)
    dj = _c64(dj)
    lib().orc_table_ray_scatter(C.byref(pl), scatter_array(scatters), len(scatters), C.byref(cfg), _tab_ptr(dj), ptr,
                                C.c_long(n), _tab_ptr(out))
    return out


def _resample(name, pl, cfg, tabs, full_shape, indices):
    ptr, n, keep = _indices(indices)
    shape = tuple(full_shape) + (3,) if indices is None else (n, 3)
    out = np.zeros(shape)
    tabs = [None if t is None else _c64(t) for t in tabs]
    arr = (c_double_p * len(tabs))(*[None if t is None else _tab_ptr(t) for t in tabs])
    getattr(lib(), name)(C.byref(pl), C.byref(cfg), arr, C.c_int(len(tabs)), ptr, C.c_long(n), _tab_ptr(out))
    return out


def table_resample_sum_4d(pl, cfg, tabs, indices=None):
    return _resample("orc_table_resample_sum_4d", pl, cfg, tabs, cfg.shape4, indices)


def table_resample_sum_e(pl, cfg, tabs, indices=None):
    return _resample("orc_table_resample_sum_e", pl, cfg, tabs, cfg.shape_e, indices)


def table_resample_sum_t(pl, cfg, tabs, indices=None):
    return _resample("orc_table_resample_sum_t", pl, cfg, tabs, cfg.shape_t, indices)


def roundtrip(pl, cfg, which):
    """forward(backward(i)) for every texel of ray-scatter-space (which=0), surface-radiance-space (1) or
    transmittance-space (2); returns an array of shape + (dims,)."""
    shape = {0: tuple(cfg.shape4), 1: tuple(cfg.shape_e), 2: tuple(cfg.shape_t)}[which]
    out = np.zeros(shape + (len(shape),))
    fn = {0: "orc_roundtrip_4d", 1: "orc_roundtrip_e", 2: "orc_roundtrip_t"}[which]
    getattr(lib(), fn)(C.byref(pl), C.byref(cfg), _tab_ptr(out))
    return out


def backward_all(pl, cfg, which):
    """backward(i) of every integer texel of ray-scatter-space (which=0), surface-radiance-space (1) or
    transmittance-space (2): (point, direction, light, above) arrays in row-major texel order."""
    shape = {0: tuple(cfg.shape4), 1: tuple(cfg.shape_e), 2: tuple(cfg.shape_t)}[which]
    n = int(np.prod(shape))
    p, d, l = np.zeros((n, 3)), np.zeros((n, 3)), np.zeros((n, 3))
    ab = np.zeros(n, dtype=np.int32)
    lib().orc_backward_all(C.byref(pl), C.byref(cfg), C.c_int(which), _tab_ptr(p), _tab_ptr(d), _tab_ptr(l),
                           ab.ctypes.data_as(C.POINTER(C.c_int)))
    return p, d, l, ab.astype(bool)


def counters_reset():
    lib().orc_counters_reset()


def counters_get():
    a, b, c = C.c_longlong(), C.c_longlong(), C.c_longlong()
    lib().orc_counters_get(C.byref(a), C.byref(b), C.byref(c))
    return {"esamples": a.value, "lookups4d": b.value, "lookups2d": c.value}


def num_threads():
    return lib().orc_num_threads()


def set_num_threads(n):
    lib().orc_set_num_threads(int(n))


# ------------------------------------------------------------------ atmosphere_lut.clj driver

EARTH = dict(radius=6378000.0, height=35000.0, brightness=(0.3, 0.3, 0.3))
MIE = dict(base=(2e-5, 2e-5, 2e-5), scale=1200.0, g=0.76, quotient=0.9)
RAYLEIGH = dict(base=(5.8e-6, 13.5e-6, 33.1e-6), scale=8000.0)


def generate_atmosphere_luts(pl, mie, rayleigh, cfg, iterations=5, record=None, log=None):
    """generate-atmosphere-luts (atmosphere_lut.clj:43-105) on the oracle's table builders.

    Returns float32 arrays in file layout: transmittance, surface_radiance, ray_scatter, mie_strength.
    `record` (a dict) receives every intermediate table as float64.
    """
    scatters = [mie, rayleigh]                                                   # :64
    rec = record if record is not None else {}
    T = table_transmittance(pl, scatters, cfg)                                   # :74
    dE = table_surface_radiance_base(pl, scatters, cfg)                          # :75
    E = None                                                                     # :76 (constantly 0)
    R1 = table_first_order(pl, scatters, cfg, rayleigh, 0)                       # :77
    M1 = table_first_order(pl, scatters, cfg, mie, 1)                            # :78
    dS = SSourceSpec(R1, M1, mie)                                                # :79-84
    S = R1                                                                       # :85
    rec.update(T=T, Ebase=dE, R1=R1, M1=M1)
    for it in range(iterations):                                                 # :86
        if log:
            log("Iteration %d/%d" % (it + 1, iterations))
        dJ = table_point_scatter(pl, scatters, cfg, dS, dE)                      # :88,90
        dE_new = table_surface_radiance(pl, cfg, dS)                             # :89,92 (old dS)
        dS_new = table_ray_scatter(pl, scatters, cfg, dJ)                        # :91,93
        dE = dE_new
        dS = SSourceSpec(dS_new)
        E = table_resample_sum_e(pl, cfg, [E, dE])                               # :94-95
        S = table_resample_sum_4d(pl, cfg, [S, dS_new])                          # :96-97
        rec["dJ%d" % it] = dJ
        rec["dE%d" % it] = dE
        rec["dS%d" % it] = dS_new
        rec["E%d" % it] = E
        rec["S%d" % it] = S
    lt = table_resample_sum_t(pl, cfg, [T])                                      # :98
    le = table_resample_sum_e(pl, cfg, [E])                                      # :99
    ls = table_resample_sum_4d(pl, cfg, [S])                                     # :100
    lm = table_resample_sum_4d(pl, cfg, [M1])                                    # :101
    rec.update(LT=lt, LE=le, LS=ls, LM=lm)
    return (pack_floats(lt), pack_floats(le), pack_floats(convert_4d_to_2d(ls)),  # :102-105
            pack_floats(convert_4d_to_2d(lm)))
