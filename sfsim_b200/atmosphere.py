"""Host-side mirror of the public surface of sfsim.atmosphere (src/clj/sfsim/atmosphere.clj:35-422).

Same names, argument order and meaning as the reference's functions; every evaluation runs on the GPU
through the C ABI's batch entry points (one item per call here; pass arrays to the *_batch helpers for
many).  `planet` and scatter components are dicts like the reference's maps:

    earth    = {"centre": (0, 0, 0), "radius": 6378000.0, "height": 100000.0, "brightness": (0.3, 0.3, 0.3)}
    rayleigh = {"base": (5.8e-6, 13.5e-6, 33.1e-6), "scale": 8000.0}
    mie      = {"base": (2e-5, 2e-5, 2e-5), "scale": 1200.0, "g": 0.76, "quotient": 0.9}

Functions that take other functions in the reference (`ray-scatter` takes a point-scatter function,
`point-scatter` takes ray-scatter and surface-radiance functions) take *source descriptors* here, the
closed set the LUT build uses: first-order components (FirstOrder) and interpolated tables
(sfsim_b200.interpolate.InterpolationTable).
"""
import ctypes as C
import math

import numpy as np

from . import _lib
from ._lib import check

# ------------------------------------------------------------------ medium (pure scalar formulas)


def scattering(component, height_or_planet, x=None):
    """atmosphere.clj:42-47: scatter-base * exp(-height / scatter-scale); 3-arity takes (planet component x)."""
    if x is not None:  # (scattering planet component x)
        planet, component, point = component, height_or_planet, x
        return scattering(component, height(planet, point))
    h = float(height_or_planet)
    return np.asarray(component["base"], dtype=np.float64) * math.exp(-(h / component["scale"]))


def extinction(scattering_type, h):
    """atmosphere.clj:50-53"""
    return scattering(scattering_type, h) / scattering_type.get("quotient", 1.0)


def phase(component, mu):
    """atmosphere.clj:56-61 (Cornette-Shanks; g defaults to 0 -> Rayleigh)"""
    g = component.get("g", 0.0) if component else 0.0
    g2 = g * g
    return (3.0 * (1.0 - g2) * (1.0 + mu * mu)) / (8.0 * math.pi * (2.0 + g2) * math.pow((1.0 + g2) - 2.0 * g * mu, 1.5))


def medium_batch(component, fn, args):
    """scattering (fn 0), extinction (fn 1) of `component` at heights `args`, or its phase function (fn 2) at cosines
    `args`, evaluated by the device functions the table kernels inline (atmlut_medium_batch): array [n][3] (fn 0, 1)
    or [n] (fn 2).  The functions above are the host-side mirror of the same formulas."""
    import ctypes as C
    from . import _lib
    a = _lib.f64(args).reshape(-1)
    out = np.zeros((len(a), 3))
    sc = _lib.make_scatter(component)
    _lib.check(_lib.load().atmlut_medium_batch(C.byref(sc), int(fn), len(a), _lib.ptr(a), _lib.ptr(out)))
    return out[:, 0] if fn == 2 else out


def height(planet, point):
    """sphere.clj:28-31"""
    d = np.asarray(point, dtype=np.float64) - np.asarray(planet.get("centre", (0, 0, 0)), dtype=np.float64)
    return float(np.linalg.norm(d)) - planet["radius"]


# ------------------------------------------------------------------ helpers

def _pts(a):
    a = _lib.f64(a)
    if a.ndim == 1:
        a = a.reshape(1, 3)
    return a


def _scatter_key(s):
    c = _lib.make_scatter(s)
    return (tuple(c.base), c.scale, c.g, c.quotient)


class FirstOrder:
    """First-order point-scatter functions of atmosphere.clj:170-189 bound to their leading arguments, i.e.
    (partial point-scatter-component planet scatter component steps intensity) and friends."""

    COMPONENT, STRENGTH, BASE = 0, 1, 2

    def __init__(self, kind, planet, scatter, component, steps, intensity):
        self.kind, self.planet, self.scatter, self.steps = kind, planet, list(scatter), int(steps)
        self.intensity = tuple(float(v) for v in intensity)
        if kind == FirstOrder.BASE:
            self.component = 0
        else:
            matches = [i for i, s in enumerate(self.scatter) if s is component or s == component]
            if not matches:
                raise ValueError("component must be one of the scatter components")
            self.component = matches[0]

    def __call__(self, x, view_direction, light_direction, above_horizon=True):
        return self.batch(_pts(x), _pts(view_direction), _pts(light_direction))[0]

    def batch(self, x, v, l):
        lib = _lib.load()
        x, v, l = _pts(x), _pts(v), _pts(l)
        out = np.zeros_like(x)
        pl = _lib.make_planet(self.planet)
        sc = _lib.make_scatter_array(self.scatter)
        check(lib.atmlut_point_scatter_first_order_batch(C.byref(pl), sc, len(self.scatter), self.kind,
                                                         self.component, self.steps, _lib.vec3(self.intensity),
                                                         len(x), _lib.ptr(x), _lib.ptr(v), _lib.ptr(l),
                                                         _lib.ptr(out)))
        return out


# ------------------------------------------------------------------ radiative quantities

def transmittance(planet, scatter, steps, x, x0_or_v, above_horizon=None):
    """atmosphere.clj:114-128: 5-arity (x -> x0) or 6-arity (x along v to the shell / the ground)."""
    return transmittance_batch(planet, scatter, steps, _pts(x), _pts(x0_or_v),
                               None if above_horizon is None else [above_horizon])[0]


def transmittance_batch(planet, scatter, steps, x, x0_or_v, above_horizon=None):
    lib = _lib.load()
    x, y = _pts(x), _pts(x0_or_v)
    out = np.zeros_like(x)
    pl = _lib.make_planet(planet)
    sc = _lib.make_scatter_array(scatter)
    if above_horizon is None:
        check(lib.atmlut_transmittance_batch(C.byref(pl), sc, len(scatter), int(steps), len(x), _lib.ptr(x),
                                             _lib.ptr(y), _lib.ptr(out)))
    else:
        ab = _lib.i32(np.asarray(above_horizon, dtype=bool))
        check(lib.atmlut_transmittance_dir_batch(C.byref(pl), sc, len(scatter), int(steps), len(x), _lib.ptr(x),
                                                 _lib.ptr(y), _lib.ptr(ab), _lib.ptr(out)))
    return out


def surface_radiance_base(planet, scatter, steps, intensity, x, light_direction):
    """atmosphere.clj:131-137"""
    lib = _lib.load()
    x, l = _pts(x), _pts(light_direction)
    out = np.zeros_like(x)
    pl = _lib.make_planet(planet)
    sc = _lib.make_scatter_array(scatter)
    check(lib.atmlut_surface_radiance_base_batch(C.byref(pl), sc, len(scatter), int(steps), _lib.vec3(intensity),
                                                 len(x), _lib.ptr(x), _lib.ptr(l), _lib.ptr(out)))
    return out[0]


def point_scatter_component(planet, scatter, component, steps, intensity, x, view_direction, light_direction,
                            above_horizon=True):
    """atmosphere.clj:170-174"""
    return FirstOrder(FirstOrder.COMPONENT, planet, scatter, component, steps, intensity)(x, view_direction,
                                                                                          light_direction)


def strength_component(planet, scatter, component, steps, intensity, x, view_direction, light_direction,
                       above_horizon=True):
    """atmosphere.clj:177-182"""
    return FirstOrder(FirstOrder.STRENGTH, planet, scatter, component, steps, intensity)(x, view_direction,
                                                                                         light_direction)


def point_scatter_base(planet, scatter, steps, intensity, x, view_direction, light_direction, above_horizon=True):
    """atmosphere.clj:185-189"""
    return FirstOrder(FirstOrder.BASE, planet, scatter, None, steps, intensity)(x, view_direction, light_direction)


def _source_tables(src):
    """(table a, table b or None, mie component or None) of a ray-scatter source: an InterpolationTable over
    ray-scatter-space or the MieCombined closure of atmosphere_lut.clj:79-84."""
    from . import interpolate
    if isinstance(src, interpolate.MieCombined):
        a = src.a.table if isinstance(src.a, interpolate.InterpolationTable) else src.a
        b = src.b.table if isinstance(src.b, interpolate.InterpolationTable) else src.b
        return _lib.f32(a), _lib.f32(b), src.mie
    if isinstance(src, interpolate.InterpolationTable):
        return _lib.f32(src.table), None, None
    raise TypeError("expected an InterpolationTable or MieCombined source, got %r" % (src,))


def ray_scatter(planet, scatter, steps, point_scatter, x, view_direction, light_direction, above_horizon):
    """atmosphere.clj:192-200.  `point_scatter` is a FirstOrder source or an InterpolationTable over
    point-scatter-space (the interpolated dJ of atmosphere_lut.clj:90-91)."""
    lib = _lib.load()
    x, v, l = _pts(x), _pts(view_direction), _pts(light_direction)
    ab = _lib.i32(np.broadcast_to(np.asarray(above_horizon, dtype=bool), (len(x),)))
    out = np.zeros_like(x)
    pl = _lib.make_planet(planet)
    sc = _lib.make_scatter_array(scatter)
    if isinstance(point_scatter, FirstOrder):
        # component and kind were resolved against the scatter list the FirstOrder source was made with: the same
        # list must be integrated here, or another component would be picked without a word
        if [_scatter_key(a) for a in point_scatter.scatter] != [_scatter_key(a) for a in scatter]:
            raise TypeError("ray_scatter: `scatter` differs from the scatter list of the first-order point-scatter source")
        check(lib.atmlut_ray_scatter_first_order_batch(C.byref(pl), sc, len(scatter), point_scatter.kind,
                                                       point_scatter.component, int(steps),
                                                       _lib.vec3(point_scatter.intensity), len(x), _lib.ptr(x),
                                                       _lib.ptr(v), _lib.ptr(l), _lib.ptr(ab), _lib.ptr(out)))
    else:
        dj, _, _ = _source_tables(point_scatter)
        shape = (C.c_int * 4)(*dj.shape[:4])
        check(lib.atmlut_ray_scatter_table_batch(C.byref(pl), sc, len(scatter), int(steps), shape, _lib.ptr(dj),
                                                 len(x), _lib.ptr(x), _lib.ptr(v), _lib.ptr(l), _lib.ptr(ab),
                                                 _lib.ptr(out)))
    return out[0]


def point_scatter(planet, scatter, ray_scatter, surface_radiance, intensity, sphere_steps, ray_steps, x,
                  view_direction, light_direction, above_horizon=True):
    """atmosphere.clj:203-222 with `ray_scatter` an InterpolationTable / MieCombined over ray-scatter-space and
    `surface_radiance` an InterpolationTable over surface-radiance-space (intensity and above-horizon are
    ignored, like the reference's underscored parameters)."""
    lib = _lib.load()
    x, v, l = _pts(x), _pts(view_direction), _pts(light_direction)
    out = np.zeros_like(x)
    a, b, mie = _source_tables(ray_scatter)
    e_tab = _lib.f32(surface_radiance.table)
    scatter = list(scatter)
    phase_component = 0
    if b is not None:
        matches = [i for i, s in enumerate(scatter) if s is mie or s == mie]
        if not matches:
            raise ValueError("the Mie component of the source must be one of the scatter components")
        phase_component = matches[0]
    pl = _lib.make_planet(planet)
    sc = _lib.make_scatter_array(scatter)
    shape4 = (C.c_int * 4)(*a.shape[:4])
    shape_e = (C.c_int * 2)(*e_tab.shape[:2])
    check(lib.atmlut_point_scatter_batch(C.byref(pl), sc, len(scatter), int(sphere_steps), int(ray_steps), shape4,
                                         _lib.ptr(a), _lib.ptr(b), phase_component, shape_e, _lib.ptr(e_tab), len(x),
                                         _lib.ptr(x), _lib.ptr(v), _lib.ptr(l), _lib.ptr(out)))
    return out[0]


def surface_radiance(planet, ray_scatter, steps, x, light_direction):
    """atmosphere.clj:225-230 with `ray_scatter` an InterpolationTable / MieCombined over ray-scatter-space."""
    lib = _lib.load()
    x, l = _pts(x), _pts(light_direction)
    out = np.zeros_like(x)
    a, b, mie = _source_tables(ray_scatter)
    scatter = [mie] if b is not None else []
    pl = _lib.make_planet(planet)
    sc = _lib.make_scatter_array(scatter)
    shape4 = (C.c_int * 4)(*a.shape[:4])
    check(lib.atmlut_surface_radiance_batch(C.byref(pl), sc, len(scatter), int(steps), shape4, _lib.ptr(a), _lib.ptr(b),
                                            0, len(x), _lib.ptr(x), _lib.ptr(l), _lib.ptr(out)))
    return out[0]


# ------------------------------------------------------------------ interpolation spaces (atmosphere.clj:233-422)

class Space:
    """interpolation-space (interpolate.clj:22): shape + forward + backward, evaluated on the device.

    The reference's index maps take raw points: `height` subtracts the planet centre (sphere.clj:28-31) but
    `is-above-horizon?` and the spaces assume the origin (atmosphere.clj:95-102).  The device maps subtract
    `planet["centre"]` from every point first, which is the same thing for the only planets the tables are built for
    (centre at the origin; the table builders refuse any other) and keeps the maps usable for a shifted planet."""

    def __init__(self, which, planet, shape):
        self.which, self.planet, self.shape = which, planet, tuple(int(s) for s in shape)

    def _call_forward(self, point, direction, light, above):
        lib = _lib.load()
        point = _pts(point)
        n = len(point)
        dims = 4 if self.which == 0 else 2
        idx = np.zeros((n, dims))
        pl = _lib.make_planet(self.planet)
        shape = (C.c_int * len(self.shape))(*self.shape)
        d = _pts(direction) if direction is not None else None
        l = _pts(light) if light is not None else None
        ab = _lib.i32(np.broadcast_to(np.asarray(above, dtype=bool), (n,))) if above is not None else None
        check(lib.atmlut_index_forward_batch(C.byref(pl), self.which, shape, n, _lib.ptr(point), _lib.ptr(d),
                                             _lib.ptr(l), _lib.ptr(ab), _lib.ptr(idx)))
        return idx

    def _call_backward(self, indices):
        lib = _lib.load()
        dims = 4 if self.which == 0 else 2
        indices = _lib.f64(indices).reshape(-1, dims)
        n = len(indices)
        p, d, l = np.zeros((n, 3)), np.zeros((n, 3)), np.zeros((n, 3))
        ab = np.zeros(n, dtype=np.int32)
        pl = _lib.make_planet(self.planet)
        shape = (C.c_int * len(self.shape))(*self.shape)
        check(lib.atmlut_index_backward_batch(C.byref(pl), self.which, shape, n, _lib.ptr(indices), _lib.ptr(p),
                                              _lib.ptr(d), _lib.ptr(l), _lib.ptr(ab)))
        return p, d, l, ab.astype(bool)


class RayScatterSpace(Space):
    """ray-scatter-space / point-scatter-space (atmosphere.clj:415-422)"""

    def __init__(self, planet, shape):
        super().__init__(0, planet, shape)

    def forward(self, point, direction, light_direction, above_horizon):
        return self._call_forward(point, direction, light_direction, above_horizon)[0]

    def backward(self, height_index, elevation_index, sun_elevation_index, sun_angle_index):
        p, d, l, ab = self._call_backward([height_index, elevation_index, sun_elevation_index, sun_angle_index])
        return p[0], d[0], l[0], bool(ab[0])


class SurfaceRadianceSpace(Space):
    """surface-radiance-space (atmosphere.clj:359-365)"""

    def __init__(self, planet, shape):
        super().__init__(1, planet, shape)

    def forward(self, point, light_direction):
        return self._call_forward(point, None, light_direction, None)[0]

    def backward(self, height_index, sun_elevation_index):
        p, _, l, _ = self._call_backward([height_index, sun_elevation_index])
        return p[0], l[0]


class TransmittanceSpace(Space):
    """transmittance-space (atmosphere.clj:313-319)"""

    def __init__(self, planet, shape):
        super().__init__(2, planet, shape)

    def forward(self, point, direction, above_horizon):
        return self._call_forward(point, direction, None, above_horizon)[0]

    def backward(self, height_index, elevation_index):
        p, d, _, ab = self._call_backward([height_index, elevation_index])
        return p[0], d[0], bool(ab[0])


def ray_scatter_space(planet, shape):
    return RayScatterSpace(planet, shape)


point_scatter_space = ray_scatter_space


def surface_radiance_space(planet, shape):
    return SurfaceRadianceSpace(planet, shape)


def transmittance_space(planet, shape):
    return TransmittanceSpace(planet, shape)


# scalar index maps (atmosphere.clj:233-384), evaluated on the device by the same code the table kernels use

def _index_map(planet, fn, size, a, b=None, flag=None):
    lib = _lib.load()
    a = _pts(a)
    n = len(a)
    bb = _pts(b) if b is not None else None
    fl = _lib.i32(np.broadcast_to(np.asarray(flag, dtype=bool), (n,))) if flag is not None else None
    out = np.zeros((n, 3))
    oflag = np.zeros(n, dtype=np.int32)
    pl = _lib.make_planet(planet if planet is not None else {"radius": 1.0, "height": 1.0})
    check(lib.atmlut_index_map_batch(C.byref(pl), fn, int(size), n, _lib.ptr(a), _lib.ptr(bb), _lib.ptr(fl),
                                     _lib.ptr(out), _lib.ptr(oflag)))
    return out, oflag


def horizon_distance(planet, radius):
    """atmosphere.clj:233-236"""
    return float(_index_map(planet, 8, 2, (radius, 0, 0))[0][0][0])


def elevation_to_index(planet, size, point, direction, above_horizon):
    """atmosphere.clj:239-253"""
    return float(_index_map(planet, 0, size, point, direction, above_horizon)[0][0][0])


def index_to_elevation(planet, size, radius, index):
    """atmosphere.clj:256-270 -> [direction above-horizon]"""
    out, flag = _index_map(planet, 1, size, (radius, index, 0))
    return out[0], bool(flag[0])


def height_to_index(planet, size, point):
    """atmosphere.clj:273-278"""
    return float(_index_map(planet, 2, size, point)[0][0][0])


def index_to_height(planet, size, index):
    """atmosphere.clj:281-288"""
    return _index_map(planet, 3, size, (index, 0, 0))[0][0]


def sun_elevation_to_index(size, point, light_direction):
    """atmosphere.clj:322-326"""
    return float(_index_map(None, 4, size, point, light_direction)[0][0][0])


def index_to_sin_sun_elevation(size, index):
    """atmosphere.clj:329-332"""
    return float(_index_map(None, 5, size, (index, 0, 0))[0][0][0])


def sun_angle_to_index(size, direction, light_direction):
    """atmosphere.clj:368-372"""
    return float(_index_map(None, 6, size, direction, light_direction)[0][0][0])


def index_to_sun_direction(size, direction, sin_sun_elevation, index):
    """atmosphere.clj:375-384"""
    return _index_map(None, 7, size, direction, (sin_sun_elevation, index, 0))[0][0]
