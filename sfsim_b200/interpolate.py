"""Host-side mirror of sfsim.interpolate (src/clj/sfsim/interpolate.clj): make-lookup-table,
interpolation-table, interpolate-function, linear-space, clip, mix, compose-space.

`make_lookup_table(fun, space)` tabulates `fun` over `space` like the reference (interpolate.clj:68-72).
When `fun` is one of the atmosphere function descriptors below and `space` one of the atmosphere spaces,
the table is integrated by the CUDA table kernels through the C ABI (that is the accelerated path);
for ordinary Python callables over a `linear_space` it is plain host logic, as in the reference.
Tables are numpy arrays: shape + (3,) float32 for RGB tables.
"""
import ctypes as C
import itertools
import math

import numpy as np

from . import _lib, atmosphere
from ._lib import check

# ------------------------------------------------------------------ generic part (interpolate.clj:25-51, 75-84, 113-118)


class LinearSpace:
    """linear-space (interpolate.clj:47-51)"""

    def __init__(self, minima, maxima, shape):
        self.minima, self.maxima, self.shape = list(minima), list(maxima), tuple(shape)

    def forward(self, *point):
        return [(x - a) / (b - a) * (n - 1) for x, a, b, n in zip(point, self.minima, self.maxima, self.shape)]

    def backward(self, *indices):
        return [i / (n - 1) * (b - a) + a for i, a, b, n in zip(indices, self.minima, self.maxima, self.shape)]


def linear_space(minima, maxima, shape):
    return LinearSpace(minima, maxima, shape)


class ComposedSpace:
    """compose-space (interpolate.clj:113-118)"""

    def __init__(self, f, g):
        self.shape = f.shape
        self._f, self._g = f, g

    def forward(self, *args):
        r = self._g.forward(*args)
        return self._f.forward(*(r if isinstance(r, (list, tuple)) else [r]))

    def backward(self, *args):
        r = self._f.backward(*args)
        return self._g.backward(*(r if isinstance(r, (list, tuple)) else [r]))


def compose_space(f, g):
    return ComposedSpace(f, g)


def clip(value, size):
    """interpolate.clj:75-78"""
    return min(max(float(value), 0.0), float(size - 1))


def mix(a, b, scalar):
    """interpolate.clj:81-84"""
    return a * (1 - scalar) + b * scalar


# ------------------------------------------------------------------ function descriptors for the CUDA table kernels

class Transmittance:
    """(partial transmittance planet scatter steps), atmosphere_lut.clj:65"""

    def __init__(self, planet, scatter, steps):
        self.planet, self.scatter, self.steps = planet, list(scatter), int(steps)


class SurfaceRadianceBase:
    """(partial surface-radiance-base planet scatter steps intensity), atmosphere_lut.clj:66"""

    def __init__(self, planet, scatter, steps, intensity):
        self.planet, self.scatter, self.steps, self.intensity = planet, list(scatter), int(steps), tuple(intensity)


class RayScatter:
    """(partial ray-scatter planet scatter steps point-scatter), atmosphere_lut.clj:71-72,91.
    point_scatter: atmosphere.FirstOrder (first order) or InterpolationTable over point-scatter-space."""

    def __init__(self, planet, scatter, steps, point_scatter):
        self.planet, self.scatter, self.steps, self.point_scatter = planet, list(scatter), int(steps), point_scatter


class MieCombined:
    """The iteration-1 dS closure, atmosphere_lut.clj:79-84: rayleigh(x v l a) + mie_strength(x v l a) * phase(mie, v.l)"""

    def __init__(self, rayleigh_table, mie_strength_table, mie, scatter):
        self.a, self.b, self.mie = rayleigh_table, mie_strength_table, mie
        self.phase_component = [i for i, s in enumerate(scatter) if s is mie or s == mie][0]


class PointScatter:
    """(partial point-scatter planet scatter ray-scatter surface-radiance intensity sphere-steps ray-steps),
    atmosphere_lut.clj:88.  ray_scatter: InterpolationTable or MieCombined; surface_radiance: InterpolationTable."""

    def __init__(self, planet, scatter, ray_scatter, surface_radiance, intensity, sphere_steps, ray_steps):
        self.planet, self.scatter = planet, list(scatter)
        self.ray_scatter, self.surface_radiance = ray_scatter, surface_radiance
        self.intensity, self.sphere_steps, self.ray_steps = tuple(intensity), int(sphere_steps), int(ray_steps)


class SurfaceRadiance:
    """(partial surface-radiance planet ray-scatter steps), atmosphere_lut.clj:89"""

    def __init__(self, planet, ray_scatter, steps):
        self.planet, self.ray_scatter, self.steps = planet, ray_scatter, int(steps)


class TableSum:
    """(fn [& args] (add (a args) (b args))) of two interpolation tables, atmosphere_lut.clj:94-97; b may be None"""

    def __init__(self, a, b=None):
        self.a, self.b = a, b


# ------------------------------------------------------------------ interpolation-table (interpolate.clj:87-110)

class InterpolationTable:
    """interpolation-table: callable doing forward map + multilinear lookup (interpolate.clj:87-104)."""

    def __init__(self, lookup_table, space):
        self.table = np.ascontiguousarray(lookup_table, dtype=np.float32)
        self.space = space

    def __call__(self, *coords):
        idx = self.space.forward(*coords)
        return interpolate_value(self.table, idx)


def interpolate_value(lookup_table, point):
    """interpolate-value (interpolate.clj:87-98) on the device."""
    lib = _lib.load()
    table = np.ascontiguousarray(lookup_table, dtype=np.float32)
    point = [float(p) for p in np.asarray(point, dtype=np.float64).reshape(-1)]
    dims = len(point)
    ncomp = 1 if table.ndim == dims else table.shape[-1]
    out = np.zeros(ncomp, dtype=np.float32)
    shape = (C.c_int * dims)(*table.shape[:dims])
    coords = np.asarray(point, dtype=np.float64)
    check(lib.atmlut_interpolate_batch(_lib.ptr(table), shape, dims, ncomp, 1, _lib.ptr(coords), _lib.ptr(out)))
    return float(out[0]) if table.ndim == dims else out.astype(np.float64)


def interpolation_table(lookup_table, space):
    return InterpolationTable(lookup_table, space)


def _config_for(shape4=None, shape_t=None, shape_e=None, **kw):
    cfg = _lib.default_config()
    return _lib.make_config(ray_scatter_shape=shape4 or cfg.ray_scatter_shape,
                            transmittance_shape=shape_t or cfg.transmittance_shape,
                            surface_radiance_shape=shape_e or cfg.surface_radiance_shape, **kw)


def _table_of(src, shape=None, what="source table"):
    """float32 host table of an InterpolationTable or a plain array.  The C entry points take no table-shape
    argument and read prod(shape) * 3 floats from the pointer, so a table of any other shape is refused here."""
    tab = _lib.f32(src.table if isinstance(src, InterpolationTable) else src)
    if shape is not None and tuple(tab.shape) != tuple(shape) + (3,):
        raise TypeError("%s has shape %s, but the space being tabulated needs %s" %
                        (what, tuple(tab.shape), tuple(shape) + (3,)))
    return tab


def _same_space(src, space, what):
    """A re-tabulation reads `src` through forward o backward of `space`: both must describe the same space."""
    other = getattr(src, "space", None)
    if other is None or not isinstance(other, atmosphere.Space):
        return
    if other.which != space.which or tuple(other.shape) != tuple(space.shape) or \
            _lib.make_planet(other.planet).radius != _lib.make_planet(space.planet).radius or \
            _lib.make_planet(other.planet).height != _lib.make_planet(space.planet).height:
        raise TypeError("%s was tabulated over a different space (kind, shape or planet) than the one it is "
                        "re-tabulated over" % what)


def make_lookup_table(fun, space):
    """make-lookup-table (interpolate.clj:68-72)."""
    lib = _lib.load()
    shape = tuple(space.shape)
    if isinstance(fun, Transmittance):
        cfg = _config_for(shape_t=shape, ray_steps=fun.steps)
        out = np.zeros(shape + (3,), np.float32)
        pl, sc = _lib.make_planet(fun.planet), _lib.make_scatter_array(fun.scatter)
        check(lib.atmlut_transmittance_table(C.byref(pl), sc, len(fun.scatter), C.byref(cfg), _lib.ptr(out)))
        return out
    if isinstance(fun, SurfaceRadianceBase):
        cfg = _config_for(shape_e=shape, ray_steps=fun.steps, intensity=fun.intensity)
        out = np.zeros(shape + (3,), np.float32)
        pl, sc = _lib.make_planet(fun.planet), _lib.make_scatter_array(fun.scatter)
        check(lib.atmlut_surface_radiance_base_table(C.byref(pl), sc, len(fun.scatter), C.byref(cfg), _lib.ptr(out)))
        return out
    if isinstance(fun, RayScatter):
        pl, sc = _lib.make_planet(fun.planet), _lib.make_scatter_array(fun.scatter)
        out = np.zeros(shape + (3,), np.float32)
        src = fun.point_scatter
        if isinstance(src, atmosphere.FirstOrder):
            if src.kind == atmosphere.FirstOrder.BASE:
                raise TypeError("tabulate point-scatter-base as the sum of its components")
            cfg = _config_for(shape4=shape, ray_steps=fun.steps, intensity=src.intensity)
            check(lib.atmlut_first_order_tables(C.byref(pl), sc, len(fun.scatter), C.byref(cfg), src.component,
                                                src.kind, _lib.ptr(out), 0, 0, None))
        else:
            cfg = _config_for(shape4=shape, ray_steps=fun.steps)
            dj = _table_of(src, shape, "the point-scatter table")
            check(lib.atmlut_ray_scatter_table(C.byref(pl), sc, len(fun.scatter), C.byref(cfg), _lib.ptr(dj),
                                               _lib.ptr(out)))
        return out
    if isinstance(fun, PointScatter):
        e_tab = _table_of(fun.surface_radiance)
        if e_tab.ndim != 3 or e_tab.shape[2] != 3:
            raise TypeError("the surface-radiance table must have shape [height][sun elevation][3]")
        cfg = _config_for(shape4=shape, shape_e=e_tab.shape[:2], ray_steps=fun.ray_steps,
                          sphere_steps=fun.sphere_steps, intensity=fun.intensity)
        pl, sc = _lib.make_planet(fun.planet), _lib.make_scatter_array(fun.scatter)
        out = np.zeros(shape + (3,), np.float32)
        rs = fun.ray_scatter
        a = _table_of(rs.a if isinstance(rs, MieCombined) else rs, shape, "the ray-scatter table")
        b = _table_of(rs.b, shape, "the Mie-strength table") if isinstance(rs, MieCombined) else None
        check(lib.atmlut_point_scatter_table(C.byref(pl), sc, len(fun.scatter), C.byref(cfg), _lib.ptr(a),
                                             _lib.ptr(b), rs.phase_component if isinstance(rs, MieCombined) else 0,
                                             _lib.ptr(e_tab), _lib.ptr(out)))
        return out
    if isinstance(fun, SurfaceRadiance):
        rs = fun.ray_scatter
        a = _table_of(rs.a if isinstance(rs, MieCombined) else rs)
        if a.ndim != 5 or a.shape[4] != 3:
            raise TypeError("the ray-scatter table must have shape [height][elevation][sun elevation][heading][3]")
        b = _table_of(rs.b, a.shape[:4], "the Mie-strength table") if isinstance(rs, MieCombined) else None
        cfg = _config_for(shape4=a.shape[:4], shape_e=shape, ray_steps=fun.steps)
        pl = _lib.make_planet(fun.planet)
        # surface-radiance takes no scatter argument (atmosphere.clj:225-230); the phase of MieCombined needs g
        scatter = [rs.mie] if isinstance(rs, MieCombined) else []
        sc = _lib.make_scatter_array(scatter)
        out = np.zeros(shape + (3,), np.float32)
        check(lib.atmlut_surface_radiance_table(C.byref(pl), sc, len(scatter), C.byref(cfg), _lib.ptr(a), _lib.ptr(b),
                                                0, _lib.ptr(out)))
        return out
    if isinstance(fun, (TableSum, InterpolationTable)) and isinstance(space, atmosphere.Space):
        ts = fun if isinstance(fun, TableSum) else TableSum(fun)
        for name, part in (("the first table", ts.a), ("the second table", ts.b)):
            if part is not None:
                _same_space(part, space, name)
        a = _table_of(ts.a, shape, "the first table") if ts.a is not None else None
        b = _table_of(ts.b, shape, "the second table") if ts.b is not None else None
        kw = {0: "shape4", 1: "shape_e", 2: "shape_t"}[space.which]
        cfg = _config_for(**{kw: shape})
        pl = _lib.make_planet(space.planet)
        out = np.zeros(shape + (3,), np.float32)
        check(lib.atmlut_resample_table(C.byref(pl), C.byref(cfg), space.which, _lib.ptr(a), _lib.ptr(b),
                                        _lib.ptr(out)))
        return out
    if callable(fun):
        # generic host path (the reference's own unit tests tabulate sqr and * over linear spaces)
        first = fun(*space.backward(*[0.0] * len(shape)))
        first = np.asarray(first, dtype=np.float64)
        out = np.zeros(shape + first.shape, dtype=np.float64)
        for idx in itertools.product(*[range(n) for n in shape]):
            out[idx] = fun(*space.backward(*[float(i) for i in idx]))
        return out
    raise TypeError("cannot tabulate %r" % (fun,))


def interpolate_function(fun, space):
    """interpolate-function (interpolate.clj:107-110)"""
    return interpolation_table(make_lookup_table(fun, space), space)
