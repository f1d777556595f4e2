"""The written `.scatter` files through the reference's consumer path, without OpenGL.

`make-atmosphere-luts` (atmosphere.clj:587-619) loads the four files into RGB32F textures of fixed sizes and
the shaders in resources/shaders/atmosphere/ sample them.  tests/glsl_emulator.py restates those shader
lookups; here the files written by the CUDA build at SHIPPED resolution are read back exactly like the loader
does (slurp-floats, width x height) and sampled at arbitrary points.  The sampled values must agree with the
direct double-precision evaluation (oracle) up to the table's interpolation error.
"""
import os

import numpy as np
import pytest

from oracle import oracle as orc
from sfsim_b200 import _lib, atmosphere_lut
from tests import glsl_emulator as glsl

pytestmark = pytest.mark.gpu

RADIUS, HEIGHT = 6378000.0, 35000.0


@pytest.fixture(scope="module")
def first_order_files(tmp_path_factory):
    """Shipped shapes, iterations = 0: ray-scatter.scatter then holds first-order Rayleigh scatter only, which
    the oracle can evaluate directly at any point."""
    out_dir = str(tmp_path_factory.mktemp("atmosphere"))
    cfg = _lib.make_config(iterations=0)
    paths = atmosphere_lut.generate_atmosphere_luts(out_dir, cfg=cfg)
    return cfg, paths


def load_like_the_renderer(cfg, paths):
    """atmosphere.clj:587-619: slurp-floats + make-vector-texture-2d / -4d with the hard-coded sizes."""
    h, e, s, a = cfg.ray_scatter_shape
    t = orc.slurp_floats(paths[0]).reshape(cfg.transmittance_height_size, cfg.transmittance_elevation_size, 3)
    surf = orc.slurp_floats(paths[1]).reshape(cfg.surface_height_size, cfg.surface_sun_elevation_size, 3)
    ray = orc.slurp_floats(paths[2]).reshape(h * s, e * a, 3)
    mie = orc.slurp_floats(paths[3]).reshape(h * s, e * a, 3)
    return glsl.Atmosphere(RADIUS, HEIGHT, t, ray, mie, (h, e, s, a)), surf


def test_file_sizes_match_the_loader(first_order_files):
    cfg, paths = first_order_files
    assert [os.path.getsize(p) for p in paths] == [64 * 255 * 12, 16 * 63 * 12, 1024 * 1016 * 12, 1024 * 1016 * 12]


def test_transmittance_outer_from_file(first_order_files):
    cfg, paths = first_order_files
    atm, _ = load_like_the_renderer(cfg, paths)
    pl = orc.planet(RADIUS, HEIGHT)
    sc = [orc.scatter(**orc.MIE), orc.scatter(**orc.RAYLEIGH)]
    rng = np.random.default_rng(5)
    errs = []
    for _ in range(200):
        hgt = rng.uniform(0, HEIGHT)
        sin_el = rng.uniform(0.02, 1.0)          # above the horizon, away from the grazing singularity
        point = np.array([0.0, 0.0, RADIUS + hgt])
        direction = np.array([np.sqrt(1 - sin_el ** 2), 0.0, sin_el])
        want = orc.transmittance(pl, sc, 100, point, direction, True)
        got = atm.transmittance_outer(point, direction)
        errs.append(np.max(np.abs(got - want) / want))
    errs = np.array(errs)
    assert np.median(errs) < 2e-3 and errs.max() < 5e-2      # interpolation error of the 64 x 255 table


def test_ray_scatter_outer_from_files(first_order_files):
    """ray_scatter_outer = ray-scatter + mie-strength * phase(0.76, mu) (ray-scatter-outer.glsl) against the direct
    first-order integral ray-scatter(point-scatter-base) of the reference."""
    cfg, paths = first_order_files
    atm, _ = load_like_the_renderer(cfg, paths)
    pl = orc.planet(RADIUS, HEIGHT)
    mie, ray = orc.scatter(**orc.MIE), orc.scatter(**orc.RAYLEIGH)
    one = (1.0, 1.0, 1.0)

    def base(p, v, l, ab):
        return orc.point_scatter_base(pl, [mie, ray], 100, one, p, v, l, ab)

    rng = np.random.default_rng(9)
    errs = []
    for _ in range(40):
        hgt = rng.uniform(100.0, 0.8 * HEIGHT)
        point = np.array([RADIUS + hgt, 0.0, 0.0])
        sin_el = rng.uniform(0.1, 0.95)
        az = rng.uniform(0, 2 * np.pi)
        direction = np.array([sin_el, np.sqrt(1 - sin_el ** 2) * np.cos(az), np.sqrt(1 - sin_el ** 2) * np.sin(az)])
        sun_el = rng.uniform(0.2, 0.95)
        light = np.array([sun_el, np.sqrt(1 - sun_el ** 2), 0.0])
        want = orc.ray_scatter(pl, [mie, ray], 100, base, point, direction, light, True)
        got = atm.ray_scatter_outer(light, point, direction)
        errs.append(np.max(np.abs(got - want) / want))
    errs = np.array(errs)
    # the reference's 4-D table is coarse (8 headings, 32 sun elevations): multilinear interpolation error is a few
    # per cent in the median and large next to the forward Mie peak -- the renderer sees exactly the same
    assert np.median(errs) < 5e-2 and np.percentile(errs, 90) < 0.5


def test_full_build_files_load_and_are_physical(tmp_path):
    """All five iterations: what the game ships.  Values sampled through the shader lookups are finite,
    non-negative, and multiple scattering adds light to the first-order table."""
    cfg = _lib.default_config()
    paths = atmosphere_lut.generate_atmosphere_luts(str(tmp_path), cfg=cfg)
    atm, surf = load_like_the_renderer(cfg, paths)
    cfg0 = _lib.make_config(iterations=0)
    first = atmosphere_lut.generate_tables(cfg=cfg0)
    # (texel-wise S_full >= S_first does not hold: every iteration re-tabulates S through forward o backward,
    # which moves values between texels where that map is not the identity)
    assert float(atm.ray_scatter.sum()) > 1.2 * float(first[2].sum()) and float(atm.ray_scatter.min()) >= 0.0
    np.testing.assert_array_equal(atm.mie_strength, first[3])           # first-order Mie is shipped separately
    assert float(surf.max()) > 0.05 and float(surf.min()) >= 0.0
    rng = np.random.default_rng(2)
    for _ in range(50):
        point = np.array([0.0, RADIUS + rng.uniform(0, HEIGHT), 0.0])
        d = rng.normal(size=3)
        d /= np.linalg.norm(d)
        d[1] = abs(d[1])
        light = np.array([0.0, 0.6, 0.8])
        s = atm.ray_scatter_outer(light, point, d)
        t = atm.transmittance_outer(point, d)
        assert np.all(np.isfinite(s)) and np.all(s >= 0) and np.all((t > 0) & (t <= 1 + 1e-6))
