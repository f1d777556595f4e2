"""The reference's own unit tests for the hot-path functions (test/clj/sfsim/t_atmosphere.clj,
t_interpolate.clj), run against the host-side mirror sfsim_b200.atmosphere / sfsim_b200.interpolate,
i.e. through the C ABI on the GPU.  Tolerances are the reference's unless noted.
"""
import math

import numpy as np
import pytest

from oracle import oracle as orc
from sfsim_b200 import atmosphere as atm
from sfsim_b200 import interpolate as itp
from tests import glsl_emulator as glsl

pytestmark = pytest.mark.gpu

E = math.e
PI = math.pi
radius = 6378000.0
max_height = 100000.0
earth = {"centre": (0, 0, 0), "radius": radius, "height": max_height, "brightness": (0.3, 0.3, 0.3)}
mie = {"base": (2e-5, 2e-5, 2e-5), "scale": 1200.0, "g": 0.76, "quotient": 0.9}
rayleigh = {"base": (5.8e-6, 13.5e-6, 33.1e-6), "scale": 8000.0}
scatter = [mie, rayleigh]


def roughly(v, tol):
    return pytest.approx(v, abs=tol, rel=0)


# t_atmosphere.clj:72-100
def test_scattering_extinction_phase():
    r = {"base": (5.8e-6,) * 3, "scale": 8000.0}
    assert atm.scattering(r, 0.0)[0] == 5.8e-6
    assert atm.scattering(r, 8000.0)[0] == roughly(5.8e-6 / E, 1e-12)
    assert atm.scattering(r, 16000.0)[0] == roughly(5.8e-6 / E / E, 1e-12)
    m = {"base": (2e-5,) * 3, "scale": 1200.0, "quotient": 0.9}
    assert atm.extinction(m, 1200.0)[0] == roughly(2e-5 / 0.9 / E, 1e-12)
    assert atm.phase({}, 0.0) == pytest.approx(3 / (16 * PI))
    assert atm.phase({}, 1.0) == pytest.approx(6 / (16 * PI))
    assert atm.phase({"g": 0.5}, 0.0) == pytest.approx((3 * 0.75) / (8 * PI * 2.25 * 1.25 ** 1.5))
    assert atm.phase({"g": 0.5}, 1.0) == pytest.approx((6 * 0.75) / (8 * PI * 2.25 * 0.25 ** 1.5))


def test_scattering_extinction_phase_on_the_device():
    """the same facts (t_atmosphere.clj:72-100) through the device functions the table kernels inline, and the device
    against the host mirror at random arguments"""
    r = {"base": (5.8e-6,) * 3, "scale": 8000.0}
    sc = atm.medium_batch(r, 0, [0.0, 8000.0, 16000.0])
    assert sc[0, 0] == 5.8e-6
    assert sc[1, 0] == roughly(5.8e-6 / E, 1e-12) and sc[2, 0] == roughly(5.8e-6 / E / E, 1e-12)
    m = {"base": (2e-5,) * 3, "scale": 1200.0, "quotient": 0.9}
    assert atm.medium_batch(m, 1, [1200.0])[0, 0] == roughly(2e-5 / 0.9 / E, 1e-12)
    assert atm.medium_batch({"base": (1, 1, 1), "scale": 1.0}, 2, [0.0, 1.0]).tolist() == pytest.approx([3 / (16 * PI), 6 / (16 * PI)])
    half = {"base": (1, 1, 1), "scale": 1.0, "g": 0.5}
    assert atm.medium_batch(half, 2, [0.0, 1.0]).tolist() == pytest.approx(
        [(3 * 0.75) / (8 * PI * 2.25 * 1.25 ** 1.5), (6 * 0.75) / (8 * PI * 2.25 * 0.25 ** 1.5)])
    rng = np.random.default_rng(3)
    hs, mus = rng.uniform(0, 1e5, 200), rng.uniform(-1, 1, 200)
    np.testing.assert_allclose(atm.medium_batch(mie, 0, hs), [atm.scattering(mie, h) for h in hs], rtol=1e-14)
    np.testing.assert_allclose(atm.medium_batch(mie, 1, hs), [atm.extinction(mie, h) for h in hs], rtol=1e-14)
    np.testing.assert_allclose(atm.medium_batch(mie, 2, mus), [atm.phase(mie, mu) for mu in mus], rtol=1e-13)


# t_atmosphere.clj:144-172
def test_transmittance_known_answers():
    r = {"base": (5.8e-6, 13.5e-6, 33.1e-6), "scale": 8000.0}
    m = {"base": (2e-5,) * 3, "scale": 1200.0, "quotient": 0.9}
    both = [r, m]
    t = atm.transmittance
    assert t(earth, [r], 50, (0, radius, 0), (0, radius, 0))[0] == roughly(1.0, 1e-6)
    assert t(earth, [r], 50, (0, radius, 0), (1000, radius, 0))[0] == roughly(math.exp(-1000 * 5.8e-6), 1e-6)
    assert (t(earth, [r], 50, (0, radius + 8000, 0), (1000, radius + 8000, 0))[0] ==
            roughly(math.exp(-(1000 * 5.8e-6) / E), 1e-6))
    assert t(earth, both, 50, (0, radius, 0), (1000, radius, 0))[0] == roughly(math.exp(-1000 * (5.8e-6 + 2e-5 / 0.9)), 1e-6)
    assert t(earth, [r], 50, (-1000, radius, 0), (1, 0, 0), False)[0] == roughly(math.exp(-1000 * 5.8e-6), 1e-6)
    assert t(earth, both, 50, (0, radius, 0), (0, 1, 0), True)[0] == roughly(0.932307, 1e-6)


def test_transmittance_batch_matches_oracle():
    rng = np.random.default_rng(3)
    n = 200
    pl = orc.planet(radius, max_height)
    om, orr = orc.scatter(**orc.MIE), orc.scatter(**orc.RAYLEIGH)
    x = rng.normal(size=(n, 3))
    x = x / np.linalg.norm(x, axis=1, keepdims=True) * (radius + rng.uniform(0, max_height, size=(n, 1)))
    v = rng.normal(size=(n, 3))
    v /= np.linalg.norm(v, axis=1, keepdims=True)
    above = np.array([orc.is_above_horizon(pl, x[i], v[i]) for i in range(n)])
    got = atm.transmittance_batch(earth, scatter, 30, x, v, above)
    want = np.array([orc.transmittance(pl, [om, orr], 30, x[i], v[i], above[i]) for i in range(n)])
    np.testing.assert_allclose(got, want, rtol=1e-9, atol=1e-300)


# t_atmosphere.clj:175-194 (transmittance is mocked to 0.5 there; with scatter = [] it is 1)
def test_surface_radiance_base():
    moved = dict(earth, centre=(0, 2 * radius, 0))
    one = (1.0, 1.0, 1.0)
    np.testing.assert_allclose(atm.surface_radiance_base(earth, [], 10, one, (0, radius, 0), (1, 0, 0)), 0.0, atol=0)
    np.testing.assert_allclose(atm.surface_radiance_base(moved, [], 10, one, (0, radius, 0), (0, -1, 0)), 1.0)
    np.testing.assert_allclose(atm.surface_radiance_base(earth, [], 10, one, (0, radius, 0), (0, 1, 0)), 1.0)
    np.testing.assert_allclose(atm.surface_radiance_base(earth, [], 10, one, (0, radius, 0), (0, -1, 0)), 0.0, atol=0)


# t_atmosphere.clj:285-296
def test_scattering_components_add_up():
    steps, one = 100, (1, 1, 1)
    x, v, l = (radius + 1000, 0, 0), (0, 1, 0), (0.36, 0.48, 0.8)
    mu = float(np.dot(v, l))
    a = atm.point_scatter_component(earth, scatter, mie, steps, one, x, v, l, True)
    b = atm.point_scatter_component(earth, scatter, rayleigh, steps, one, x, v, l, True)
    base = atm.point_scatter_base(earth, scatter, steps, one, x, v, l, True)
    np.testing.assert_allclose(a + b, base, atol=1e-12, rtol=0)
    s = atm.strength_component(earth, scatter, mie, steps, one, x, v, l, True)
    np.testing.assert_allclose(s * atm.phase(mie, mu), a, atol=1e-12, rtol=0)
    pl = orc.planet(radius, max_height)
    want = orc.point_scatter_base(pl, [orc.scatter(**orc.MIE), orc.scatter(**orc.RAYLEIGH)], steps, one, x, v, l, True)
    np.testing.assert_allclose(base, want, rtol=1e-9)
    # sun below the horizon: no direct light (atmosphere.clj:154-160)
    np.testing.assert_array_equal(atm.point_scatter_base(earth, scatter, steps, one, (radius, 0, 0), v, (-1, 0, 0)), 0)


def test_ray_scatter_of_first_order_source_matches_oracle():
    one = (1, 1, 1)
    pl = orc.planet(radius, max_height)
    om, orr = orc.scatter(**orc.MIE), orc.scatter(**orc.RAYLEIGH)
    src = atm.FirstOrder(atm.FirstOrder.COMPONENT, earth, scatter, rayleigh, 10, one)
    x, l = (0, 0, radius + 2000.0), (0.6, 0, 0.8)
    for v, above in [((0, 0, 1), True), ((0.6, 0.8, 0), True), ((0, 0.6, -0.8), False)]:
        got = atm.ray_scatter(earth, scatter, 10, src, x, v, l, above)

        def j(p, vv, ll, ab):
            return orc.point_scatter_component(pl, [om, orr], orr, 10, one, p, vv, ll, ab)

        want = orc.ray_scatter(pl, [om, orr], 10, j, x, v, l, above)
        np.testing.assert_allclose(got, want, rtol=1e-9)


# t_atmosphere.clj:386-459, 505-536: index maps
def test_index_maps():
    small = {"radius": 4.0, "height": 1.0}
    assert atm.horizon_distance(small, 4.0) == 0.0
    assert atm.horizon_distance(small, 5.0) == 3.0
    cases = [(2, (4, 0, 0), (-1, 0, 0), False, 0.5), (2, (5, 0, 0), (-1, 0, 0), False, 1 / 3),
             (2, (5, 0, 0), (-math.sqrt(0.5), math.sqrt(0.5), 0), False, 0.223), (3, (4, 0, 0), (-1, 0, 0), False, 1.0),
             (2, (5, 0, 0), (-0.6, 0.8, 0), False, 0.0), (2, (4, 0, 0), (1, 0, 0), True, 2 / 3),
             (2, (5, 0, 0), (0, 1, 0), True, 0.5), (2, (5, 0, 0), (-0.6, 0.8, 0), True, 1.0),
             (2, (4, 0, 0), (0, 1, 0), True, 1.0), (3, (4, 0, 0), (0, 1, 0), True, 2.0),
             (2, (5, 0, 0), (-1, 0, 0), True, 1.0), (2, (4, 0, 0), (1, 0, 0), False, 0.5)]
    for size, point, direction, above, expected in cases:
        assert atm.elevation_to_index(small, size, point, direction, above) == roughly(expected, 1e-3)
    inv = [(2, 5.0, 1 / 3, (-1, 0, 0), False), (3, 5.0, 2 / 3, (-1, 0, 0), None),
           (2, 5.0, 0.222549, (-math.sqrt(0.5), math.sqrt(0.5), 0), None), (2, 5.0, 0.4, (-1, 0, 0), None),
           (2, 4.0, 0.4, (0, 1, 0), None), (2, 4.0, 2 / 3, (1, 0, 0), True), (3, 4.0, 4 / 3, (1, 0, 0), None),
           (2, 4.0, 1.0, (0, 1, 0), None), (2, 5.0, 1.0, (-0.6, 0.8, 0), None), (2, 5.0, 0.5, (0, 1, 0), True),
           (2, 5.0, 0.5001, (0, 1, 0), None), (2, 4.0, 0.5, (0, 1, 0), False), (2, 4.0, 0.5001, (1, 0, 0), None)]
    for size, r, index, direction, above in inv:
        d, a = atm.index_to_elevation(small, size, r, index)
        np.testing.assert_allclose(d, direction, atol=1e-3)
        if above is not None:
            assert a is above
    assert atm.height_to_index(small, 2, (4, 0, 0)) == 0.0
    assert atm.height_to_index(small, 2, (5, 0, 0)) == 1.0
    assert atm.height_to_index(small, 2, (4.5, 0, 0)) == roughly(0.687, 1e-3)
    assert atm.height_to_index(small, 17, (5, 0, 0)) == 16.0
    e35 = {"radius": radius, "height": 35000.0}
    assert atm.height_to_index(e35, 32, (6377999.999549146, -16.87508805500576, 73.93459155883768)) == roughly(0, 1e-6)
    np.testing.assert_allclose(atm.index_to_height(small, 2, 0.0), (4, 0, 0))
    np.testing.assert_allclose(atm.index_to_height(small, 2, 0.68718), (4.5, 0, 0), atol=1e-3)
    np.testing.assert_allclose(atm.index_to_height(small, 3, 2.0), (5, 0, 0))
    assert atm.sun_elevation_to_index(2, (4, 0, 0), (1, 0, 0)) == 1.0
    assert atm.sun_elevation_to_index(2, (4, 0, 0), (0, 1, 0)) == roughly(0.464, 1e-3)
    assert atm.sun_elevation_to_index(2, (4, 0, 0), (-1, 0, 0)) == 0.0
    assert atm.sun_elevation_to_index(17, (4, 0, 0), (1, 0, 0)) == 16.0
    assert atm.index_to_sin_sun_elevation(2, 1.0) == roughly(1.0, 1e-3)
    assert atm.index_to_sin_sun_elevation(2, 0.0) == roughly(-0.2, 1e-3)
    assert atm.index_to_sin_sun_elevation(2, 0.463863) == roughly(0.0, 1e-3)
    assert atm.index_to_sin_sun_elevation(3, 1.0) == roughly(0.022, 1e-3)
    assert atm.sun_angle_to_index(2, (0, 1, 0), (0, 1, 0)) == 1.0
    assert atm.sun_angle_to_index(2, (0, 1, 0), (0, -1, 0)) == 0.0
    assert atm.sun_angle_to_index(2, (0, 1, 0), (0, 0, 1)) == 0.5
    assert atm.sun_angle_to_index(17, (0, 1, 0), (1, 0, 0)) == 8.0
    sd = atm.index_to_sun_direction
    np.testing.assert_allclose(sd(2, (0, 1, 0), 0.0, 1.0), (0, 1, 0))
    np.testing.assert_allclose(sd(2, (0, 1, 0), 0.0, 0.0), (0, -1, 0))
    np.testing.assert_allclose(sd(2, (0, 1, 0), 1.0, 0.5), (1, 0, 0))
    np.testing.assert_allclose(sd(2, (0, 1, 0), 1.00001, 0.5), (1, 0, 0), atol=1e-3)
    np.testing.assert_allclose(sd(2, (0, 1, 0), 0.0, 0.5), (0, 0, 1))
    np.testing.assert_allclose(sd(2, (1, 0, 0), 1.0, 1.0), (1, 0, 0))
    np.testing.assert_allclose(sd(2, (0, -1, 0), 0.0, 1.0), (0, -1, 0))
    np.testing.assert_allclose(sd(3, (0, 1, 0), 0.0, 1.0), (0, 0, 1))


# t_atmosphere.clj:462-502, 539-566: interpolation spaces
def test_spaces():
    e35 = {"radius": radius, "height": 35000.0}
    space = atm.transmittance_space(e35, [15, 17])
    assert space.shape == (15, 17)
    np.testing.assert_allclose(space.forward((radius, 0, 0), (0, 1, 0), True), (0, 16))
    np.testing.assert_allclose(space.forward((radius + 35000, 0, 0), (0, 1, 0), True), (14, 8), atol=1e-9)
    np.testing.assert_allclose(space.forward((radius, 0, 0), (-1, 0, 0), False), (0, 8))
    p, d, a = space.backward(0.0, 16.0)
    np.testing.assert_allclose(p, (radius, 0, 0))
    np.testing.assert_allclose(d, (0, 1, 0), atol=1e-6)
    assert a is True
    assert space.backward(0.0, 8.0)[2] is False
    sr = atm.surface_radiance_space(e35, [15, 17])
    np.testing.assert_allclose(sr.forward((radius, 0, 0), (1, 0, 0)), (0, 16))
    np.testing.assert_allclose(sr.forward((radius, 0, 0), (-1, 0, 0)), (0, 0))
    np.testing.assert_allclose(sr.forward((radius, 0, 0), (0, 1, 0)), (0, 7.422), atol=1e-3)
    np.testing.assert_allclose(sr.backward(14.0, 0.0)[1], (-0.2, 0.980, 0), atol=1e-3)
    rs = atm.ray_scatter_space(earth, [21, 19, 17, 15])
    np.testing.assert_allclose(rs.forward((radius, 0, 0), (1, 0, 0), (1, 0, 0), True), (0, 9.794, 16, 14), atol=1e-3)
    np.testing.assert_allclose(rs.forward((radius + max_height, 0, 0), (1, 0, 0), (1, 0, 0), True), (20, 9, 16, 14),
                               atol=1e-9)
    np.testing.assert_allclose(rs.forward((0, radius, 0), (0, -1, 0), (0, 1, 0), False), (0, 9, 16, 0))
    np.testing.assert_allclose(rs.forward((radius, 0, 0), (0, 0, 1), (0, 0, -1), True), (0, 18, 7.422, 0), atol=1e-3)
    np.testing.assert_allclose(rs.backward(0.0, 9.7937607, 7.421805, 7.0)[2], (0, 0, 1), atol=1e-3)
    assert rs.backward(0.0, 9.79376, 0.0, 0.0)[3] is True
    assert rs.backward(20.0, 8.206, 16.0, 0.0)[3] is False
    # every integer texel of a small space agrees with the oracle bit for bit
    pl = orc.planet(radius, max_height)
    for idx in [(0, 0, 0, 0), (3, 9, 5, 2), (20, 18, 16, 14), (7, 9, 0, 7), (0, 9, 3, 14)]:
        got = rs.backward(*[float(i) for i in idx])
        want = orc.ray_scatter_backward(pl, (21, 19, 17, 15), *[float(i) for i in idx])
        np.testing.assert_array_equal(got[0], want[0])      # sqrt and arithmetic only: bit-exact
        np.testing.assert_array_equal(got[1], want[1])
        np.testing.assert_allclose(got[2], want[2], rtol=0, atol=1e-15)   # log/exp differ by an ulp between libms
        assert got[3] == want[3]
        np.testing.assert_allclose(rs.forward(*got), orc.ray_scatter_forward(pl, (21, 19, 17, 15), *want), rtol=0,
                                   atol=1e-12)


# t_interpolate.clj:22-117
def test_interpolate_module():
    sp = itp.linear_space([-2.0], [4.0], [16])
    assert sp.forward(-2.0) == [0.0] and sp.forward(4.0) == [15.0] and sp.forward(0.0) == [5.0]
    assert sp.backward(0.0) == [-2.0] and sp.backward(15.0) == [4.0] and sp.backward(5.0) == [0.0]
    assert itp.linear_space([-2.0, -1.0], [4.0, 1.0], [16, 5]).forward(4.0, 0.0) == [15.0, 2.0]
    sqr = lambda x: x * x
    np.testing.assert_array_equal(itp.make_lookup_table(sqr, itp.linear_space([-3.0], [2.0], [6])), [9, 4, 1, 0, 1, 4])
    np.testing.assert_array_equal(itp.make_lookup_table(lambda a, b: a * b, itp.linear_space([1.0, 3.0], [2.0, 5.0], [2, 3])),
                                  [[3, 4, 5], [6, 8, 10]])
    assert [itp.clip(v, 16) for v in (0.0, 15.0, -2.0, 16.0)] == [0.0, 15.0, 0.0, 15.0]
    assert [itp.mix(-2.0, 4.0, s) for s in (0.0, 1.0, 0.5, 0.25)] == [-2.0, 4.0, 1.0, -0.5]
    table = itp.interpolation_table([9, 4, 1, 0, 1, 4], itp.linear_space([-3.0], [2.0], [6]))
    assert [table(x) for x in (-3.0, 1.5, -5.0, 3.0)] == [9.0, 2.5, 9.0, 4.0]
    vec = itp.interpolation_table([[2, 3, 5], [3, 5, 9]], itp.linear_space([-1.0], [1.0], [2]))
    np.testing.assert_array_equal(vec(0.0), (2.5, 4.0, 7.0))
    t2 = itp.interpolation_table([[2, 3, 5], [7, 11, 13]], itp.linear_space([0.0, 0.0], [1.0, 2.0], [2, 3]))
    assert [t2(0, 0), t2(0, 2), t2(0, 1.5), t2(1, 0), t2(0.5, 0)] == [2.0, 5.0, 4.0, 7.0, 4.5]
    f = itp.interpolate_function(sqr, itp.linear_space([-3.0], [2.0], [6]))
    assert [f(x) for x in (-3.0, 1.5, -5.0, 3.0)] == [9.0, 2.5, 9.0, 4.0]
    radius_space = type("R", (), {"shape": None, "forward": staticmethod(lambda a, b: [math.hypot(a, b)]),
                                  "backward": staticmethod(lambda r: [r, 0.0])})
    combined = itp.compose_space(itp.linear_space([0.0], [1.0], [101]), radius_space)
    assert combined.forward(3.0, 4.0) == [500.0]
    assert combined.backward(500.0) == [5.0, 0.0]


# t_atmosphere.clj:569-578, 629-635, 661-667, 769-775: LUTs built by the CUDA kernels, sampled like the GLSL
def test_luts_through_the_shader_lookups():
    size, steps = 12, 10
    one = (1, 1, 1)
    t_space = atm.transmittance_space(earth, [size, size])
    rs_space = atm.ray_scatter_space(earth, [size] * 4)
    T = itp.make_lookup_table(itp.Transmittance(earth, scatter, steps), t_space)
    point_scatter_rayleigh = atm.FirstOrder(atm.FirstOrder.COMPONENT, earth, scatter, rayleigh, steps, one)
    S = itp.make_lookup_table(itp.RayScatter(earth, scatter, steps, point_scatter_rayleigh), rs_space)
    M = S                                                         # t_atmosphere.clj:574
    tiled = orc.convert_4d_to_2d(S.astype(np.float64)).astype(np.float32)
    a = glsl.Atmosphere(radius, max_height, T, tiled, tiled, (size,) * 4)
    assert a.transmittance_track((0, 0, 6478000), (0, 0, 6478000))[0] == roughly(1.0, 1e-4)
    assert a.transmittance_track((0, 0, 6378000), (0, 0, 6478000))[0] == roughly(0.976549, 1e-4)
    assert a.transmittance_track((6378000, 0, 0), (6378000, 0, 100000))[0] == roughly(0.079658, 1e-4)
    assert a.transmittance_outer((0, 0, 6478000), (0, 0, 1))[0] == roughly(0.976359, 1e-4)
    assert a.transmittance_outer((0, 0, 6378000), (0, 0, 1))[0] == roughly(0.953463, 1e-4)
    assert a.transmittance_outer((0, 0, 6378000), (1, 0, 0))[0] == roughly(0.016916, 1e-4)
    assert a.ray_scatter_track((0, 0, 1), (0, 0, 6378000), (0, 0, 6378000))[2] == roughly(0.0, 1e-4)
    assert a.ray_scatter_track((0, 0, 1), (0, 0, 6378000), (0, 0, 6478000))[2] == roughly(0.043302, 1e-4)
    assert a.ray_scatter_track((0, 0, 1), (0, 0, 6378000), (100000, 0, 6378000))[2] == roughly(0.008272, 1e-4)
    # the interpolated closure agrees with the direct evaluation at a texel centre
    t_fn = itp.interpolation_table(T, t_space)
    x, v, above = t_space.backward(3.0, 9.0)
    np.testing.assert_allclose(t_fn(x, v, above), atm.transmittance(earth, scatter, steps, x, v, above), rtol=1e-6)


def test_make_lookup_table_chain_matches_oracle():
    """One iteration of atmosphere_lut.clj:74-97 written with the mirrored public functions."""
    planet = {"centre": (0, 0, 0), "radius": radius, "height": 35000.0, "brightness": (0.3, 0.3, 0.3)}
    shape4, shape_e, steps, sphere_steps, one = (4, 9, 4, 2), (3, 7), 20, 8, (1, 1, 1)
    rs_space = atm.ray_scatter_space(planet, shape4)
    e_space = atm.surface_radiance_space(planet, shape_e)
    dE = itp.interpolate_function(itp.SurfaceRadianceBase(planet, scatter, steps, one), e_space)
    first_rayleigh = itp.interpolate_function(
        itp.RayScatter(planet, scatter, steps, atm.FirstOrder(atm.FirstOrder.COMPONENT, planet, scatter, rayleigh, steps, one)),
        rs_space)
    first_mie = itp.interpolate_function(
        itp.RayScatter(planet, scatter, steps, atm.FirstOrder(atm.FirstOrder.STRENGTH, planet, scatter, mie, steps, one)),
        rs_space)
    dS = itp.MieCombined(first_rayleigh, first_mie, mie, scatter)
    dJ = itp.interpolate_function(itp.PointScatter(planet, scatter, dS, dE, one, sphere_steps, steps), rs_space)
    dE2 = itp.interpolate_function(itp.SurfaceRadiance(planet, dS, steps), e_space)
    dS2 = itp.interpolate_function(itp.RayScatter(planet, scatter, steps, dJ), rs_space)
    S = itp.make_lookup_table(itp.TableSum(first_rayleigh, dS2), rs_space)
    pl = orc.planet(radius, 35000.0)
    cfg = orc.config(shape4, (2, 2), shape_e, steps, sphere_steps)
    rec = {}
    orc.generate_atmosphere_luts(pl, orc.scatter(**orc.MIE), orc.scatter(**orc.RAYLEIGH), cfg, iterations=1, record=rec)

    def err(got, want):
        got, want = np.asarray(got, np.float64), np.asarray(want, np.float64)
        return float(np.max(np.abs(got - want) / np.maximum(np.abs(want), 1e-20)))

    assert err(dJ.table, rec["dJ0"]) <= 1e-4
    assert err(dE2.table, rec["dE0"]) <= 1e-4
    assert err(dS2.table, rec["dS0"]) <= 1e-4
    assert err(S, rec["S0"]) <= 1e-4


def test_integrals_with_table_sources_at_arbitrary_points():
    """point-scatter, surface-radiance and ray-scatter with interpolated tables as their function arguments
    (atmosphere.clj:192-230), at points that are NOT texel centres and not in the (r, 0, 0) frame."""
    planet = {"centre": (0, 0, 0), "radius": radius, "height": 35000.0, "brightness": (0.3, 0.3, 0.3)}
    shape4, shape_e, steps, sphere_steps, one = (4, 9, 4, 2), (3, 7), 12, 8, (1, 1, 1)
    pl = orc.planet(radius, 35000.0)
    om, orr = orc.scatter(**orc.MIE), orc.scatter(**orc.RAYLEIGH)
    cfg = orc.config(shape4, (2, 2), shape_e, steps, sphere_steps)
    rec = {}
    orc.generate_atmosphere_luts(pl, om, orr, cfg, iterations=1, record=rec)
    rs_space = atm.ray_scatter_space(planet, shape4)
    e_space = atm.surface_radiance_space(planet, shape_e)
    r1 = itp.interpolation_table(rec["R1"], rs_space)
    m1 = itp.interpolation_table(rec["M1"], rs_space)
    e0 = itp.interpolation_table(rec["Ebase"], e_space)
    dj = itp.interpolation_table(rec["dJ0"], rs_space)
    ds_closure = itp.MieCombined(r1, m1, mie, scatter)
    osrc = orc.SSourceSpec(r1.table.astype(np.float64), m1.table.astype(np.float64), om)

    def o_lookup4(table):
        t = table.astype(np.float64)
        return lambda p, v, l, ab: orc.interpolate(t, orc.ray_scatter_forward(pl, shape4, p, v, l, ab))

    def o_ds(p, v, l, ab):
        return o_lookup4(r1.table)(p, v, l, ab) + o_lookup4(m1.table)(p, v, l, ab) * orc.phase(om, float(np.dot(v, l)))

    def o_e(p, l):
        return orc.interpolate(e0.table.astype(np.float64), orc.surface_radiance_forward(pl, shape_e, p, l))

    x = np.array([0.3, 0.5, 0.81]) / np.linalg.norm([0.3, 0.5, 0.81]) * (radius + 4321.0)
    l = np.array([0.48, 0.6, 0.64])
    for v in (np.array([0.0, 0.6, 0.8]), np.array([0.6, -0.8, 0.0]), -x / np.linalg.norm(x)):
        got = atm.point_scatter(planet, scatter, ds_closure, e0, one, sphere_steps, steps, x, v, l, True)
        want = orc.point_scatter(pl, [om, orr], o_ds, o_e, one, sphere_steps, steps, x, v, l, True)
        np.testing.assert_allclose(got, want, rtol=1e-9)
        above = orc.is_above_horizon(pl, x, v)
        got = atm.ray_scatter(planet, scatter, steps, dj, x, v, l, above)
        want = orc.ray_scatter(pl, [om, orr], steps, o_lookup4(dj.table), x, v, l, above)
        np.testing.assert_allclose(got, want, rtol=1e-9)
    got = atm.surface_radiance(planet, ds_closure, steps, x, l)
    want = orc.surface_radiance(pl, o_ds, steps, x, l)
    np.testing.assert_allclose(got, want, rtol=1e-9)
    got = atm.surface_radiance(planet, dj, steps, x, l)
    want = orc.surface_radiance(pl, o_lookup4(dj.table), steps, x, l)
    np.testing.assert_allclose(got, want, rtol=1e-9)
    assert osrc is not None
