"""The noise oracle (oracle/noise_oracle.c) against the reference's own facts: test/clj/sfsim/t_worley.clj:30-74 and
t_perlin.clj:69-146.  The random grids are inputs, as in the reference's with-redefs."""
import numpy as np
import pytest

from oracle import noise as orc


def v(*a):
    return [float(x) for x in a]


# t_worley.clj:30-40
def test_extract_point_from_grid():
    one = [[[v(1, 2, 3)]]]
    np.testing.assert_array_equal(orc.extract_point_from_grid(one, 10, 0, 0, 0), (1, 2, 3))
    np.testing.assert_array_equal(orc.extract_point_from_grid([[[v(1, 2, 3), v(4, 5, 6)]]], 10, 0, 0, 1), (4, 5, 6))
    np.testing.assert_array_equal(orc.extract_point_from_grid([[[v(1, 2, 3)], [v(4, 5, 6)]]], 10, 0, 1, 0), (4, 5, 6))
    np.testing.assert_array_equal(orc.extract_point_from_grid([[[v(1, 2, 3)]], [[v(4, 5, 6)]]], 10, 1, 0, 0), (4, 5, 6))
    np.testing.assert_array_equal(orc.extract_point_from_grid(one, 10, 0, 0, 1), (11, 2, 3))
    np.testing.assert_array_equal(orc.extract_point_from_grid(one, 10, 0, 1, 0), (1, 12, 3))
    np.testing.assert_array_equal(orc.extract_point_from_grid(one, 10, 1, 0, 0), (1, 2, 13))
    np.testing.assert_array_equal(orc.extract_point_from_grid(one, 10, 0, 0, -1), (-9, 2, 3))
    np.testing.assert_array_equal(orc.extract_point_from_grid(one, 10, 0, -1, 0), (1, -8, 3))
    np.testing.assert_array_equal(orc.extract_point_from_grid(one, 10, -1, 0, 0), (1, 2, -7))


# t_worley.clj:43-54
@pytest.mark.parametrize("grid,divisions,size,point,expected", [
    ([[[v(1, 1, 1)]]], 1, 2, (1, 1, 1), 0.0),
    ([[[v(1, 1, 1)]]], 1, 2, (0.5, 1, 1), 0.5),
    ([[[v(1, 1, 1), v(3, 1, 1)]]], 2, 4, (3, 1, 1), 0.0),
    ([[[v(1, 1, 1)], [v(1, 3, 1)]]], 2, 4, (1, 3, 1), 0.0),
    ([[[v(1, 1, 1)]], [[v(1, 1, 3)]]], 2, 4, (1, 1, 3), 0.0),
    ([[[v(0.25, 1, 1)]]], 1, 2, (1.75, 1, 1), 0.5),
    ([[[v(1.75, 1, 1)]]], 1, 2, (0.25, 1, 1), 0.5),
    ([[[v(1, 0.25, 1)]]], 1, 2, (1, 1.75, 1), 0.5),
    ([[[v(1, 1.75, 1)]]], 1, 2, (1, 0.25, 1), 0.5),
    ([[[v(1, 1, 0.25)]]], 1, 2, (1, 1, 1.75), 0.5),
    ([[[v(1, 1, 1.75)]]], 1, 2, (1, 1, 0.25), 0.5),
])
def test_closest_distance_to_point_in_grid(grid, divisions, size, point, expected):
    assert orc.closest_distance_to_point_in_grid(grid, divisions, size, point) == expected


# t_worley.clj:70-74
def test_worley_noise():
    noise = orc.worley_noise([[[v(0.5, 0.5, 0.5)]]], 2)
    assert noise[0] == 1.0 and len(noise) == 8 and noise.min() == 0.0


# t_perlin.clj:100-107
@pytest.mark.parametrize("t,expected", [(0.0, 0.0), (1.0, 1.0), (0.5, 0.5), (0.2, 0.05792), (0.8, 0.94208)])
def test_ease_curve(t, expected):
    assert orc.ease_curve(t) == pytest.approx(expected, abs=1e-4)


GRADIENT_GRID = [[[v(1, 1, 0), v(1, 0, -1)], [v(1, 0, 1), v(-1, -1, 0)]],
                 [[v(1, 0, 1), v(-1, 0, 1)], [v(1, 0, -1), v(-1, 0, -1)]]]    # t_perlin.clj:126-128


# t_perlin.clj:137-146
def test_perlin_noise_sample_and_noise():
    assert orc.perlin_noise_sample(GRADIENT_GRID, 2, 4, (0.5, 0.5, 0.5)) == pytest.approx(0.30273, abs=1e-5)
    assert orc.perlin_noise_sample(GRADIENT_GRID, 2, 4, (1.5, 0.5, 0.5)) == pytest.approx(-0.21457, abs=1e-5)
    noise = orc.perlin_noise(GRADIENT_GRID, 4)
    assert noise[0] == pytest.approx(0.74821, abs=1e-5)
    assert len(noise) == 64 and noise.min() == 0.0 and noise.max() == 1.0


# t_perlin.clj:69-97 (corner-gradients on the identity array, through the sample function): a grid whose gradient at
# [z][y][x] is (x, y, z) makes the wrap-around of the "+1" corner visible
def test_perlin_corner_wraparound():
    grid = [[[v(x, y, z) for x in range(4)] for y in range(4)] for z in range(4)]
    # at the centre of cell (3, 3, 3) the +1 corners wrap to index 0 (t_perlin.clj:89-96)
    got = orc.perlin_noise_sample(grid, 4, 4, (3.5, 3.5, 3.5))
    want = 0.0
    for z in (0, 1):
        for y in (0, 1):
            for x in (0, 1):
                g = np.array([(3 + x) % 4, (3 + y) % 4, (3 + z) % 4], dtype=float)
                corner = np.array([0.5 - x, 0.5 - y, 0.5 - z])
                want += 0.125 * float(g @ corner)          # ease-curve(1/2) = 1/2 on every axis
    assert got == pytest.approx(want, abs=1e-12)


# ---------------------------------------------------------------- blue noise, t_bluenoise.clj:48-130

def test_blue_noise_helpers():
    f1 = orc.density_function(1.0)
    assert f1(0, 0) == pytest.approx(1.0, abs=1e-6)
    for dx, dy in ((1, 0), (0, 1), (-1, 0), (0, -1)):
        assert f1(dx, dy) == pytest.approx(np.exp(-0.5))
    assert orc.density_function(2.0)(2, 0) == pytest.approx(np.exp(-0.5))
    assert orc.argmax_with_mask([5, 3, 2], [True] * 3) == 0
    assert orc.argmax_with_mask([3, 5, 2], [True] * 3) == 1
    assert orc.argmax_with_mask([3, 5, 2], [True, False, True]) == 0
    assert orc.argmax_with_mask([3, 2, 5], [True, False, True]) == 2
    assert orc.argmin_with_mask([2, 3, 5], [False] * 3) == 0
    assert orc.argmin_with_mask([3, 2, 5], [False] * 3) == 1
    assert orc.argmin_with_mask([3, 2, 5], [False, True, False]) == 0
    assert orc.argmin_with_mask([3, 5, 2], [False, True, False]) == 2
    assert [orc.wrap(x, m) for x, m in ((0, 5), (2, 5), (5, 5), (3, 5), (2, 4))] == [0, 2, 0, -2, -2]


def test_blue_noise_density():
    one, dx, dy = (lambda a, b: 1.0), (lambda a, b: float(a)), (lambda a, b: float(b))
    assert orc.density_sample([False] * 4, 2, one, 0, 0) == 0.0
    assert orc.density_sample([True, False, False, False], 2, one, 0, 0) == 1.0
    assert orc.density_sample([True, False, False, False], 2, lambda a, b: 2.0, 0, 0) == 2.0
    assert orc.density_sample([True] * 4, 2, one, 0, 0) == 4.0
    assert orc.density_sample([True] * 9, 3, dx, 1, 1) == 0.0
    assert orc.density_sample([True] * 9, 3, dy, 1, 1) == 0.0
    assert orc.density_sample([False, False, True, False], 2, dx, 1, 1) == -1.0
    assert orc.density_sample([False, True, False, False], 2, dy, 1, 1) == -1.0
    assert orc.density_sample([True] * 9, 3, dx, 2, 2) == 0.0
    assert orc.density_sample([True] * 9, 3, dy, 2, 2) == 0.0
    np.testing.assert_array_equal(orc.density_array([True, False, False, False], 2, dx), [0.0, -1.0, 0.0, -1.0])
    np.testing.assert_array_equal(orc.density_change([0.0, -1.0, 0.0, -1.0], 2, +1, dx, 0), [0.0, -2.0, 0.0, -2.0])
    np.testing.assert_array_equal(orc.density_change([0.0, -1.0, 0.0, -1.0], 2, -1, dx, 0), [0.0, 0.0, 0.0, 0.0])


def test_blue_noise_phases():
    f19, f15 = orc.density_function(1.9), orc.density_function(1.5)
    np.testing.assert_array_equal(orc.seed_pattern([True, False, False, False], 2, f19), [False, False, False, True])
    np.testing.assert_array_equal(orc.seed_pattern([True, True, False, False], 2, f19), [True, False, False, True])
    np.testing.assert_array_equal(orc.dither_phase1([True, False, False, False], 2, 1, f15), [0, 0, 0, 0])
    np.testing.assert_array_equal(orc.dither_phase1([True, False, False, True], 2, 2, f15), [0, 0, 0, 1])
    d, mask = orc.dither_phase2([True, False, False, True], 2, 2, [0, 0, 0, 1], f15)
    np.testing.assert_array_equal(d, [0, 0, 0, 1])
    np.testing.assert_array_equal(mask, [True, False, False, True])
    d, mask = orc.dither_phase2([True, False, False, False], 2, 1, [0, 0, 0, 0], f15)
    np.testing.assert_array_equal(d, [0, 0, 0, 1])
    np.testing.assert_array_equal(mask, [True, False, False, True])
    np.testing.assert_array_equal(orc.dither_phase3([True, False, False, True], 2, 2, [0, 0, 0, 1], f15), [0, 3, 2, 1])


def test_blue_noise_is_a_permutation_with_blue_spectrum():
    """blue-noise (bluenoise.clj:175-185): every rank 0 .. m^2 - 1 exactly once; thresholding at any level leaves no
    two set pixels closer than the seed pattern allows (weak check: low frequencies are suppressed)."""
    m, n = 16, 25
    rng = np.random.default_rng(3)
    picks = rng.permutation(m * m)[:n]
    dither = orc.blue_noise(m, picks, sigma=1.5)
    assert sorted(dither.tolist()) == list(range(m * m))
    spectrum = np.abs(np.fft.fft2(dither.reshape(m, m) - dither.mean())) ** 2
    low = spectrum[:3, :3].sum() - spectrum[0, 0]
    assert low < 0.02 * spectrum.sum()
