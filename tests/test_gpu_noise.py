"""3-D noise textures on the GPU (include/sfsim_noise.h) against the CPU oracle and the reference's own facts.

Integer-free but deterministic arithmetic (+, -, *, /, sqrt, floor in IEEE double, no fused multiply-add): the
un-normalised samples must equal the oracle's BIT FOR BIT, the float32 textures likewise."""
import os

import numpy as np
import pytest

from oracle import noise as orc
from sfsim_b200 import _lib, perlin, worley
from tests.test_noise_oracle import GRADIENT_GRID

pytestmark = pytest.mark.gpu


# t_worley.clj:70-74
def test_worley_noise_reference_facts():
    noise = worley.worley_noise(1, 2, grid=[[[(0.5, 0.5, 0.5)]]])
    assert noise.dtype == np.float32 and len(noise) == 8
    assert noise[0] == 1.0 and noise.min() == 0.0


# t_perlin.clj:137-146
def test_perlin_noise_reference_facts():
    samples = perlin.perlin_samples(GRADIENT_GRID, 4)
    assert samples[0] == pytest.approx(0.30273, abs=1e-5)          # cell (0.5, 0.5, 0.5)
    assert samples[1] == pytest.approx(-0.21457, abs=1e-5)         # cell (1.5, 0.5, 0.5): i runs fastest
    noise = perlin.perlin_noise(2, 4, gradients=GRADIENT_GRID)
    assert noise[0] == pytest.approx(0.74821, abs=1e-5)
    assert len(noise) == 64 and noise.min() == 0.0 and noise.max() == 1.0


@pytest.mark.parametrize("divisions,size", [(4, 16), (1, 3), (3, 12), (8, 64), (5, 5)])   # build.clj:36 default: 4, 16
def test_worley_matches_oracle_bit_for_bit(divisions, size):
    rng = np.random.default_rng(divisions * 100 + size)
    grid = worley.random_point_grid(divisions, size, random=lambda n: rng.random() * n)
    raw = worley.closest_distances(grid, size)
    want = orc.worley_noise(grid, size)
    np.testing.assert_array_equal(worley.worley_noise(divisions, size, grid=grid), want.astype(np.float32))
    # un-normalised distances: the oracle only exposes them point by point
    for t in (0, size ** 3 // 2, size ** 3 - 1):
        k, j, i = t // (size * size), (t // size) % size, t % size
        assert raw[t] == orc.closest_distance_to_point_in_grid(grid, divisions, size, (k + 0.5, j + 0.5, i + 0.5))
    assert raw.max() > 0 and np.all(raw >= 0)


@pytest.mark.parametrize("divisions,size", [(4, 16), (1, 2), (3, 10), (8, 64), (7, 7)])   # build.clj:42 default: 4, 16
def test_perlin_matches_oracle_bit_for_bit(divisions, size):
    rng = np.random.default_rng(divisions * 1000 + size)
    gradients = perlin.random_gradient_grid(divisions, lambda: perlin.random_gradient(lambda v: v[rng.integers(12)]))
    raw = perlin.perlin_samples(gradients, size)
    for t in (0, 1, size ** 3 // 3, size ** 3 - 1):
        k, j, i = t // (size * size), (t // size) % size, t % size
        assert raw[t] == orc.perlin_noise_sample(gradients, divisions, size, (i + 0.5, j + 0.5, k + 0.5))
    np.testing.assert_array_equal(perlin.perlin_noise(divisions, size, gradients=gradients),
                                  orc.perlin_noise(gradients, size).astype(np.float32))


def test_worley_texture_is_periodic_and_bounded():
    """size-independent properties at a larger size than the oracle test: values in [0, 1], the maximum distance maps
    to 0, and shifting the grid by one cell along x (the k axis of the texture, worley.clj:111) shifts the texture"""
    divisions, size = 8, 96
    rng = np.random.default_rng(5)
    grid = worley.random_point_grid(divisions, size, random=lambda n: rng.random() * n)
    noise = worley.worley_noise(divisions, size, grid=grid).reshape(size, size, size)
    assert noise.min() == 0.0 and noise.max() <= 1.0
    cell = size // divisions
    shifted = np.roll(grid, 1, axis=2).copy()                  # cell i -> i + 1 along x ...
    shifted[..., 0] = (shifted[..., 0] + cell) % size          # ... and its point with it
    other = worley.worley_noise(divisions, size, grid=shifted).reshape(size, size, size)
    np.testing.assert_allclose(other, np.roll(noise, cell, axis=0), rtol=0, atol=2e-6)


def test_invalid_arguments():
    with pytest.raises(_lib.AtmlutError):
        worley.worley_noise(3, 16, grid=np.zeros((3, 3, 3, 3)))        # size not a multiple of divisions
    with pytest.raises(TypeError):
        worley.worley_noise(2, 4, grid=np.zeros((2, 2, 3)))


# ---------------------------------------------------------------- blue noise (bluenoise.clj)

def test_blue_noise_matches_the_oracle_decision_for_decision():
    """Every arg-max / arg-min of the void-and-cluster phases must fall on the same element as in the CPU oracle (which
    reproduces t_bluenoise.clj:105-130): the dither arrays are equal as integers."""
    from sfsim_b200 import bluenoise
    for m, n, sigma, seed in ((16, 25, 1.5, 1), (2, 1, 1.5, 2), (9, 8, 1.9, 3), (64, 409, 1.5, 4)):   # 64: build.clj:47-50
        rng = np.random.default_rng(seed)
        picks = rng.permutation(m * m)[:n]
        got = bluenoise.blue_noise(m, n, sigma, picks=picks)
        want = orc.blue_noise(m, picks, sigma=sigma)
        np.testing.assert_array_equal(got, want)
        assert sorted(got.tolist()) == list(range(m * m))


def test_blue_noise_reference_facts():
    """t_bluenoise.clj:113-130 chained: seed [true false false true] of size 2 gives the dither array [0 3 2 1]"""
    from sfsim_b200 import bluenoise
    np.testing.assert_array_equal(bluenoise.blue_noise(2, 2, 1.5, picks=[0, 3]), [0, 3, 2, 1])
    tex = bluenoise.blue_noise_texture(2, 2, 1.5, picks=[0, 3])
    np.testing.assert_array_equal(tex, np.array([0, 3, 2, 1], dtype=np.float32) / 4)


def test_blue_noise_texture_of_the_build_task():
    from sfsim_b200 import bluenoise
    rng = np.random.default_rng(11)
    m = bluenoise.noise_size
    picks = rng.permutation(m * m)[:m * m // 10]
    tex = bluenoise.blue_noise_texture(picks=picks)
    want = orc.blue_noise(m, picks, sigma=1.5).astype(np.float64) / m / m
    np.testing.assert_array_equal(tex, want.astype(np.float32))
    spectrum = np.abs(np.fft.fft2(tex.reshape(m, m) - tex.mean())) ** 2
    assert spectrum[:4, :4].sum() - spectrum[0, 0] < 0.01 * spectrum.sum()      # low frequencies are suppressed


def test_build_tasks_write_the_files_of_build_clj(tmp_path):
    """sfsim_b200.build mirrors build.clj:34-51,84-87: file names, sizes and value ranges of the noise textures"""
    from sfsim_b200 import build
    clouds, data = str(tmp_path / "clouds"), str(tmp_path)
    build.worley(size=8, divisions=2, out_dir=clouds)
    build.perlin(size=8, divisions=2, out_dir=clouds)
    build.bluenoise(size=16, out_dir=data)
    for name in ("worley-north.raw", "worley-south.raw", "worley-cover.raw", "perlin.raw"):
        a = np.fromfile(os.path.join(clouds, name), dtype="<f4")
        assert a.size == 8 ** 3 and a.min() >= 0.0 and a.max() <= 1.0 and a.max() - a.min() > 0.5
    b = np.fromfile(os.path.join(data, "bluenoise.raw"), dtype="<f4")
    assert sorted(np.rint(b * 256).astype(int).tolist()) == list(range(256))        # every rank once, scaled by 1 / size^2
