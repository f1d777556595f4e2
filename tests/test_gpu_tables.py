"""Parity of the CUDA table kernels (through the C ABI) against the CPU oracle and the golden fixtures.

Tolerance (BASELINE.json north_star): max relative error <= 1e-4 per LUT entry.  Entries are compared as
|gpu - ref| / max(|ref|, FLOOR) with FLOOR = 1e-20: the tables hold radiances up to ~1, and reference
entries below 1e-20 are themselves rounding noise of the double-precision reference (e.g. a lookup weight
of one ulp on a neighbour of an all-zero row), so they are compared absolutely.
"""
import ctypes as C
import os

import numpy as np
import pytest

from oracle import oracle as orc
from sfsim_b200 import _lib, atmosphere_lut

pytestmark = pytest.mark.gpu

TOL = 1e-4
FLOOR = 1e-20
GOLDEN = os.path.join(os.path.dirname(__file__), "golden")

REDUCED = dict(shape4=(8, 31, 8, 2), shape_t=(16, 63), shape_e=(4, 15), ray_steps=100, sphere_steps=15)


def rel_err(gpu, ref):
    gpu = np.asarray(gpu, dtype=np.float64).reshape(-1)
    ref = np.asarray(ref, dtype=np.float64).reshape(-1)
    assert gpu.shape == ref.shape
    assert np.isfinite(gpu).all()
    return float(np.max(np.abs(gpu - ref) / np.maximum(np.abs(ref), FLOOR)))


def lib_config(c, iterations=5, **kw):
    return _lib.make_config(ray_scatter_shape=c["shape4"], transmittance_shape=c["shape_t"],
                            surface_radiance_shape=c["shape_e"], ray_steps=c["ray_steps"],
                            sphere_steps=c["sphere_steps"], iterations=iterations, **kw)


def orc_config(c, **kw):
    return orc.config(c["shape4"], c["shape_t"], c["shape_e"], c["ray_steps"], c["sphere_steps"], **kw)


class Lib:
    """Thin caller of the per-table entry points with Earth defaults."""

    def __init__(self, cfg, planet=atmosphere_lut.earth, scatter=(atmosphere_lut.mie, atmosphere_lut.rayleigh)):
        self.lib = _lib.load()
        self.cfg = cfg
        self.pl = _lib.make_planet(planet)
        self.sc = _lib.make_scatter_array(scatter)
        self.n = len(scatter)
        self.s4 = cfg.ray_scatter_shape + (3,)
        self.st = cfg.transmittance_shape + (3,)
        self.se = cfg.surface_radiance_shape + (3,)

    def transmittance(self):
        out = np.zeros(self.st, np.float32)
        _lib.check(self.lib.atmlut_transmittance_table(C.byref(self.pl), self.sc, self.n, C.byref(self.cfg),
                                                       _lib.ptr(out)))
        return out

    def surface_radiance_base(self):
        out = np.zeros(self.se, np.float32)
        _lib.check(self.lib.atmlut_surface_radiance_base_table(C.byref(self.pl), self.sc, self.n, C.byref(self.cfg),
                                                               _lib.ptr(out)))
        return out

    def first_order(self, comp_a=1, strength_a=0, comp_b=0, strength_b=1):
        a, b = np.zeros(self.s4, np.float32), np.zeros(self.s4, np.float32)
        _lib.check(self.lib.atmlut_first_order_tables(C.byref(self.pl), self.sc, self.n, C.byref(self.cfg), comp_a,
                                                      strength_a, _lib.ptr(a), comp_b, strength_b, _lib.ptr(b)))
        return a, b

    def point_scatter(self, ds_a, ds_b, de):
        out = np.zeros(self.s4, np.float32)
        ds_a, de = _lib.f32(ds_a), _lib.f32(de)
        ds_b = _lib.f32(ds_b) if ds_b is not None else None
        _lib.check(self.lib.atmlut_point_scatter_table(C.byref(self.pl), self.sc, self.n, C.byref(self.cfg),
                                                       _lib.ptr(ds_a), _lib.ptr(ds_b), 0, _lib.ptr(de),
                                                       _lib.ptr(out)))
        return out

    def surface_radiance(self, ds_a, ds_b):
        out = np.zeros(self.se, np.float32)
        ds_a = _lib.f32(ds_a)
        ds_b = _lib.f32(ds_b) if ds_b is not None else None
        _lib.check(self.lib.atmlut_surface_radiance_table(C.byref(self.pl), self.sc, self.n, C.byref(self.cfg),
                                                          _lib.ptr(ds_a), _lib.ptr(ds_b), 0, _lib.ptr(out)))
        return out

    def ray_scatter(self, dj):
        out = np.zeros(self.s4, np.float32)
        dj = _lib.f32(dj)
        _lib.check(self.lib.atmlut_ray_scatter_table(C.byref(self.pl), self.sc, self.n, C.byref(self.cfg),
                                                     _lib.ptr(dj), _lib.ptr(out)))
        return out

    def resample(self, which, a, b=None):
        shape = {0: self.s4, 1: self.se, 2: self.st}[which]
        out = np.zeros(shape, np.float32)
        a = _lib.f32(a) if a is not None else None
        b = _lib.f32(b) if b is not None else None
        _lib.check(self.lib.atmlut_resample_table(C.byref(self.pl), C.byref(self.cfg), which, _lib.ptr(a),
                                                  _lib.ptr(b), _lib.ptr(out)))
        return out


@pytest.fixture(scope="module")
def reduced_oracle():
    """All intermediate tables of BASELINE.json configs[0] from the oracle (about a second on the host)."""
    pl = orc.planet(**orc.EARTH)
    rec = {}
    files = orc.generate_atmosphere_luts(pl, orc.scatter(**orc.MIE), orc.scatter(**orc.RAYLEIGH),
                                         orc_config(REDUCED), iterations=3, record=rec)
    return rec, files


# ---------------------------------------------------------------- stage by stage, oracle inputs

def test_transmittance_and_surface_radiance_base_tables(reduced_oracle):
    rec, _ = reduced_oracle
    lib = Lib(lib_config(REDUCED))
    assert rel_err(lib.transmittance(), rec["T"]) <= TOL
    assert rel_err(lib.surface_radiance_base(), rec["Ebase"]) <= TOL


def test_first_order_tables(reduced_oracle):
    rec, _ = reduced_oracle
    lib = Lib(lib_config(REDUCED))
    r1, m1 = lib.first_order()
    assert rel_err(r1, rec["R1"]) <= TOL
    assert rel_err(m1, rec["M1"]) <= TOL
    assert float(np.max(r1)) > 0.05          # not trivially zero


def test_point_scatter_surface_radiance_ray_scatter_tables(reduced_oracle):
    rec, _ = reduced_oracle
    lib = Lib(lib_config(REDUCED))
    ds_a, ds_b, de = rec["R1"], rec["M1"], rec["Ebase"]
    s_prev = rec["R1"]
    for it in range(3):
        assert rel_err(lib.point_scatter(ds_a, ds_b, de), rec["dJ%d" % it]) <= TOL, "dJ iteration %d" % it
        assert rel_err(lib.surface_radiance(ds_a, ds_b), rec["dE%d" % it]) <= TOL, "dE iteration %d" % it
        assert rel_err(lib.ray_scatter(rec["dJ%d" % it]), rec["dS%d" % it]) <= TOL, "dS iteration %d" % it
        assert rel_err(lib.resample(0, s_prev, rec["dS%d" % it]), rec["S%d" % it]) <= TOL, "S iteration %d" % it
        e_prev = rec["E%d" % (it - 1)] if it else None
        assert rel_err(lib.resample(1, e_prev, rec["dE%d" % it]), rec["E%d" % it]) <= TOL, "E iteration %d" % it
        ds_a, ds_b, de, s_prev = rec["dS%d" % it], None, rec["dE%d" % it], rec["S%d" % it]


def test_resampled_transmittance(reduced_oracle):
    rec, _ = reduced_oracle
    lib = Lib(lib_config(REDUCED))
    assert rel_err(lib.resample(2, rec["T"]), rec["LT"]) <= TOL


def test_first_order_sample_count_matches_the_reference():
    """The device counts every overall-extinction sample it evaluates.  The reference evaluates steps^2 view-ray
    samples per TEXEL plus steps sun-ray samples per (texel, outer sample) from which the sun is visible; the kernel
    evaluates the view ray once per (height, elevation) PAIR.  The sun-ray counts must agree exactly -- a bit-exact
    check of every is-above-horizon? decision (atmosphere.clj:95-102,154-160) of the table."""
    cfg = lib_config(REDUCED, iterations=0)
    builder = atmosphere_lut.AtmosphereLutBuilder(cfg=cfg)
    builder.run()
    builder.sync()
    gpu = builder.work()["esamples_first_order"]
    builder.close()
    h, e, s, a = REDUCED["shape4"]
    steps = REDUCED["ray_steps"]
    orc.counters_reset()
    orc.table_first_order(orc.planet(**orc.EARTH), [orc.scatter(**orc.MIE), orc.scatter(**orc.RAYLEIGH)],
                          orc_config(REDUCED), orc.scatter(**orc.RAYLEIGH), 0)
    ref = orc.counters_get()["esamples"]
    sun_ref = ref - h * e * s * a * steps * steps
    sun_gpu = gpu - h * e * steps * steps
    assert sun_ref > 0 and sun_gpu == sun_ref


# ---------------------------------------------------------------- whole build

def test_generate_matches_oracle(reduced_oracle):
    _, files = reduced_oracle
    tables = atmosphere_lut.generate_tables(cfg=lib_config(REDUCED, iterations=3))
    for name, got, want in zip(atmosphere_lut.FILE_NAMES, tables, files):
        assert rel_err(got, want) <= TOL, name


def test_generate_matches_golden_files():
    """BASELINE.json configs[0] + all 5 iterations against the committed oracle outputs (tests/golden)."""
    tables = atmosphere_lut.generate_tables(cfg=lib_config(REDUCED, iterations=5))
    for name, got in zip(("transmittance", "surface-radiance", "ray-scatter", "mie-strength"), tables):
        want = np.fromfile(os.path.join(GOLDEN, "reduced_%s.scatter" % name), dtype="<f4")
        assert rel_err(got, want) <= TOL, name


def test_written_files_are_byte_compatible(tmp_path):
    """spit-floats layout: headerless little-endian float32, sizes of SURVEY.md App. A.8 (scaled)."""
    cfg = lib_config(REDUCED, iterations=1)
    tables = atmosphere_lut.generate_tables(cfg=cfg)
    paths = atmosphere_lut.write_tables(tables, str(tmp_path))
    assert [os.path.basename(p) for p in paths] == ["transmittance.scatter", "surface-radiance.scatter",
                                                    "ray-scatter.scatter", "mie-strength.scatter"]
    h, e, s, a = REDUCED["shape4"]
    sizes = [16 * 63 * 12, 4 * 15 * 12, h * s * e * a * 12, h * s * e * a * 12]
    for p, t, size in zip(paths, tables, sizes):
        assert os.path.getsize(p) == size
        assert open(p, "rb").read() == np.asarray(t, dtype="<f4").tobytes()
        np.testing.assert_array_equal(orc.slurp_floats(p), np.asarray(t).reshape(-1))


def test_file_layout_is_convert_4d_to_2d(reduced_oracle):
    """The 4-D files are tiled exactly like image.clj:299-312: row = h*S + s, column = e*A + a."""
    rec, _ = reduced_oracle
    cfg = lib_config(REDUCED, iterations=0)
    tables = atmosphere_lut.generate_tables(cfg=cfg)
    lib = Lib(cfg)
    m1 = lib.first_order()[1]
    logical = lib.resample(0, m1)                     # make-lookup-table of first-order-mie-strength (:101)
    tiled = orc.convert_4d_to_2d(logical.astype(np.float64))
    np.testing.assert_array_equal(tables[3], tiled.astype(np.float32))
    out = np.zeros_like(tiled, dtype=np.float32)
    shape = (C.c_int * 4)(*REDUCED["shape4"])
    _lib.check(_lib.load().atmlut_convert_4d_to_2d(_lib.ptr(logical), shape, 3, _lib.ptr(out)))
    np.testing.assert_array_equal(out, tables[3])
    # iterations = 0: S stays first-order Rayleigh, E stays 0 (atmosphere_lut.clj:76,85)
    assert float(np.max(np.abs(tables[1]))) == 0.0
    assert rel_err(tables[2], orc.pack_floats(orc.convert_4d_to_2d(
        orc.table_resample_sum_4d(orc.planet(**orc.EARTH), orc_config(REDUCED), [rec["R1"]])))) <= TOL


# ---------------------------------------------------------------- edge cases

@pytest.mark.parametrize("case", [
    dict(shape4=(2, 2, 2, 2), shape_t=(2, 2), shape_e=(2, 2), ray_steps=4, sphere_steps=2),       # minimum sizes
    dict(shape4=(3, 7, 5, 3), shape_t=(4, 9), shape_e=(3, 5), ray_steps=7, sphere_steps=5),       # ragged, odd
    dict(shape4=(2, 5, 20, 16), shape_t=(3, 5), shape_e=(2, 3), ray_steps=16, sphere_steps=6),    # > 256 texels per pair
    dict(shape4=(4, 6, 3, 2), shape_t=(4, 6), shape_e=(3, 4), ray_steps=33, sphere_steps=9),      # even elevation size
    dict(shape4=(2, 3, 64, 16), shape_t=(3, 5), shape_e=(2, 3), ray_steps=8, sphere_steps=4),     # 1024 texels per pair (stress layout)
    dict(shape4=(2, 3, 40, 32), shape_t=(3, 5), shape_e=(2, 3), ray_steps=8, sphere_steps=4),     # 1280 > 1024: two chunks, wide rows
    dict(shape4=(2, 3, 5, 6), shape_t=(3, 5), shape_e=(2, 3), ray_steps=9, sphere_steps=4),       # heading size not a power of two
])
def test_small_and_ragged_shapes(case):
    pl = orc.planet(**orc.EARTH)
    files = orc.generate_atmosphere_luts(pl, orc.scatter(**orc.MIE), orc.scatter(**orc.RAYLEIGH), orc_config(case),
                                         iterations=2)
    tables = atmosphere_lut.generate_tables(cfg=lib_config(case, iterations=2))
    for name, got, want in zip(atmosphere_lut.FILE_NAMES, tables, files):
        assert rel_err(got, want) <= TOL, name


def test_thick_atmosphere_uses_the_general_height_path():
    """A small planet with a deep atmosphere: (Rt^2 - R^2)/R^2 is far beyond the series of the fast sampler."""
    planet = dict(radius=1000.0, height=400.0, brightness=(0.2, 0.3, 0.4))
    mie = dict(base=(2e-3, 2e-3, 2e-3), scale=40.0, g=0.6, quotient=0.8)
    ray = dict(base=(1e-3, 2e-3, 4e-3), scale=120.0)
    case = dict(shape4=(4, 9, 4, 2), shape_t=(5, 9), shape_e=(3, 5), ray_steps=12, sphere_steps=6)
    pl = orc.planet(planet["radius"], planet["height"], planet["brightness"])
    files = orc.generate_atmosphere_luts(pl, orc.scatter(**mie), orc.scatter(**ray), orc_config(case), iterations=2)
    tables = atmosphere_lut.generate_tables(planet, (mie, ray), lib_config(case, iterations=2))
    for name, got, want in zip(atmosphere_lut.FILE_NAMES, tables, files):
        assert rel_err(got, want) <= TOL, name


def test_test_suite_planet_height_100km():
    """The reference's own test atmosphere (t_atmosphere.clj:44-47): 100 km, size 12, ray-steps 10."""
    planet = dict(atmosphere_lut.earth, height=100000.0)
    case = dict(shape4=(12, 12, 12, 12), shape_t=(12, 12), shape_e=(12, 12), ray_steps=10, sphere_steps=15)
    gold = np.load(os.path.join(GOLDEN, "small_shader_luts.npz"))
    lib = Lib(lib_config(case), planet=planet)
    assert rel_err(lib.transmittance(), gold["T"]) <= TOL
    assert rel_err(lib.first_order()[0], gold["S"]) <= TOL


def test_intensity_scales_linearly():
    cfg1 = lib_config(REDUCED, iterations=1)
    cfg2 = lib_config(REDUCED, iterations=1, intensity=(2.0, 3.0, 0.5))
    t1 = atmosphere_lut.generate_tables(cfg=cfg1)
    t2 = atmosphere_lut.generate_tables(cfg=cfg2)
    np.testing.assert_array_equal(t1[0], t2[0])                       # transmittance does not see the sun
    scale = np.array([2.0, 3.0, 0.5])
    for a, b in zip(t1[1:], t2[1:]):
        assert rel_err(b, a.astype(np.float64) * scale) <= 1e-5


def test_invalid_arguments_raise():
    with pytest.raises(_lib.AtmlutError):
        atmosphere_lut.generate_tables(cfg=_lib.make_config(ray_steps=0))
    with pytest.raises(_lib.AtmlutError):
        atmosphere_lut.generate_tables(cfg=_lib.make_config(height_size=1))
    with pytest.raises(_lib.AtmlutError, match="6144"):
        atmosphere_lut.generate_tables(cfg=_lib.make_config(ray_scatter_shape=(2, 2, 100, 100), iterations=0))
    with pytest.raises(_lib.AtmlutError):
        atmosphere_lut.generate_tables(cfg=_lib.make_config(ray_steps=300))
    with pytest.raises(_lib.AtmlutError):
        atmosphere_lut.generate_tables(planet=dict(atmosphere_lut.earth, centre=(1.0, 0.0, 0.0)),
                                       cfg=lib_config(REDUCED, iterations=0))
    with pytest.raises(_lib.AtmlutError):
        atmosphere_lut.generate_tables(scatter=(atmosphere_lut.mie,), cfg=lib_config(REDUCED, iterations=0))


# ---------------------------------------------------------------- shipped resolution (BASELINE.json configs[1..2])

@pytest.fixture(scope="module")
def shipped():
    cfg = _lib.default_config()
    builder = atmosphere_lut.AtmosphereLutBuilder(cfg=cfg)
    builder.run()
    builder.sync()
    tables = builder.download()
    work = builder.work()
    builder.close()
    return cfg, tables, work


def test_shipped_2d_tables_match_oracle(shipped):
    """configs[1]: transmittance [64,255] in full; surface radiance checked through its inputs below."""
    cfg, tables, _ = shipped
    pl = orc.planet(**orc.EARTH)
    ocfg = orc.config(cfg.ray_scatter_shape, cfg.transmittance_shape, cfg.surface_radiance_shape)
    mie, ray = orc.scatter(**orc.MIE), orc.scatter(**orc.RAYLEIGH)
    T = orc.table_transmittance(pl, [mie, ray], ocfg)
    LT = orc.table_resample_sum_t(pl, ocfg, [T])
    assert tables[0].shape == (64, 255, 3)
    assert rel_err(tables[0], LT) <= TOL
    assert tables[0].tobytes().__len__() == 195840
    assert tables[1].tobytes().__len__() == 12096
    assert tables[2].tobytes().__len__() == 12484608 == tables[3].tobytes().__len__()


def test_shipped_first_order_sampled_texels():
    """Random texels of the shipped-resolution first-order tables against the oracle (the oracle needs
    about 1.3 ms per texel and pass, so the full 1 040 384-texel table is sampled)."""
    cfg = _lib.default_config()
    lib = Lib(cfg)
    r1, m1 = lib.first_order()
    rng = np.random.default_rng(7)
    idx = np.sort(rng.choice(r1.size // 3, size=384, replace=False))
    pl = orc.planet(**orc.EARTH)
    ocfg = orc.config(cfg.ray_scatter_shape, cfg.transmittance_shape, cfg.surface_radiance_shape)
    mie, ray = orc.scatter(**orc.MIE), orc.scatter(**orc.RAYLEIGH)
    want_r = orc.table_first_order(pl, [mie, ray], ocfg, ray, 0, idx)
    want_m = orc.table_first_order(pl, [mie, ray], ocfg, mie, 1, idx)
    assert rel_err(r1.reshape(-1, 3)[idx], want_r) <= TOL
    assert rel_err(m1.reshape(-1, 3)[idx], want_m) <= TOL


def test_shipped_iteration_kernels_sampled_texels():
    """One scattering iteration at shipped resolution: every kernel's output is checked on random texels
    against the oracle evaluated on the SAME (GPU-produced) input tables."""
    cfg = _lib.default_config()
    lib = Lib(cfg)
    r1, m1 = lib.first_order()
    e0 = lib.surface_radiance_base()
    dj = lib.point_scatter(r1, m1, e0)
    de = lib.surface_radiance(r1, m1)
    ds = lib.ray_scatter(dj)
    s1 = lib.resample(0, r1, ds)
    dj2 = lib.point_scatter(ds, None, de)
    rng = np.random.default_rng(11)
    n4 = r1.size // 3
    idx = np.sort(rng.choice(n4, size=256, replace=False))
    idx_e = np.sort(rng.choice(e0.size // 3, size=32, replace=False))
    pl = orc.planet(**orc.EARTH)
    ocfg = orc.config(cfg.ray_scatter_shape, cfg.transmittance_shape, cfg.surface_radiance_shape)
    mie, ray = orc.scatter(**orc.MIE), orc.scatter(**orc.RAYLEIGH)
    src1 = orc.SSourceSpec(r1, m1, mie)
    assert rel_err(dj.reshape(-1, 3)[idx], orc.table_point_scatter(pl, [mie, ray], ocfg, src1, e0, idx)) <= TOL
    assert rel_err(de.reshape(-1, 3)[idx_e], orc.table_surface_radiance(pl, ocfg, src1, idx_e)) <= TOL
    assert rel_err(ds.reshape(-1, 3)[idx], orc.table_ray_scatter(pl, [mie, ray], ocfg, dj, idx)) <= TOL
    assert rel_err(s1.reshape(-1, 3)[idx], orc.table_resample_sum_4d(pl, ocfg, [r1, ds], idx)) <= TOL
    src2 = orc.SSourceSpec(ds)
    assert rel_err(dj2.reshape(-1, 3)[idx], orc.table_point_scatter(pl, [mie, ray], ocfg, src2, de, idx)) <= TOL


def test_shipped_build_properties(shipped):
    """Size-independent properties of the full shipped build."""
    cfg, (t, e, s, m), work = shipped
    for tab in (t, e, s, m):
        assert np.isfinite(tab).all() and float(tab.min()) >= 0.0
    assert float(t.max()) <= 1.0 + 1e-6 and float(t.min()) > 0.0
    # energy: every order adds light, and orders decay -> total S is between first order and a small multiple
    assert 0.2 < float(s.max()) < 1.0
    H, E, S, A = cfg.ray_scatter_shape
    # on the ground every below-horizon view ray has length zero (surface-intersection distance 0), and the
    # re-tabulation maps that half row onto its own middle texel: exactly no light
    m4 = m.reshape(H, S, E, A, 3)
    assert float(np.abs(m4[0, :, :E // 2 + 1]).max()) == 0.0
    assert float(np.abs(s.reshape(H, S, E, A, 3)[0, :, :E // 2 + 1]).max()) == 0.0
    assert float(m4[0, S - 1, E // 2 + 1:].min()) > 0.0
    # blue scatters more than red looking up from the ground with the sun at the zenith
    s4 = s.reshape(H, S, E, A, 3)
    assert np.all(s4[0, S - 1, E // 2 + 1, :, 2] > s4[0, S - 1, E // 2 + 1, :, 0])
    # the executed sample count is below the reference's bound N4 * steps * 2 steps
    assert 0 < work["esamples_first_order"] <= 2.0 * H * E * S * A * 100 * 100
    assert work["kernel_launches"] > 20


# ---------------------------------------------------------------- sampler variants (atm_device.cuh density_sums_seq)

@pytest.mark.parametrize("name,first,second,height", [
    # Earth: scale heights 1200 m : 8000 m = 3 : 20, thin atmosphere -> alternating one- / two-exponential pairs
    ("earth", dict(atmosphere_lut.mie), dict(atmosphere_lut.rayleigh), 35000.0),
    # the same media listed the other way round: the powers t^20 / t^3 go to the other component
    ("swapped", dict(atmosphere_lut.rayleigh), dict(atmosphere_lut.mie), 35000.0),
    # scale heights in no small ratio: two exponentials on every pair, series cut after u^2
    ("incommensurable", dict(atmosphere_lut.mie), dict(atmosphere_lut.rayleigh, scale=7994.0), 35000.0),
    # 60 km of atmosphere: the u^3 term matters, full series and two exponentials
    ("thicker", dict(atmosphere_lut.mie), dict(atmosphere_lut.rayleigh), 60000.0),
])
def test_every_sampler_variant_meets_the_tolerance(name, first, second, height):
    """The library picks the overall-extinction sampler from the planet and the media (series degree, one or two
    exponentials per sample); whichever it picks, the build matches the oracle within the tolerance."""
    c = dict(shape4=(4, 15, 4, 4), shape_t=(8, 31), shape_e=(4, 7), ray_steps=100, sphere_steps=8)
    planet = dict(atmosphere_lut.earth, height=height)
    got = atmosphere_lut.generate_tables(planet, (first, second), lib_config(c, iterations=1))
    want = orc.generate_atmosphere_luts(orc.planet(planet["radius"], height, planet["brightness"]),
                                        orc.scatter(**{k: first[k] for k in first}),
                                        orc.scatter(**{k: second[k] for k in second}), orc_config(c), iterations=1)
    for file_name, g, w in zip(atmosphere_lut.FILE_NAMES, got, want):
        if file_name == "mie-strength.scatter":
            # A single-order table re-tabulated through forward . backward: where a texel and its own row are exactly 0
            # the entry is a neighbour times a one-ulp interpolation weight -- rounding noise of the reference's own
            # arithmetic (2.6e-20 against an exact 0 here with the 8000 m medium listed first).
            # In the other files that noise drowns in the sum over the orders.
            g64, w64 = np.asarray(g, np.float64).reshape(-1), np.asarray(w, np.float64).reshape(-1)
            floor = max(FLOOR, 1e-12 * float(np.abs(w64).max()))
            assert float(np.max(np.abs(g64 - w64) / np.maximum(np.abs(w64), floor))) <= TOL, (name, file_name)
        else:
            assert rel_err(g, w) <= TOL, (name, file_name)


def test_the_builder_reports_the_sampler_it_picked():
    """atmlut_builder_counter(3): MUFU.EX2 per overall-extinction sample -- 1.5 for Earth (every second pair of samples
    from one exponential), 2 where the scale heights are in no 3 : 20 ratio or the atmosphere is not thin"""
    cfg = lib_config(REDUCED, iterations=0)
    for scatter, height, want in (((atmosphere_lut.mie, atmosphere_lut.rayleigh), 35000.0, 1.5),
                                  ((atmosphere_lut.rayleigh, atmosphere_lut.mie), 35000.0, 1.5),
                                  ((atmosphere_lut.mie, dict(atmosphere_lut.rayleigh, scale=7994.0)), 35000.0, 2.0),
                                  ((atmosphere_lut.mie, atmosphere_lut.rayleigh), 100000.0, 2.0)):
        b = atmosphere_lut.AtmosphereLutBuilder(planet=dict(atmosphere_lut.earth, height=height), scatter=scatter, cfg=cfg)
        b.run()
        b.sync()
        assert b.work()["mufu_ex2_per_esample"] == want
        b.close()


def test_generate_into_page_locked_and_pageable_buffers_gives_the_same_bytes():
    """atmlut_generate makes the four downloads nodes of the build graph when the destinations are page-locked and
    re-captures the graph when they change; pageable destinations go through the library's staging buffers.  Every
    variant, in any order, must deliver the same bytes."""
    cfg = lib_config(REDUCED, iterations=2)
    want = [t.copy() for t in atmosphere_lut.generate_tables(cfg=cfg)]                    # pageable
    pinned_a = atmosphere_lut.allocate_outputs(cfg, pinned=True)
    pinned_b = atmosphere_lut.allocate_outputs(cfg, pinned=True)
    lib = _lib.load()
    nbytes = [t.nbytes for t in want]
    owned = [lib.atmlut_host_alloc(n) for n in nbytes]                                    # the library's own allocator
    assert all(owned)
    owned_views = [np.frombuffer((C.c_char * n).from_address(p), dtype=np.float32).reshape(t.shape)
                   for p, n, t in zip(owned, nbytes, want)]
    try:
        for out in (pinned_a, pinned_b, pinned_a, None, owned_views, pinned_b, None):
            if out is not None:
                for t in out:
                    t.fill(-1.0)
            got = atmosphere_lut.generate_tables(cfg=cfg, out=out)
            for g, w in zip(got, want):
                assert g.tobytes() == w.tobytes()
    finally:
        for p in owned:
            lib.atmlut_host_free(C.c_void_p(p))
