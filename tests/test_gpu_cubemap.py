"""GPU parity of the cube-map tile path (include/sfsim_cubemap.h) against oracle/cubemap_oracle.c, which
tests/test_cubemap_oracle.py pins to t_cubemap.clj.  Integer outputs (colour bytes, water bytes, encoded normals) must be
identical (see assert_bytes_equal for the one class of pixels no implementation can pin); float outputs (surface offsets, normals) may differ by the last float32 bit where the two libms differ in the
last double bit of atan2 / sin / cos (tolerances below)."""
import numpy as np
import pytest

from oracle import cubemap as ocm
from sfsim_b200 import _lib, cubemap

pytestmark = pytest.mark.gpu
R = 6378000.0


def make_worlds(width, levels_elevation, levels_color, seed):
    elev, day, night = ocm.synthetic_world(width, levels_elevation, levels_color, seed=seed)
    ow = ocm.OracleWorld(width, elev, day, night)
    gw = cubemap.World(width)
    for level, a in elev.items():
        gw.set_elevation(level, a)
    for level, a in day.items():
        gw.set_color(False, level, a)
    for level, a in night.items():
        gw.set_color(True, level, a)
    return ow, gw, (elev, day, night)


def assert_bytes_equal(got, want, raw, where):
    """identical, except where the reference's own un-truncated double sits on an integer (regions where the raster is
    constant: v (w0 + w1 + w2 + w3) with weights summing to 1 or to 1 - 1e-16): there the byte is decided by the last
    bit of atan2 in whichever libm computed lon / lat, and the neighbouring integer is the other legitimate answer"""
    bad = got != want
    if bad.any():
        on_integer = np.abs(raw - np.rint(raw)) < 1e-9
        one_off = np.abs(got.astype(np.int32) - want.astype(np.int32)) == 1
        where_bad = np.argwhere(bad & ~(on_integer & one_off))
        assert len(where_bad) == 0, "%s: %d bytes differ, first at %s: %s vs %s (raw %r)" % (
            where, len(where_bad), where_bad[0], got[tuple(where_bad[0])], want[tuple(where_bad[0])], raw[tuple(where_bad[0])])


def assert_tile_equal(got, want, where):
    ct = want["day"].shape[0]
    assert_bytes_equal(got["day"][..., :3], want["day"][..., :3], want["raw"][..., 0:3], (where, "day"))
    assert_bytes_equal(got["night"][..., :3], want["night"][..., :3], want["raw"][..., 3:6], (where, "night"))
    assert_bytes_equal(got["water"][:, :ct], want["water"][:, :ct], want["raw"][..., 6], (where, "water"))
    assert (got["day"][..., 3] == 255).all() and (got["night"][..., 3] == 255).all() and (got["water"][:, ct:] == 0).all()
    # surface offsets: float32 of (point - centre); the doubles agree to ~1e-9 m, so at most the last float bit moves
    ref = want["surface"].astype(np.float64)
    err = np.abs(got["surface"].astype(np.float64) - ref)
    assert (err <= np.spacing(np.abs(want["surface"]).astype(np.float32)).astype(np.float64) + 1e-6).all(), (where, err.max())
    assert (got["surface"] != want["surface"]).mean() < 0.01, where
    err = np.abs(got["normals"].astype(np.float64) - want["normals"].astype(np.float64))
    assert err.max() <= 1.2e-7, (where, err.max())
    assert (got["normals"] != want["normals"]).mean() < 0.01, where
    # the encoded normals: identical unless the value sits on a rounding boundary of the encoder (a component that is
    # exactly 0 over flat terrain in one libm and -1e-17 in the other; a float that moved by its last bit): then by one
    scaled = want["normals"].astype(np.float64) * 127.5
    want_bytes = np.floor((scaled - 0.5) + 0.5).astype(np.int8)
    d = np.abs(got["normal_bytes"].astype(np.int32) - want_bytes.astype(np.int32))
    on_boundary = np.abs(scaled - np.rint(scaled)) < 2e-5
    assert (d[~on_boundary] == 0).all() and d.max() <= 1, (where, d.max(), np.argwhere((d != 0) & ~on_boundary)[:3])
    same = got["normals"] == want["normals"]
    assert (got["normal_bytes"][same] == want_bytes[same]).all(), where


@pytest.mark.parametrize("in_level,out_level,tiles", [
    (-3, 0, [(0, 0, 0), (1, 0, 0), (2, 0, 0), (3, 0, 0), (4, 0, 0), (5, 0, 0)]),     # build.clj:300-302, every face
    (-1, 2, [(0, 1, 2), (2, 3, 0), (5, 0, 3)]),                                        # in-level below 0 -> level 0 / 0
    (0, 3, [(1, 7, 7), (3, 2, 5), (4, 0, 0)]),                                         # levels 0 / 1 / 1
    (1, 4, [(0, 8, 8), (2, 15, 0), (5, 3, 12)]),                                       # levels 1 / 2 / 2
    (2, 5, [(4, 31, 16)]),                                                             # clamped to the maximum levels
])
def test_tiles_match_the_oracle(in_level, out_level, tiles):
    width, st = 24, 17
    ow, gw, _ = make_worlds(width, [0, 1, 2], [0, 1, 2], seed=100 + out_level)
    cfg = cubemap.make_config(in_level, out_level, width=width, surface_tilesize=st, max_surface_level=2, max_color_level=2)
    got = gw.make_cube_map_tiles(cfg, tiles)
    for t, (face, b, a) in enumerate(tiles):
        want = ow.make_cube_map_tile(face, in_level, out_level, b, a, surface_tilesize=st, max_surface_level=2,
                                     max_color_level=2)
        assert_tile_equal({k: v[t] for k, v in got.items()}, want, (face, b, a))
    gw.close()


def test_a_tile_with_the_shipped_constants():
    """width 675, 65 / 129 pixel tiles, levels as make-cube-map picks them for (cube-map {:in-level 0 :out-level 3})"""
    ow, gw, _ = make_worlds(675, [0, 1], [1], seed=7)
    cfg = cubemap.make_config(0, 3)
    tiles = [(2, 3, 4), (5, 7, 0)]
    got = gw.make_cube_map_tiles(cfg, tiles)
    assert got["day"].shape == (2, 129, 129, 4) and got["water"].shape == (2, 129, 132) and got["surface"].shape == (2, 65, 65, 3)
    for t, (face, b, a) in enumerate(tiles):
        assert_tile_equal({k: v[t] for k, v in got.items()}, ow.make_cube_map_tile(face, 0, 3, b, a), (face, b, a))
    gw.close()


def test_pointwise_functions_match_the_oracle():
    width = 32
    ow, gw, _ = make_worlds(width, [0, 2], [1], seed=21)
    rng = np.random.default_rng(5)
    p = rng.normal(size=(300, 3)) * rng.uniform(0.5, 2.0, size=(300, 1)) * R
    p[:6] = np.array([(1, 0, 0), (-1, 0, 0), (0, 1, 0), (0, -1, 0), (0, 0, 1), (0, 0, -1)]) * R   # poles and the date line
    for level in (0, 2):
        got = gw.project_onto_globe(p, level)
        want = np.array([ow.project_onto_globe(q, level, R) for q in p])
        assert np.abs(got - want).max() <= 1e-8, level                        # metres, on a 6.4e6 m sphere
    got = gw.normal_for_point(p, 2, 5, 129)
    want = np.array([ow.normal_for_point(q, 2, 5, 129, R) for q in p])
    assert np.abs(got - want).max() <= 1e-10
    lon = rng.uniform(-np.pi, np.pi, 500)
    lat = rng.uniform(-np.pi / 2, np.pi / 2, 500)
    lon[:4], lat[:4] = (-np.pi, np.pi, 0.0, np.pi - 1e-9), (np.pi / 2, -np.pi / 2, 0.0, 0.3)
    for level in (0, 2):
        assert gw.elevation_geodetic(level, lon, lat).tolist() == [ow.elevation_geodetic(level, a, b) for a, b in zip(lon, lat)]
        assert gw.water_geodetic(level, lon, lat).tolist() == [ow.water_geodetic(level, a, b) for a, b in zip(lon, lat)]
    assert gw.color_geodetic_day(1, lon, lat).tolist() == [ow.color_geodetic(0, 1, a, b).tolist() for a, b in zip(lon, lat)]
    assert gw.color_geodetic_night(1, lon, lat).tolist() == [ow.color_geodetic(1, 1, a, b).tolist() for a, b in zip(lon, lat)]
    gw.close()


def test_reference_facts_on_the_gpu():
    # t_cubemap.clj:361-383 with constant rasters in place of the mocked elevation-geodetic; :322-358 water
    width = 4
    gw = cubemap.World(width)
    gw.set_elevation(0, np.full(cubemap.level_shape(width, 0), 2777, np.int16))
    assert np.linalg.norm(gw.project_onto_globe([(0, 0, -1)], 0)[0] - np.array([0, 0, -6380777.0])) < 1e-6
    gw.set_elevation(0, np.full(cubemap.level_shape(width, 0), -500, np.int16))
    assert np.linalg.norm(gw.project_onto_globe([(1, 0, 0)], 0)[0] - np.array([6378000.0, 0, 0])) < 1e-6
    assert gw.water_geodetic(0, [0.0], [0.0]).tolist() == [255]
    gw.set_elevation(0, np.full(cubemap.level_shape(width, 0), 100, np.int16))
    assert gw.water_geodetic(0, [0.0], [0.0]).tolist() == [0]
    # flat terrain: the normal is the direction of the point (t_cubemap.clj:407-411)
    n = gw.normal_for_point([(R, 0, 0), (0, 0, -R), (0.6 * R, 0, 0.8 * R)], 0, 5, 33)
    np.testing.assert_allclose(n, [(1, 0, 0), (0, 0, -1), (0.6, 0, 0.8)], atol=1e-9)
    gw.close()


def test_tile_uploads_equal_a_level_upload_and_missing_rasters_are_errors():
    width = 16
    elev, day, night = ocm.synthetic_world(width, [0, 1], [1], seed=9)
    whole = cubemap.World(width)
    piecewise = cubemap.World(width)
    cfg = cubemap.make_config(0, 1, width=width, surface_tilesize=9)
    with pytest.raises(_lib.AtmlutError, match="elevation raster of level 0"):
        piecewise.make_cube_map_tiles(cfg, [(0, 0, 0)])
    for level in (0, 1):
        whole.set_elevation(level, elev[level])
        for ty in range(elev[level].shape[0]):
            for tx in range(elev[level].shape[1]):
                piecewise.set_elevation_tile(level, ty, tx, elev[level][ty, tx])
    whole.set_color(False, 1, day[1])
    whole.set_color(True, 1, night[1])
    with pytest.raises(_lib.AtmlutError, match="day raster of level 1"):
        piecewise.make_cube_map_tiles(cfg, [(0, 0, 0)])
    for ty in range(day[1].shape[0]):
        for tx in range(day[1].shape[1]):
            piecewise.set_color_tile(False, 1, ty, tx, day[1][ty, tx])
            piecewise.set_color_tile(True, 1, ty, tx, night[1][ty, tx])
    tiles = cubemap.tile_shard(1)
    a, b = whole.make_cube_map_tiles(cfg, tiles), piecewise.make_cube_map_tiles(cfg, tiles)
    for k in a:
        assert a[k].tobytes() == b[k].tobytes(), k
    with pytest.raises(_lib.AtmlutError, match="out of range"):
        whole.make_cube_map_tiles(cfg, [(0, 2, 0)])
    with pytest.raises(_lib.AtmlutError, match="out of range"):
        whole.set_elevation_tile(0, 2, 0, elev[0][0, 0])
    with pytest.raises(TypeError):
        whole.set_elevation(1, elev[0])
    assert whole.make_cube_map_tiles(cfg, np.zeros((0, 3), np.int32))["day"].shape == (0, 17, 17, 4)
    # only the outputs asked for are produced
    assert sorted(whole.make_cube_map_tiles(cfg, tiles[:2], outputs=("water", "surface"))) == ["surface", "water"]
    whole.close()
    piecewise.close()


def test_a_whole_level_has_the_properties_of_the_pyramid():
    """Size-independent checks on every tile of an output level at the shipped tile sizes: shards reproduce the single
    batch byte for byte, neighbouring tiles share their border pixels, normals are unit vectors, alpha is 255, the pad
    columns of the water image are zero."""
    width = 64
    elev, day, night = ocm.synthetic_world(width, [0, 1], [1], seed=13)
    gw = cubemap.World(width)
    for level in (0, 1):
        gw.set_elevation(level, elev[level])
    gw.set_color(False, 1, day[1])
    gw.set_color(True, 1, night[1])
    cfg = cubemap.make_config(0, 2, width=width)
    tiles = cubemap.tile_shard(2)
    out = gw.make_cube_map_tiles(cfg, tiles)
    parts = [gw.make_cube_map_tiles(cfg, cubemap.tile_shard(2, r, 4)) for r in range(4)]
    for k, v in out.items():
        for r in range(4):
            assert v[r::4].tobytes() == parts[r][k].tobytes(), (k, r)
    index = {tuple(t): n for n, t in enumerate(tiles.tolist())}
    for face in range(6):
        for b in range(4):
            for a in range(3):
                left, right = index[(face, b, a)], index[(face, b, a + 1)]
                for k in ("day", "night", "normal_bytes"):
                    assert (out[k][left][:, -1] == out[k][right][:, 0]).all(), (k, face, b, a)
                assert (out["water"][left][:, 128] == out["water"][right][:, 0]).all()
                up, down = index[(face, a, b)], index[(face, a + 1, b)]
                assert (out["day"][up][-1] == out["day"][down][0]).all()
    assert np.abs(np.linalg.norm(out["normals"].astype(np.float64), axis=-1) - 1).max() < 1e-6
    assert (out["day"][..., 3] == 255).all() and (out["night"][..., 3] == 255).all()
    assert (out["water"][:, :, 129:] == 0).all()
    assert np.isfinite(out["surface"]).all() and np.abs(out["surface"]).max() < R
    gw.close()


def test_the_streamed_level_delivers_every_tile_once_and_equals_the_batch_call():
    width = 32
    elev, day, night = ocm.synthetic_world(width, [0, 1], [1], seed=17)
    gw = cubemap.World(width)
    for level in (0, 1):
        gw.set_elevation(level, elev[level])
    gw.set_color(False, 1, day[1])
    gw.set_color(True, 1, night[1])
    cfg = cubemap.make_config(0, 2, width=width, surface_tilesize=17)
    for rank, world_size, batch in ((0, 1, 7), (1, 3, 5), (0, 1, 200)):      # ragged last batch, one batch, a shard
        tiles = cubemap.tile_shard(2, rank, world_size)
        want = gw.make_cube_map_tiles(cfg, tiles)
        seen = []

        def on_tile(key, tile):
            n = len(seen)
            assert key == tuple(tiles[n])
            for k, v in tile.items():
                assert v.tobytes() == want[k][n].tobytes(), (key, k)
            seen.append(key)

        assert gw.make_cube_map(0, 2, on_tile, rank=rank, world_size=world_size, batch=batch, surface_tilesize=17) == len(tiles)
        assert len(seen) == len(tiles)

    # a subset of the arrays: the others are neither brought back nor handed over; encoded normals without float normals
    tiles = cubemap.tile_shard(2)
    want = gw.make_cube_map_tiles(cfg, tiles)
    got = []

    def subset(key, tile):
        assert sorted(tile) == ["normal_bytes", "water"]
        n = len(got)
        assert tile["water"].tobytes() == want["water"][n].tobytes()
        assert tile["normal_bytes"].tobytes() == want["normal_bytes"][n].tobytes()
        got.append(key)

    assert gw.make_cube_map(0, 2, subset, batch=50, outputs=("water", "normal_bytes"), surface_tilesize=17) == len(tiles)
    with pytest.raises(TypeError):
        gw.make_cube_map(0, 2, subset, outputs=("water", "colour"), surface_tilesize=17)

    class Stop(Exception):
        pass

    def failing(key, tile):
        raise Stop()

    with pytest.raises(Stop):
        gw.make_cube_map(0, 2, failing, batch=4, surface_tilesize=17)
    assert gw.make_cube_map(0, 0, lambda key, tile: None, rank=7, world_size=8, surface_tilesize=17) == 0
    gw.close()


def test_committed_golden_tiles():
    """the GPU against tests/golden/cubemap_tiles.npz (made by the oracle, tests/golden/make_cubemap_golden.py)"""
    import importlib.util
    import os
    here = os.path.join(os.path.dirname(__file__), "golden")
    spec = importlib.util.spec_from_file_location("make_cubemap_golden", os.path.join(here, "make_cubemap_golden.py"))
    gold_mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(gold_mod)
    gold = np.load(os.path.join(here, "cubemap_tiles.npz"))
    elev, day, night = gold_mod.world()
    gw = cubemap.World(gold_mod.WIDTH)
    for level, a in elev.items():
        gw.set_elevation(level, a)
    for level, a in day.items():
        gw.set_color(False, level, a)
        gw.set_color(True, level, night[level])
    for n, t in enumerate(gold_mod.TILES):
        cfg = cubemap.make_config(t["in_level"], t["out_level"], width=gold_mod.WIDTH, surface_tilesize=gold_mod.SURFACE_TILESIZE)
        got = gw.make_cube_map_tiles(cfg, [(t["face"], t["b"], t["a"])])
        want = {k: gold["%d_%s" % (n, k)] for k in ("day", "night", "water", "surface", "normals", "raw")}
        assert_tile_equal({k: v[0] for k, v in got.items()}, want, ("golden", n))
    gw.close()


def test_make_cube_map_writes_the_files_of_a_level(tmp_path):
    """sfsim_b200.globe.make_cube_map (globe.clj:29-96): every tile of a level as the five files of the reference's layout,
    read back against the arrays of the batch call; then the tar step"""
    import os
    import tarfile
    from PIL import Image
    from sfsim_b200 import globe
    width = 16
    elev, day, night = ocm.synthetic_world(width, [0, 1], [1], seed=19)
    gw = cubemap.World(width)
    for level in (0, 1):
        gw.set_elevation(level, elev[level])
    gw.set_color(False, 1, day[1])
    gw.set_color(True, 1, night[1])
    prefix = str(tmp_path / "globe")
    assert globe.make_cube_map(gw, 0, 1, prefix=prefix, batch=7, surface_tilesize=9) == 24
    tiles = cubemap.tile_shard(1)
    want = gw.make_cube_map_tiles(cubemap.make_config(0, 1, width=width, surface_tilesize=9), tiles)
    for n, (face, b, a) in enumerate(tiles.tolist()):
        path = lambda suffix: globe.cube_path(prefix, face, 1, b, a, suffix)
        assert globe.slurp_bytes_gz(path(".water.gz")).tobytes() == want["water"][n].tobytes()
        assert globe.slurp_floats_gz(path(".surf.gz")).tobytes() == want["surface"][n].tobytes()
        assert np.asarray(Image.open(path(".png"))).view(np.int8).tobytes() == want["normal_bytes"][n].tobytes()
        assert np.abs(globe.slurp_normals(path(".png")) - want["normals"][n]).max() <= 0.5 / 127.5 + 1e-6
        jpg = np.asarray(Image.open(path(".jpg")).convert("RGB")).astype(np.int32)
        assert jpg.shape == (17, 17, 3)                       # lossy: only the size and a loose resemblance
        assert Image.open(path(".night.jpg")).size == (17, 17)
    globe.make_cube_map_tars(1, prefix)
    for face in range(6):
        for a in range(2):
            with tarfile.open(globe.cube_tar(prefix, face, 1, a)) as tar:
                assert len(tar.getnames()) == 2 * 5           # two rows b, five files each
            assert not os.path.exists(globe.cube_dir(prefix, face, 1, a))
    gw.close()
