"""Pins oracle/cubemap_oracle.c to the reference's own facts (test/clj/sfsim/t_cubemap.clj).  Facts that the reference
states with `with-redefs` / `provided` mocks are restated with rasters that make the mocked function return the mocked
value (a constant elevation raster for `elevation-geodetic => 2777.0`, ...) or through the oracle's helper that holds the
mocked step's output (`normal_from_points`, `surrounding_offsets`, `interpolate4`)."""
import numpy as np
import pytest

from oracle import cubemap as cm

PI = cm.PI
R = 6378000.0


# t_cubemap.clj:34-92 "First face of cube" .. "Sixth face of cube"
@pytest.mark.parametrize("face,fn,j,i,want", [
    (0, "z", 0.0, 0.0, 1.0), (0, "x", 0.0, 0.0, -1.0), (0, "x", 0.0, 1.0, 1.0), (0, "y", 0.0, 0.0, 1.0), (0, "y", 1.0, 0.0, -1.0),
    (1, "y", 0.0, 0.0, -1.0), (1, "x", 0.0, 0.0, -1.0), (1, "x", 0.0, 1.0, 1.0), (1, "z", 0.0, 0.0, 1.0), (1, "z", 1.0, 0.0, -1.0),
    (2, "x", 0.0, 0.0, 1.0), (2, "z", 0.0, 0.0, 1.0), (2, "z", 1.0, 0.0, -1.0), (2, "y", 0.0, 0.0, -1.0), (2, "y", 0.0, 1.0, 1.0),
    (3, "y", 0.0, 0.0, 1.0), (3, "x", 0.0, 0.0, 1.0), (3, "x", 0.0, 1.0, -1.0), (3, "z", 0.0, 0.0, 1.0), (3, "z", 1.0, 0.0, -1.0),
    (4, "x", 0.0, 0.0, -1.0), (4, "z", 0.0, 0.0, 1.0), (4, "z", 1.0, 0.0, -1.0), (4, "y", 0.0, 0.0, 1.0), (4, "y", 0.0, 1.0, -1.0),
    (5, "z", 0.0, 0.0, -1.0), (5, "x", 0.0, 0.0, -1.0), (5, "x", 0.0, 1.0, 1.0), (5, "y", 0.0, 0.0, -1.0), (5, "y", 1.0, 0.0, 1.0),
])
def test_cube_faces(face, fn, j, i, want):
    assert getattr(cm, "cube_map_" + fn)(face, j, i) == want


def test_vector_to_cube_face():
    # t_cubemap.clj:95-96
    assert cm.cube_map(5, 0.0, 0.5).tolist() == [0.0, -1.0, -1.0]


@pytest.mark.parametrize("p", [(1.0, 0.2, 0.4), (-1.0, 0.2, 0.4), (0.2, 1.0, 0.4), (0.2, -1.0, 0.4), (0.2, 0.4, 1.0),
                               (0.2, 0.4, -1.0)])
def test_face_round_trip(p):
    # t_cubemap.clj:99-112
    face = cm.determine_face(p)
    np.testing.assert_allclose(cm.cube_map(face, cm.cube_j(face, p), cm.cube_i(face, p)), p, atol=1e-6)


def test_cube_coordinates():
    # t_cubemap.clj:115-118
    assert cm.cube_coordinate(0, 256, 0, 0.0) == 0.0
    assert cm.cube_coordinate(0, 256, 0, 127.5) == 0.5
    assert cm.cube_coordinate(1, 256, 1, 127.5) == 0.75


@pytest.mark.parametrize("face,level,row,column,idx,want", [
    (0, 0, 0, 0, 0, (-1, 1, 1)), (0, 0, 0, 0, 1, (1, 1, 1)), (0, 0, 0, 0, 2, (-1, -1, 1)), (0, 0, 0, 0, 3, (1, -1, 1)),
    (5, 2, 3, 1, 0, (-0.5, 0.5, -1.0)), (5, 2, 3, 1, 1, (0.0, 0.5, -1.0)), (5, 2, 3, 1, 2, (-0.5, 1.0, -1.0)),
    (5, 2, 3, 1, 3, (0.0, 1.0, -1.0))])
def test_cube_map_corners(face, level, row, column, idx, want):
    # t_cubemap.clj:121-132
    assert cm.cube_map_corners(face, level, row, column)[idx].tolist() == list(want)


def test_longitude_latitude():
    # t_cubemap.clj:135-144
    assert abs(cm.longitude((1, 0, 0))) < 1e-6
    assert abs(cm.longitude((0, 1, 0)) - PI / 2) < 1e-6
    assert abs(cm.latitude((0, 6378000, 0))) < 1e-6
    assert abs(cm.latitude((0, 0, 6357000)) - PI / 2) < 1e-6
    assert abs(cm.latitude((6378000, 0, 0))) < 1e-6


@pytest.mark.parametrize("lon,lat,h,want", [
    (0.0, 0.0, 0.0, (6378000.0, 0.0, 0.0)), (PI / 2, 0.0, 0.0, (0.0, 6378000.0, 0.0)), (0.0, PI / 2, 0.0, (0.0, 0.0, 6378000.0)),
    (0.0, 0.0, 1000.0, (6379000.0, 0.0, 0.0)), (PI / 2, 0.0, 1000.0, (0.0, 6379000.0, 0.0))])
def test_geodetic_to_cartesian(lon, lat, h, want):
    # t_cubemap.clj:147-156; roughly-vector 1e-6 is a norm bound; 4e-10 is cos(pi/2) * radius
    assert np.linalg.norm(cm.geodetic_to_cartesian(lon, lat, h, R) - np.array(want)) < 1e-6


@pytest.mark.parametrize("p,want", [
    ((6378000.0, 0.0, 0.0), (0, 0, 0)), ((0.0, 6378000.0, 0.0), (PI / 2, 0, 0)), ((0.0, 0.0, 6378000.0), (0, PI / 2, 0)),
    ((0.0, 0.0, 6379000.0), (0, PI / 2, 1000)), ((0.0, 0.0, -6379000.0), (0, -PI / 2, 1000)),
    ((6379000.0, 0.0, 0.0), (0, 0, 1000)), ((6377900.0, 0.0, 0.0), (0, 0, -100))])
def test_cartesian_to_geodetic(p, want):
    # t_cubemap.clj:159-173
    np.testing.assert_allclose(cm.cartesian_to_geodetic(p, R), want, atol=1e-6)


def test_project_onto_sphere_and_cube():
    # t_cubemap.clj:176-195
    for p, want in (((1, 0, 0), (R, 0, 0)), ((0, 1, 0), (0, R, 0)), ((0, 0, 1), (0, 0, R))):
        np.testing.assert_allclose(cm.project_onto_sphere(p, R), want, atol=1e-6)
    for p, want in (((1, 0, 0), (1, 0, 0)), ((2, 1, 1), (1, 0.5, 0.5)), ((-2, 1, 1), (-1, 0.5, 0.5)), ((0, 1, 0), (0, 1, 0)),
                    ((1, 2, -1), (0.5, 1, -0.5)), ((1, -2, -1), (0.5, -1, -0.5)), ((0, 0, 1), (0, 0, 1)),
                    ((-1, 1, 2), (-0.5, 0.5, 1)), ((1, -1, -2), (0.5, -0.5, -1))):
        assert cm.project_onto_cube(p).tolist() == list(want)


def test_raster_coordinates():
    # t_cubemap.clj:198-219 (exact equalities in the reference)
    assert cm.map_x(-PI, 675, 3) == 0.0
    assert cm.map_x(0.0, 675, 3) == 675 * 2.0 * 8
    assert cm.map_y(PI / 2, 675, 3) == 0.0
    assert cm.map_y(0.0, 675, 3) == 675 * 8.0
    assert cm.map_pixels_x(0.0, 675, 3) == [675 * 2 * 8, 675 * 2 * 8 + 1, 1.0, 0.0]
    assert cm.map_pixels_x(-PI, 675, 3) == [0, 1, 1.0, 0.0]
    assert cm.map_pixels_x(PI - PI / (256 * 4), 256, 0) == [256 * 4 - 1, 0, 0.5, 0.5]
    assert cm.map_pixels_y(0.0, 675, 3) == [675 * 8, 675 * 8 + 1, 1.0, 0.0]
    assert cm.map_pixels_y(-PI / 2, 675, 3) == [675 * 2 * 8 - 1, 675 * 2 * 8 - 1, 1.0, 0.0]
    assert cm.map_pixels_y(PI / (4 * 256), 256, 0) == [256 - 1, 256, 0.5, 0.5]


def test_offsets():
    # t_cubemap.clj:222-236 (these pin fastmath's rotation-matrix-3d-z / -y and mulv)
    d = 2 * PI / (4 * 675)
    np.testing.assert_allclose(cm.offset_longitude((1, 0, 0), 0, 675), (0, d, 0), atol=1e-6)
    np.testing.assert_allclose(cm.offset_longitude((0, -1, 0), 0, 675), (d, 0, 0), atol=1e-6)
    np.testing.assert_allclose(cm.offset_longitude((0, -2, 0), 0, 675), (2 * d, 0, 0), atol=1e-6)
    np.testing.assert_allclose(cm.offset_longitude((0, -2, 0), 1, 675), (d, 0, 0), atol=1e-6)
    np.testing.assert_allclose(cm.offset_latitude((1, 0, 0), 0, 675), (0, 0, d), atol=1e-6)
    np.testing.assert_allclose(cm.offset_latitude((0, 0, 1), 0, 675), (-d, 0, 0), atol=1e-6)
    np.testing.assert_allclose(cm.offset_latitude((2, 0, 0), 0, 675), (0, 0, 2 * d), atol=1e-6)
    np.testing.assert_allclose(cm.offset_latitude((2, 0, 0), 1, 675), (0, 0, d), atol=1e-6)
    np.testing.assert_allclose(cm.offset_latitude((0, -1e-8, 1), 0, 675), (0, d, 0), atol=1e-6)
    # the offsets are small against the tolerance above: check the direction at relative accuracy as well
    np.testing.assert_allclose(cm.offset_longitude((0, -1, 0), 0, 675), (d, 0, 0), atol=1e-12)
    np.testing.assert_allclose(cm.offset_latitude((0, 0, 1), 0, 675), (-d, 0, 0), atol=1e-12)


def _tiles(width, level, dtype, tail=()):
    n = 1 << level
    return np.zeros((2 * n, 4 * n, width, width) + tail, dtype)


def test_pixel_reads_address_the_tile_and_the_pixel_inside_it():
    # t_cubemap.clj:254-273: (dy, dx) = (2 * 675 + 240, 1 * 675 + 320) reads pixel (240, 320) of tile (2, 1)
    width, level = 675, 1
    elev = _tiles(width, level, np.int16)
    day = _tiles(width, level, np.uint8, (4,))
    elev[0, 0, 240, 320] = 42
    elev[2, 1, 240, 320] = 43
    day[0, 0, 240, 320] = (1, 2, 3, 255)
    day[2, 1, 240, 320] = (4, 5, 6, 255)
    w = cm.OracleWorld(width, {level: elev}, {level: day}, {level: day})
    assert w.elevation_pixel(240, 320, level) == 42
    assert w.elevation_pixel(2 * 675 + 240, 1 * 675 + 320, level) == 43
    assert w.world_map_pixel(0, 240, 320, level).tolist() == [1, 2, 3]
    assert w.world_map_pixel(0, 2 * 675 + 240, 1 * 675 + 320, level).tolist() == [4, 5, 6]


def test_interpolation_of_map_pixels():
    # t_cubemap.clj:276-287: x-info [0 1 0.75 0.25], y-info [8 9 0.5 0.5], pixels {[8 0] 2, [8 1] 3, [9 0] 5, [9 1] 7}
    assert cm.interpolate4([2, 3, 5, 7], [0.75, 0.25], [0.5, 0.5]) == 3.875


def test_interpolation_reads_the_four_pixels_map_pixels_name():
    width, level = 8, 0
    elev = _tiles(width, level, np.int16)
    flat = np.arange(2 * width * 4 * width, dtype=np.int16).reshape(2 * width, 4 * width)
    elev[:] = flat.reshape(2, width, 4, width).transpose(0, 2, 1, 3)
    w = cm.OracleWorld(width, {level: elev})
    lon, lat = 0.3, -0.2
    x0, x1, xf0, xf1 = cm.map_pixels_x(lon, width, level)
    y0, y1, yf0, yf1 = cm.map_pixels_y(lat, width, level)
    want = cm.interpolate4([flat[y0, x0], flat[y0, x1], flat[y1, x0], flat[y1, x1]], [xf0, xf1], [yf0, yf1])
    assert w.elevation_geodetic(level, lon, lat) == want
    # wrap-around in longitude (map-pixels-x takes mod size): the last column blends with column 0
    lon = PI - 1e-4
    x0, x1, xf0, xf1 = cm.map_pixels_x(lon, width, level)
    assert (x0, x1) == (4 * width - 1, 0)
    assert w.elevation_geodetic(level, lon, 0.1) == cm.interpolate4(
        [flat[y, x] for y in cm.map_pixels_y(0.1, width, level)[:2] for x in (x0, x1)], [xf0, xf1],
        cm.map_pixels_y(0.1, width, level)[2:])


def test_tile_center():
    # t_cubemap.clj:290-295: tile-center face2 3 7 1 projects (1.0 -0.625 -0.875) onto the sphere
    p = np.array([1.0, -0.625, -0.875])
    assert cm.cube_map(2, cm.cube_coordinate(3, 3, 7, 1.0), cm.cube_coordinate(3, 3, 1, 1.0)).tolist() == p.tolist()
    np.testing.assert_allclose(cm.tile_center(2, 3, 7, 1, R), p / np.linalg.norm(p) * R, rtol=1e-15)


def test_water():
    # t_cubemap.clj:322-358: height 0 -> 0, -500 -> 255, 100 -> 0
    assert cm.water_from_height(0.0) == 0
    assert cm.water_from_height(-500.0) == 255
    assert cm.water_from_height(100.0) == 0
    assert cm.water_from_height(-250.0) == 127      # (int 127.5)
    width, level = 4, 0
    w = cm.OracleWorld(width, {level: np.full((2, 4, width, width), -500, np.int16)})
    assert w.water_geodetic(level, 0.0, 0.0) == 255


def test_project_onto_globe():
    # t_cubemap.clj:361-383: elevation 2777 under (0 0 -1) -> (0 0 -6380777); negative heights are clipped to zero
    width, level = 4, 0
    w = cm.OracleWorld(width, {level: np.full((2, 4, width, width), 2777, np.int16)})
    assert np.linalg.norm(w.project_onto_globe((0, 0, -1), level, R) - np.array([0, 0, -6380777.0])) < 1e-6
    w = cm.OracleWorld(width, {level: np.full((2, 4, width, width), -500, np.int16)})
    assert np.linalg.norm(w.project_onto_globe((1, 0, 0), level, R) - np.array([6378000.0, 0, 0])) < 1e-6


def test_surrounding_points_order():
    # t_cubemap.clj:386-404: offset-longitude => (0 0 -0.1), offset-latitude => (0 0.1 0); point k = 3 (j + 1) + (i + 1)
    pts = cm.surrounding_offsets((1, 0, 0), (0, 0, -0.1), (0, 0.1, 0))
    for j in (-1, 0, 1):
        for i in (-1, 0, 1):
            np.testing.assert_allclose(pts[3 * (j + 1) + (i + 1)], (1, 0.1 * j, -0.1 * i), atol=1e-6)


def test_surrounding_points_project_every_offset_point():
    width, level = 16, 1
    elev, _, _ = cm.synthetic_world(width, [level], [], seed=3)
    w = cm.OracleWorld(width, elev)
    p = np.array([0.3, -0.8, 0.52]) * R
    pts = w.surrounding_points(p, level, 3, 33, R)
    off = cm.surrounding_offsets(p, cm.offset_longitude(p, 3, 33), cm.offset_latitude(p, 3, 33))
    for k in range(9):
        assert pts[k].tolist() == w.project_onto_globe(off[k], level, R).tolist()
    assert w.normal_for_point(p, level, 3, 33, R).tolist() == cm.normal_from_points(pts).tolist()


def test_normal_for_point():
    # t_cubemap.clj:407-427: flat, sloped in longitudinal and in latitudinal direction
    flat = [(6378000, j, -i) for j in (-1, 0, 1) for i in (-1, 0, 1)]
    assert cm.normal_from_points(flat).tolist() == [1, 0, 0]
    lon = [(6378000 + i, j, -i) for j in (-1, 0, 1) for i in (-1, 0, 1)]
    np.testing.assert_allclose(cm.normal_from_points(lon), (np.sqrt(0.5), 0, np.sqrt(0.5)), atol=1e-6)
    lat = [(6378000 + j, j, -i) for j in (-1, 0, 1) for i in (-1, 0, 1)]
    np.testing.assert_allclose(cm.normal_from_points(lat), (np.sqrt(0.5), -np.sqrt(0.5), 0), atol=1e-6)


def test_normal_byte_encoding():
    # image.clj:126-136 spit-normals: round(x * 127.5 - 0.5) as a signed byte
    assert cm.normal_byte(1.0) == 127
    assert cm.normal_byte(-1.0) == -128
    assert cm.normal_byte(0.0) == 0       # Math.round(-0.5) = 0 (ties towards positive infinity)
    assert cm.normal_byte(0.5) == 63


def test_tile_against_the_pointwise_functions():
    """globe.clj:29-80: a tile is nothing but the point-wise functions at the pixel grid, levels clamped as in the source."""
    width = 8
    elev, day, night = cm.synthetic_world(width, [0, 1], [1], seed=5)
    w = cm.OracleWorld(width, elev, day, night)
    face, in_level, out_level, b, a, st = 3, 0, 1, 1, 0, 5
    t = w.make_cube_map_tile(face, in_level, out_level, b, a, surface_tilesize=st)
    ct = 2 * (st - 1) + 1
    assert t["day"].shape == (ct, ct, 4) and t["water"].shape == (ct, 12) and t["surface"].shape == (st, st, 3)
    center = cm.tile_center(face, out_level, b, a, R)
    for v, u in ((0, 0), (2, 3), (4, 4)):
        p = cm.cube_map(face, cm.cube_coordinate(out_level, st, b, v), cm.cube_coordinate(out_level, st, a, u))
        want = (w.project_onto_globe(p, 0, R) - center).astype(np.float32)
        assert t["surface"][v, u].tolist() == want.tolist()
    for v, u in ((0, 0), (5, 7), (8, 8)):
        p = cm.cube_map(face, cm.cube_coordinate(out_level, ct, b, v), cm.cube_coordinate(out_level, ct, a, u))
        point = w.project_onto_globe(p, 0, R)
        lon, lat, _ = cm.cartesian_to_geodetic(point, R)
        assert t["normals"][v, u].tolist() == w.normal_for_point(point, 0, out_level, ct, R).astype(np.float32).tolist()
        assert t["day"][v, u].tolist() == [int(c) for c in w.color_geodetic(0, 1, lon, lat)] + [255]
        assert t["night"][v, u].tolist() == [int(c) for c in w.color_geodetic(1, 1, lon, lat)] + [255]
        assert t["water"][v, u] == w.water_geodetic(1, lon, lat)
    assert (t["water"][:, ct:] == 0).all()
    # neighbouring tiles share their border pixels (the pixel grid includes both tile edges)
    t2 = w.make_cube_map_tile(face, in_level, out_level, b, a + 1, surface_tilesize=st)
    assert (t["day"][:, -1] == t2["day"][:, 0]).all()


def test_oracle_reproduces_the_committed_golden_tiles():
    """tests/golden/cubemap_tiles.npz (tests/golden/make_cubemap_golden.py): the oracle of this checkout must still
    produce the committed tiles -- integer arrays exactly, float arrays to the last bit on the machine that made them
    and within one float32 ulp elsewhere (another libm may differ in the last double bit of atan2 / sin / cos)."""
    import importlib.util
    import os
    here = os.path.join(os.path.dirname(__file__), "golden")
    spec = importlib.util.spec_from_file_location("make_cubemap_golden", os.path.join(here, "make_cubemap_golden.py"))
    gold_mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(gold_mod)
    gold = np.load(os.path.join(here, "cubemap_tiles.npz"))
    elev, day, night = gold_mod.world()
    ow = cm.OracleWorld(gold_mod.WIDTH, elev, day, night)
    for n, t in enumerate(gold_mod.TILES):
        tile = ow.make_cube_map_tile(t["face"], t["in_level"], t["out_level"], t["b"], t["a"],
                                     surface_tilesize=gold_mod.SURFACE_TILESIZE)
        for k in ("day", "night", "water"):
            diff = tile[k].astype(np.int32) - gold["%d_%s" % (n, k)].astype(np.int32)
            assert np.abs(diff).max() <= 1 and (diff != 0).mean() < 1e-3, (n, k)
        for k in ("surface", "normals"):
            want = gold["%d_%s" % (n, k)]
            assert np.abs(tile[k].astype(np.float64) - want).max() <= float(np.spacing(np.abs(want).max())), (n, k)
