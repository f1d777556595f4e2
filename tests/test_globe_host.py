"""Host-only tests of the sfsim.globe / sfsim.util mirror (sfsim_b200/globe.py): file names and formats of the cube-map
tile files, against the reference's facts (t_util.clj:34-131) and, where the reference checkout is present, against
its fixtures."""
import os
import tarfile

import numpy as np
import pytest

from sfsim_b200 import globe

FIXTURES = "/root/reference/test/clj/sfsim/fixtures/util"


def test_paths():
    # t_util.clj:115-131
    assert globe.tile_path("world", 1, 3, 2, ".png") == "world/1/2/3.png"
    assert globe.cube_path("globe", 5, 2, 3, 1, ".png") == "globe/5/2/1/3.png"
    assert globe.cube_dir("globe", 5, 2, 1) == "globe/5/2/1"
    assert globe.cube_tar("globe", 5, 2, 1) == "globe/5/2/1.tar"


def test_gz_round_trips(tmp_path):
    # t_util.clj:76-78,94-96
    name = str(tmp_path / "spit.gz")
    globe.spit_bytes_gz(name, [2, 3, 5, 7])
    assert globe.slurp_bytes_gz(name).tolist() == [2, 3, 5, 7]
    globe.spit_floats_gz(name, [2.0, 3.0, 5.0, 7.0])
    assert globe.slurp_floats_gz(name).tolist() == [2.0, 3.0, 5.0, 7.0]


@pytest.mark.skipif(not os.path.isdir(FIXTURES), reason="the reference checkout is only present in the build container")
def test_reads_the_reference_fixtures():
    # t_util.clj:38,58
    assert globe.slurp_bytes_gz(os.path.join(FIXTURES, "bytes.gz")).tolist() == [2, 3, 5, 7]
    assert globe.slurp_floats_gz(os.path.join(FIXTURES, "floats.gz")).tolist() == [2.0, 3.0, 5.0, 7.0]


def test_normals_round_trip_like_the_reference(tmp_path):
    # image.clj:126-156: spit-normals then slurp-normals returns the vector to within half a byte step
    rng = np.random.default_rng(1)
    n = rng.normal(size=(9, 9, 3))
    n = (n / np.linalg.norm(n, axis=-1, keepdims=True)).astype(np.float32)
    path = str(tmp_path / "n.png")
    globe.spit_normals(path, normals=n)
    back = globe.slurp_normals(path)
    assert np.abs(back - n).max() <= 0.5 / 127.5 + 1e-6
    # the bytes the library delivers give the same file
    scaled = n.astype(np.float64) * 127.5
    globe.spit_normals(str(tmp_path / "b.png"), normal_bytes=np.floor((scaled - 0.5) + 0.5).astype(np.int8))
    assert (globe.slurp_normals(str(tmp_path / "b.png")) == back).all()


def test_a_tile_lands_in_the_files_of_the_reference_layout(tmp_path):
    ct, st = 9, 5
    rng = np.random.default_rng(2)
    tile = {"day": rng.integers(0, 256, (ct, ct, 4), dtype=np.uint8), "night": rng.integers(0, 256, (ct, ct, 4), dtype=np.uint8),
            "water": rng.integers(0, 256, (ct, 12), dtype=np.uint8), "surface": rng.normal(size=(st, st, 3)).astype(np.float32),
            "normal_bytes": rng.integers(-128, 128, (ct, ct, 3), dtype=np.int8)}
    prefix = str(tmp_path / "globe")
    globe.write_cube_map_tile(prefix, 3, 2, 1, 0, tile)
    d = os.path.join(prefix, "3", "2", "0")
    assert sorted(os.listdir(d)) == ["1.jpg", "1.night.jpg", "1.png", "1.surf.gz", "1.water.gz"]     # globe.clj:74-78
    assert globe.slurp_bytes_gz(os.path.join(d, "1.water.gz")).tobytes() == tile["water"].tobytes()
    assert globe.slurp_floats_gz(os.path.join(d, "1.surf.gz")).tobytes() == tile["surface"].tobytes()
    from PIL import Image
    assert np.asarray(Image.open(os.path.join(d, "1.png"))).view(np.int8).tobytes() == tile["normal_bytes"].tobytes()
    assert Image.open(os.path.join(d, "1.jpg")).size == (ct, ct)
    globe.make_cube_map_tars(2, prefix)                                                               # globe.clj:83-96
    assert not os.path.exists(d)
    with tarfile.open(os.path.join(prefix, "3", "2", "0.tar")) as tar:
        assert sorted(tar.getnames()) == ["1.jpg", "1.night.jpg", "1.png", "1.surf.gz", "1.water.gz"]
