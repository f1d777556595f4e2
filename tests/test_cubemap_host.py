"""Host-only tests of the sfsim.cubemap mirror (sfsim_b200/cubemap.py): the raster-free coordinate functions against the
oracle, the struct layout, tile sharding (also across two gloo ranks) and the no-device behaviour of the entry points."""
import ctypes as C
import os
import subprocess
import sys

import numpy as np
import pytest

from oracle import cubemap as ocm
from sfsim_b200 import _lib, cubemap

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
R = 6378000.0


def test_coordinate_functions_equal_the_oracle_bit_for_bit():
    rng = np.random.default_rng(11)
    for _ in range(200):
        face = int(rng.integers(0, 6))
        j, i = rng.random(2)
        assert cubemap.cube_map(face, j, i).tolist() == ocm.cube_map(face, j, i).tolist()
        p = rng.normal(size=3)
        f = cubemap.determine_face(p)
        assert f == ocm.determine_face(p)
        assert cubemap.cube_i(f, p) == ocm.cube_i(f, p) and cubemap.cube_j(f, p) == ocm.cube_j(f, p)
        assert cubemap.project_onto_cube(p).tolist() == ocm.project_onto_cube(p).tolist()
        assert cubemap.project_onto_sphere(p, R).tolist() == ocm.project_onto_sphere(p, R).tolist()
        assert cubemap.longitude(p) == ocm.longitude(p) and cubemap.latitude(p) == ocm.latitude(p)
        assert cubemap.cartesian_to_geodetic(p * R, R) == ocm.cartesian_to_geodetic(p * R, R).tolist()
        lon, lat = rng.uniform(-np.pi, np.pi), rng.uniform(-np.pi / 2, np.pi / 2)
        assert cubemap.geodetic_to_cartesian(lon, lat, 100.0, R).tolist() == ocm.geodetic_to_cartesian(lon, lat, 100.0, R).tolist()
        level = int(rng.integers(0, 6))
        assert cubemap.map_pixels_x(lon, 675, level) == ocm.map_pixels_x(lon, 675, level)
        assert cubemap.map_pixels_y(lat, 675, level) == ocm.map_pixels_y(lat, 675, level)
        tile = int(rng.integers(0, 1 << level))
        pixel = float(rng.integers(0, 129))
        assert cubemap.cube_coordinate(level, 129, tile, pixel) == ocm.cube_coordinate(level, 129, tile, pixel)
        assert cubemap.tile_center(face, level, tile, tile, R).tolist() == ocm.tile_center(face, level, tile, tile, R).tolist()
    assert [c.tolist() for c in cubemap.cube_map_corners(5, 2, 3, 1)] == ocm.cube_map_corners(5, 2, 3, 1).tolist()


def test_reference_facts_hold_for_the_mirror():
    # t_cubemap.clj:95-96,115-118,198-219
    assert cubemap.cube_map(5, 0.0, 0.5).tolist() == [0.0, -1.0, -1.0]
    assert cubemap.cube_coordinate(1, 256, 1, 127.5) == 0.75
    assert cubemap.map_x(0.0, 675, 3) == 675 * 2.0 * 8 and cubemap.map_y(0.0, 675, 3) == 675 * 8.0
    assert cubemap.map_pixels_x(np.pi - np.pi / (256 * 4), 256, 0) == [256 * 4 - 1, 0, 0.5, 0.5]
    assert cubemap.map_pixels_y(-np.pi / 2, 675, 3) == [675 * 2 * 8 - 1, 675 * 2 * 8 - 1, 1.0, 0.0]


def test_config_defaults_are_the_constants_of_make_cube_map():
    cfg = cubemap.make_config(-3, 0)                      # globe.clj:32-40, build.clj:300-302
    assert (cfg.width, cfg.surface_tilesize, cfg.sublevel, cfg.max_surface_level, cfg.max_color_level) == (675, 65, 1, 4, 5)
    assert cfg.radius == 6378000.0 and cfg.color_tilesize == 129
    assert C.sizeof(cubemap.CubemapConfig) == 7 * 4 + 4 + 8
    with pytest.raises(TypeError):
        cubemap.make_config(0, 0, no_such_field=1)


def test_tile_shards_partition_the_level_in_reference_order():
    n = 4
    every = [(k, b, a) for k in range(6) for b in range(n) for a in range(n)]      # globe.clj:41
    assert [tuple(t) for t in cubemap.tile_shard(2)] == every
    parts = [cubemap.tile_shard(2, r, 8) for r in range(8)]
    assert sorted(tuple(t) for p in parts for t in p) == every
    assert [len(p) for p in parts] == [12] * 8
    assert [tuple(t) for t in parts[3][:2]] == [every[3], every[11]]
    assert len(cubemap.tile_shard(0, 7, 8)) == 0                  # 6 tiles on 8 ranks: the last two have nothing to do
    with pytest.raises(_lib.AtmlutError):
        cubemap.tile_shard(2, 8, 8)


def test_no_cpu_fallback_without_a_device():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a device is present")
    with pytest.raises(_lib.AtmlutError, match="no CUDA device"):
        cubemap.World(16)


GLOO_WORKER = r"""
import os, sys
sys.path.insert(0, sys.argv[1])
import numpy as np
import torch
import torch.distributed as dist
from sfsim_b200 import cubemap
dist.init_process_group("gloo", init_method="tcp://127.0.0.1:%s" % sys.argv[2], rank=int(sys.argv[3]), world_size=2)
rank = dist.get_rank()
mine = cubemap.tile_shard(1, rank, 2)                 # 24 tiles over 2 ranks, no data-path collective
keys = torch.tensor([(k * 2 + b) * 2 + a for k, b, a in mine], dtype=torch.int64)
both = [torch.zeros(12, dtype=torch.int64) for _ in range(2)]
dist.all_gather(both, keys)                           # test-only: proves the shards are disjoint and complete
if rank == 0:
    assert sorted(torch.cat(both).tolist()) == list(range(24))
    print("SHARDS_OK")
dist.destroy_process_group()
"""


def test_two_ranks_split_a_level(tmp_path):
    script = tmp_path / "worker.py"
    script.write_text(GLOO_WORKER)
    port = str(29500 + os.getpid() % 1000)
    procs = [subprocess.Popen([sys.executable, str(script), ROOT, port, str(r)], stdout=subprocess.PIPE,
                              stderr=subprocess.STDOUT, text=True) for r in range(2)]
    outs = [p.communicate(timeout=120)[0] for p in procs]
    assert all(p.returncode == 0 for p in procs), outs
    assert "SHARDS_OK" in outs[0]
