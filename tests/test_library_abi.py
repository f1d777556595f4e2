"""CPU-only checks of the C-ABI library: it loads, exports every symbol the headers under include/ declare,
its host-only helpers work, and the compute entry points fail loudly without a CUDA device (no fallback)."""
import ctypes as C
import os
import re
import struct

import numpy as np
import pytest

from oracle import oracle as orc
from sfsim_b200 import _lib, atmosphere_lut, sharding

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    """every function the headers under include/ declare (atmlut_* in sfsim_atmosphere.h, sfsim_* in sfsim_noise.h)"""
    names = set()
    for header in sorted(os.listdir(os.path.join(ROOT, "include"))):
        text = open(os.path.join(ROOT, "include", header)).read()
        text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
        names |= set(re.findall(r"\b((?:atmlut|sfsim)_\w+)\s*\(", text))
    return sorted(names)


def test_library_exports_every_declared_symbol():
    lib = _lib.load()
    declared = declared_symbols()
    assert len(declared) >= 30
    for name in declared:
        assert hasattr(lib, name), name
    # the Python binding's list is the header's list
    assert sorted(_lib.EXPORTS) == declared


def test_struct_layouts_match_the_header():
    assert C.sizeof(_lib.Planet) == 8 * 8
    assert C.sizeof(_lib.Scatter) == 6 * 8
    assert C.sizeof(_lib.Config) == 11 * 4 + 4 + 3 * 8       # 11 ints, padding, 3 doubles
    assert _lib.Config.intensity.offset == 48


def test_default_config_is_the_shipped_one():
    cfg = _lib.default_config()                                 # atmosphere_lut.clj:47-63
    assert cfg.ray_scatter_shape == (32, 127, 32, 8)
    assert cfg.transmittance_shape == (64, 255)
    assert cfg.surface_radiance_shape == (16, 63)
    assert (cfg.ray_steps, cfg.sphere_steps, cfg.iterations) == (100, 15, 5)
    assert list(cfg.intensity) == [1.0, 1.0, 1.0]
    shapes = atmosphere_lut.output_shapes(cfg)
    assert shapes == ((64, 255, 3), (16, 63, 3), (1024, 1016, 3), (1024, 1016, 3))
    assert [int(np.prod(s)) * 4 for s in shapes] == [195840, 12096, 12484608, 12484608]


def test_constants_match_the_reference():
    assert atmosphere_lut.radius == 6378000.0 and atmosphere_lut.height == 35000.0           # atmosphere_lut.clj:20-21
    assert atmosphere_lut.mie == {"base": (2e-5, 2e-5, 2e-5), "scale": 1200.0, "g": 0.76, "quotient": 0.9}
    assert atmosphere_lut.rayleigh == {"base": (5.8e-6, 13.5e-6, 33.1e-6), "scale": 8000.0}
    assert atmosphere_lut.FILE_NAMES == ("transmittance.scatter", "surface-radiance.scatter", "ray-scatter.scatter",
                                         "mie-strength.scatter")


def test_float_files_round_trip(tmp_path):
    """spit-floats / slurp-floats (util.clj:188-240): headerless little-endian float32."""
    lib = _lib.load()
    data = np.array([2.0, 3.0, 5.0, 7.0], dtype=np.float32)
    path = str(tmp_path / "floats.raw")
    assert lib.atmlut_write_floats(path.encode(), _lib.ptr(data), C.c_long(4)) == 0
    golden = os.path.join(ROOT, "tests", "golden", "floats.raw")
    assert open(path, "rb").read() == open(golden, "rb").read() == struct.pack("<4f", 2.0, 3.0, 5.0, 7.0)
    back = np.zeros(8, dtype=np.float32)
    assert lib.atmlut_read_floats(golden.encode(), _lib.ptr(back), C.c_long(8)) == 4
    np.testing.assert_array_equal(back[:4], data)
    assert lib.atmlut_write_floats(b"/nonexistent-dir/x.raw", _lib.ptr(data), C.c_long(4)) != 0
    assert b"cannot open" in lib.atmlut_last_error()


def test_convert_4d_to_2d_host_helper():
    """image.clj:299-312, the two cases of t_image.clj:187-191 plus a random RGB table against the oracle."""
    lib = _lib.load()

    def convert(a, ncomp=1):
        a = np.ascontiguousarray(a, dtype=np.float32)
        d, c, b, aa = a.shape[:4]
        out = np.zeros((d * b, c * aa) + (() if ncomp == 1 else (ncomp,)), dtype=np.float32)
        shape = (C.c_int * 4)(d, c, b, aa)
        assert lib.atmlut_convert_4d_to_2d(_lib.ptr(a), shape, ncomp, _lib.ptr(out)) == 0
        return out

    np.testing.assert_array_equal(convert(np.arange(1, 17).reshape(2, 2, 2, 2)),
                                  [[1, 2, 5, 6], [3, 4, 7, 8], [9, 10, 13, 14], [11, 12, 15, 16]])
    np.testing.assert_array_equal(convert(np.arange(1, 25).reshape(1, 2, 3, 4)),
                                  [[1, 2, 3, 4, 13, 14, 15, 16], [5, 6, 7, 8, 17, 18, 19, 20],
                                   [9, 10, 11, 12, 21, 22, 23, 24]])
    rng = np.random.default_rng(0)
    t = rng.random((3, 5, 4, 2, 3)).astype(np.float32)
    np.testing.assert_array_equal(convert(t, 3), orc.convert_4d_to_2d(t.astype(np.float64)).astype(np.float32))


def test_quadrature_tables_match_the_oracle():
    """sphere.clj:70-105: the direction lists the kernels integrate over are the reference's, bit for bit
    (ring counts ceil(sin(theta) * steps) sit on an integer at theta = pi/2)."""
    lib = _lib.load()
    for steps, half, theta_steps, theta_range in [(15, 0, 7, np.pi), (100, 1, 25, np.pi / 2), (6, 0, 3, np.pi),
                                                  (64, 0, 32, np.pi), (64, 1, 16, np.pi / 2), (7, 1, 1, np.pi / 2)]:
        n = lib.atmlut_sphere_directions(steps, half, None, None, 0)
        dirs = np.zeros((n, 3))
        weights = np.zeros(n)
        assert lib.atmlut_sphere_directions(steps, half, _lib.ptr(dirs), _lib.ptr(weights), n) == n
        want_dirs, want_w = orc.sphere_directions(theta_steps, steps, theta_range, (1, 0, 0))
        assert n == len(want_w)
        np.testing.assert_array_equal(dirs, want_dirs)
        np.testing.assert_array_equal(weights, want_w)
    assert lib.atmlut_sphere_directions(15, 0, None, None, 0) == 71
    assert lib.atmlut_sphere_directions(100, 1, None, None, 0) == 1605


def test_slab_partition():
    """N4 pairs split in equal contiguous slabs; the library and its mirror agree."""
    for n_pairs in (4064, 16256, 7, 1, 250):
        for world in (1, 2, 3, 4, 8):
            covered = []
            for rank in range(world):
                got = sharding.slab_from_library(n_pairs, rank, world)
                assert got == sharding.slab(n_pairs, rank, world)
                begin, count, per_rank = got
                assert per_rank * world >= n_pairs and count <= per_rank
                covered.extend(range(begin, begin + count))
            assert covered == list(range(n_pairs))
    assert sharding.slab(4064, 3, 8) == (1524, 508, 508)          # shipped: 32 * 127 = 8 * 508


@pytest.mark.skipif(_lib.load().atmlut_device_count() > 0, reason="a CUDA device is present")
def test_no_cpu_fallback_without_a_device():
    """Without a GPU every compute entry point fails with a message; nothing is computed on the host."""
    lib = _lib.load()
    assert lib.atmlut_init(0) != 0
    assert b"no CPU fallback" in lib.atmlut_last_error()
    with pytest.raises(_lib.AtmlutError, match="no CPU fallback"):
        atmosphere_lut.generate_tables(cfg=_lib.make_config(ray_scatter_shape=(2, 2, 2, 2), iterations=0))
    from sfsim_b200 import atmosphere
    with pytest.raises(_lib.AtmlutError):
        atmosphere.transmittance(atmosphere_lut.earth, [atmosphere_lut.rayleigh], 10, (0, 6378000.0, 0), (0, 1, 0), True)


def test_header_is_plain_c_and_links(tmp_path):
    """A C99 program includes the header, links the shared library and calls the host-only entry points
    (what a JNI / FFM / cgo consumer does)."""
    import subprocess
    exe = str(tmp_path / "c_abi_smoke")
    subprocess.check_call(["gcc", "-std=c99", "-Wall", "-Werror", "-pedantic", "-I", os.path.join(ROOT, "include"),
                           os.path.join(ROOT, "tests", "c_abi_smoke.c"), "-o", exe, "-L", ROOT, "-lsfsim_atmosphere",
                           "-Wl,-rpath," + ROOT])
    out = subprocess.run([exe], cwd=str(tmp_path), capture_output=True, text=True)
    assert out.returncode == 0, out.stdout + out.stderr
    assert "devices=" in out.stdout


def test_missing_library_fails_loudly(monkeypatch, tmp_path):
    monkeypatch.setattr(_lib, "_lib", None)
    monkeypatch.setattr(_lib, "LIB_PATH", str(tmp_path / "libsfsim_atmosphere.so"))
    with pytest.raises(_lib.AtmlutError, match="no CPU fallback"):
        _lib.load()


def test_product_package_does_not_touch_the_oracle():
    """Only tests/, __graft_entry__.smoke() and bench.py's CPU-baseline legs may use oracle/."""
    pkg = os.path.join(ROOT, "sfsim_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                text = open(os.path.join(dirpath, f)).read()
                assert "oracle" not in text.lower().replace("no oracle", ""), os.path.join(dirpath, f)


def test_hot_kernel_uses_blackwell_packed_fp32_and_mufu():
    """SASS of the built library: the first-order kernel's sampler runs on FFMA2/FMUL2/FADD2 (sm_100 packed FP32),
    MUFU.EX2 and DADD, and the library is compiled for sm_100a only."""
    import shutil
    import subprocess
    cuobjdump = shutil.which("cuobjdump") or "/usr/local/cuda/bin/cuobjdump"
    if not os.path.exists(cuobjdump):
        pytest.skip("cuobjdump not available")
    lib = os.path.join(ROOT, "libsfsim_atmosphere.so")
    arch = subprocess.run([cuobjdump, "-lelf", lib], capture_output=True, text=True).stdout
    assert "sm_100a" in arch and "sm_90" not in arch and "sm_80" not in arch
    sass = subprocess.run([cuobjdump, "-sass", lib], capture_output=True, text=True).stdout
    start = sass.index("k_first_order")
    end = sass.find("Function :", start + 1)
    body = sass[start:end if end > 0 else len(sass)]
    for mnemonic in ("FFMA2", "FMUL2", "FADD2", "MUFU.EX2", "DADD"):
        assert mnemonic in body, mnemonic
    assert body.count("MUFU.EX2") >= 8
