"""Sharded build on 2+ GPUs == single-GPU build, byte for byte (skipped on a one-GPU box; the CPU-side
world_size-2 logic is covered by tests/test_sharding_gloo.py)."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.parametrize("extra", [["--configs", "reduced", "--modes", "p2p"], ["--configs", "shipped", "--modes", "p2p"],
                                   ["--configs", "reduced", "--modes", "nccl"], ["--configs", "shipped", "--modes", "nccl"]])
def test_sharded_build_is_identical(extra):
    import torch
    n = torch.cuda.device_count()
    if n < 2:
        pytest.skip("needs at least 2 GPUs")
    world = 2 if n < 4 else 4
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(world),
           "--master-addr", "127.0.0.1", "--master-port", "29533", os.path.join(ROOT, "tools", "multi_gpu_check.py")] + extra
    res = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    assert res.returncode == 0, res.stdout[-2000:] + res.stderr[-2000:]
    assert "MULTI_GPU_CHECK OK" in res.stdout


def test_single_process_multi_gpu_build_is_identical():
    """atmlut_generate_multi: one process drives several GPUs (the form the Clojure host uses)."""
    import numpy as np
    import torch
    from sfsim_b200 import _lib, atmosphere_lut
    n = min(torch.cuda.device_count(), 8)
    if n < 2:
        pytest.skip("needs at least 2 GPUs")
    for cfg in (_lib.make_config(ray_scatter_shape=(5, 7, 8, 2), transmittance_shape=(8, 15),
                                 surface_radiance_shape=(4, 7), ray_steps=20, sphere_steps=8, iterations=2),
                _lib.default_config()):
        want = atmosphere_lut.generate_tables(cfg=cfg)
        for _ in range(2):                                   # second call reuses the cached group
            got = atmosphere_lut.generate_tables(cfg=cfg, num_gpus=n)
            for name, g, w in zip(atmosphere_lut.FILE_NAMES, got, want):
                assert np.array_equal(g, w), name
