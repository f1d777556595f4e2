"""Short run for profilers: three full shipped-resolution builds on one GPU, launched kernel by kernel
(no timing claims).  `stress` as first argument profiles the stress configuration instead."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from sfsim_b200 import _lib, atmosphere_lut  # noqa: E402

cfg = None
if len(sys.argv) > 1 and sys.argv[1] == "stress":
    cfg = _lib.make_config(ray_scatter_shape=(64, 253, 64, 16), iterations=2)
b = atmosphere_lut.AtmosphereLutBuilder(cfg=cfg)
for _ in range(3):
    b.run_timed()
    b.sync()
b.close()
