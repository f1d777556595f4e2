"""Short run for profilers: two full shipped-resolution builds on one GPU (no timing claims)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from sfsim_b200 import atmosphere_lut  # noqa: E402

b = atmosphere_lut.AtmosphereLutBuilder()
for _ in range(2):
    b.run()
    b.sync()
b.close()
