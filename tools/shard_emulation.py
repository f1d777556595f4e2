"""Stage times of ONE rank's share of a sharded build on a single GPU: rank 0 of `world` in the all-gather mode with a
callback that gathers nothing (the tables stay incomplete, so the numbers are timings only).  Used to tune the kernels
for the small grids of an 8-GPU shard without occupying eight GPUs.

    python tools/shard_emulation.py [world] [shipped|stress]
"""
import ctypes as C
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from sfsim_b200 import _lib, atmosphere_lut  # noqa: E402

world = int(sys.argv[1]) if len(sys.argv) > 1 else 8
workload = sys.argv[2] if len(sys.argv) > 2 else "shipped"
cfg = _lib.default_config() if workload == "shipped" else _lib.make_config(ray_scatter_shape=(64, 253, 64, 16), iterations=10)
lib = _lib.load()
pl = _lib.make_planet(atmosphere_lut.earth)
sc = _lib.make_scatter_array((atmosphere_lut.mie, atmosphere_lut.rayleigh))
handle = C.c_void_p()
_lib.check(lib.atmlut_builder_create(C.byref(pl), sc, 2, C.byref(cfg), 0, world, C.byref(handle)))
noop = _lib.ALLGATHER_FN(lambda user, buf, nbytes, stream: 0)
_lib.check(lib.atmlut_builder_set_allgather(handle, noop, None))
for _ in range(3):
    _lib.check(lib.atmlut_builder_run_timed(handle))
    _lib.check(lib.atmlut_builder_sync(handle))
stages = {}
for i in range(lib.atmlut_builder_stage_count(handle)):
    ms = C.c_float()
    _lib.check(lib.atmlut_builder_stage_ms(handle, i, C.byref(ms)))
    name = lib.atmlut_builder_stage_name(handle, i).decode()
    stages[name] = stages.get(name, 0.0) + ms.value
print("world %d %s K4_CONSUMERS=%s:" % (world, workload, os.environ.get("ATMLUT_K4_CONSUMERS", "auto")),
      {k: round(v, 3) for k, v in stages.items() if not k.endswith("exchange")}, "sum %.3f ms" % sum(stages.values()))
lib.atmlut_builder_destroy(handle)
