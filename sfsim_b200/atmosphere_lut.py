"""Host-side mirror of sfsim.atmosphere-lut (src/clj/sfsim/atmosphere_lut.clj) over the C ABI.

`generate_atmosphere_luts` is the drop-in for `clj -T:build atmosphere-lut` (build.clj:84-87): it
produces the four `data/atmosphere/*.scatter` files with the same names, shapes and byte layout.
`AtmosphereLutBuilder` is the device-resident form used for benchmarking and for sharding one
build over the GPUs of a box (one process per GPU, torch.distributed/NCCL all-gather per table).
"""
import ctypes as C
import os

import numpy as np

from . import _lib
from ._lib import check

# atmosphere_lut.clj:20-40
radius = 6378000.0
height = 35000.0
earth = {"centre": (0.0, 0.0, 0.0), "radius": radius, "height": height, "brightness": (0.3, 0.3, 0.3)}
mie = {"base": (2e-5, 2e-5, 2e-5), "scale": 1200.0, "g": 0.76, "quotient": 0.9}
rayleigh = {"base": (5.8e-6, 13.5e-6, 33.1e-6), "scale": 8000.0}

FILE_NAMES = ("transmittance.scatter", "surface-radiance.scatter", "ray-scatter.scatter", "mie-strength.scatter")


def output_shapes(cfg):
    """File-layout shapes of the four outputs."""
    h, e, s, a = cfg.ray_scatter_shape
    return (cfg.transmittance_shape + (3,), cfg.surface_radiance_shape + (3,), (h * s, e * a, 3), (h * s, e * a, 3))


def allocate_outputs(cfg, pinned=False):
    shapes = output_shapes(cfg)
    if pinned:
        import torch
        return [torch.empty(s, dtype=torch.float32).pin_memory().numpy() for s in shapes]
    return [np.empty(s, dtype=np.float32) for s in shapes]


def generate_tables(planet=earth, scatter=(mie, rayleigh), cfg=None, out=None, num_gpus=1):
    """generate-atmosphere-luts up to (not including) the file writes; returns the four float32 arrays
    in file layout.  One call into the library: atmlut_generate, or atmlut_generate_multi when the build
    is to be spread over `num_gpus` GPUs from this one process."""
    lib = _lib.load()
    cfg = cfg or _lib.default_config()
    out = out or allocate_outputs(cfg)
    pl = _lib.make_planet(planet)
    sc = _lib.make_scatter_array(scatter)
    if num_gpus == 1:
        check(lib.atmlut_generate(C.byref(pl), sc, len(scatter), C.byref(cfg), *[_lib.ptr(o) for o in out]))
    else:
        check(lib.atmlut_generate_multi(C.byref(pl), sc, len(scatter), C.byref(cfg), int(num_gpus),
                                        *[_lib.ptr(o) for o in out]))
    return out


def write_tables(tables, out_dir):
    """spit-floats of the four tables (atmosphere_lut.clj:102-105)."""
    lib = _lib.load()
    os.makedirs(out_dir, exist_ok=True)
    paths = []
    for name, t in zip(FILE_NAMES, tables):
        path = os.path.join(out_dir, name)
        t = _lib.f32(t)
        check(lib.atmlut_write_floats(path.encode(), _lib.ptr(t), C.c_long(t.size)))
        paths.append(path)
    return paths


def generate_atmosphere_luts(out_dir="data/atmosphere", planet=earth, scatter=(mie, rayleigh), cfg=None, num_gpus=1):
    """Program to generate lookup tables for atmospheric scattering (atmosphere_lut.clj:43-105)."""
    return write_tables(generate_tables(planet, scatter, cfg, num_gpus=num_gpus), out_dir)


class AtmosphereLutBuilder:
    """Device-resident build.  With world > 1 the 4-D tables are sharded over the ranks (one process per GPU):

    * mode "p2p" (default on GPUs): ranks exchange CUDA IPC handles once through torch.distributed, the kernels
      store every finished texel straight into all GPUs' tables over NVLink and a flag barrier closes each table;
    * mode "nccl": each rank fills a contiguous slab and torch.distributed all-gathers every table (NCCL)."""

    def __init__(self, planet=earth, scatter=(mie, rayleigh), cfg=None, rank=0, world=1, device=None,
                 process_group=None, mode="p2p"):
        self.lib = _lib.load()
        self.cfg = cfg or _lib.default_config()
        self.rank, self.world = rank, world
        if device is not None:
            check(self.lib.atmlut_init(int(device)))
        pl = _lib.make_planet(planet)
        sc = _lib.make_scatter_array(scatter)
        self.handle = C.c_void_p()
        check(self.lib.atmlut_builder_create(C.byref(pl), sc, len(scatter), C.byref(self.cfg), rank, world,
                                             C.byref(self.handle)))
        self._callback = None
        self.gathers = 0
        self.mode = mode if world > 1 else "single"
        if world > 1:
            if mode == "p2p":
                self._exchange_ipc_handles(process_group)
            elif mode == "nccl":
                self._install_allgather(process_group)
            else:
                raise ValueError("mode must be 'p2p' or 'nccl'")

    def _exchange_ipc_handles(self, process_group):
        import torch.distributed as dist
        mine = C.create_string_buffer(self.lib.atmlut_builder_ipc_handle_bytes())
        check(self.lib.atmlut_builder_ipc_export(self.handle, mine, len(mine)))
        everyone = [None] * self.world
        dist.all_gather_object(everyone, bytes(mine.raw), group=process_group)
        check(self.lib.atmlut_builder_ipc_import(self.handle, b"".join(everyone), self.world))
        dist.barrier(group=process_group)

    def _install_allgather(self, process_group):
        from . import sharding

        def allgather(_user, buf, bytes_per_rank, stream):
            try:
                sharding.allgather_device_table(buf, bytes_per_rank, self.rank, self.world, stream, process_group)
                self.gathers += 1
                return 0
            except Exception as exc:  # surfaces as "allgather callback failed" from the library
                import traceback
                traceback.print_exc()
                self._error = exc
                return 1

        self._callback = _lib.ALLGATHER_FN(allgather)
        check(self.lib.atmlut_builder_set_allgather(self.handle, self._callback, None))

    def run(self):
        """Enqueue one full build on the builder's stream (asynchronous; one CUDA graph launch after the first run)."""
        check(self.lib.atmlut_builder_run(self.handle))

    def run_timed(self):
        """The same build with per-stage events recorded on the device (see stage_times)."""
        check(self.lib.atmlut_builder_run_timed(self.handle))

    def set_option(self, option, value):
        check(self.lib.atmlut_builder_set_option(self.handle, int(option), int(value)))

    @property
    def stream(self):
        """cudaStream_t of the build (an integer handle)."""
        return self.lib.atmlut_builder_stream(self.handle)

    def sync(self):
        check(self.lib.atmlut_builder_sync(self.handle))

    def download(self, out=None):
        out = out or allocate_outputs(self.cfg)
        check(self.lib.atmlut_builder_download(self.handle, *[_lib.ptr(o) for o in out]))
        return out

    def stage_times(self):
        """[(stage name, device milliseconds)] of the last run_timed()."""
        n = self.lib.atmlut_builder_stage_count(self.handle)
        res = []
        for i in range(n):
            ms = C.c_float()
            check(self.lib.atmlut_builder_stage_ms(self.handle, i, C.byref(ms)))
            res.append((self.lib.atmlut_builder_stage_name(self.handle, i).decode(), ms.value))
        return res

    def work(self):
        e, l4, l2 = C.c_double(), C.c_double(), C.c_double()
        check(self.lib.atmlut_builder_work(self.handle, C.byref(e), C.byref(l4), C.byref(l2)))
        res = {"esamples": e.value, "lookups4d": l4.value, "lookups2d": l2.value}
        for which, key in enumerate(("esamples_first_order", "esamples_ray_scatter", "kernel_launches", "mufu_ex2_per_esample")):
            v = C.c_double()
            check(self.lib.atmlut_builder_counter(self.handle, which, C.byref(v)))
            res[key] = v.value
        return res

    def close(self):
        if self.handle:
            self.lib.atmlut_builder_destroy(self.handle)
            self.handle = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
