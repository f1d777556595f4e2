"""Host side of the all-gather exchange mode: one atmosphere-LUT build sharded over the GPUs of a box.

Within a table every texel is independent; between tables the next kernel needs the whole previous table (lookups
are scattered).  In this mode each rank integrates a contiguous slab of (height, elevation) pairs in place in the
full-size (padded) table, and one all-gather per table reassembles it.  The slab arithmetic lives in the
library (atmlut_slab, mirrored by `slab` so that it can be checked without a device); `allgather_table` is the
collective `AtmosphereLutBuilder` issues from the library's all-gather callback -- on NCCL for device tables, and on
gloo in the CPU test of exactly this function (tests/test_sharding_gloo.py).
"""
import ctypes as C

from . import _lib


def slab(n_pairs, rank, world):
    """(begin, count, per_rank) of `rank`; identical to the C library's atmlut_slab."""
    per_rank = (n_pairs + world - 1) // world
    begin = rank * per_rank
    count = max(0, min(n_pairs, begin + per_rank) - begin)
    return begin, count, per_rank


def slab_from_library(n_pairs, rank, world):
    lib = _lib.load()
    b, c, p = C.c_int(), C.c_int(), C.c_int()
    _lib.check(lib.atmlut_slab(n_pairs, rank, world, C.byref(b), C.byref(c), C.byref(p)))
    return b.value, c.value, p.value


def padded_pairs(n_pairs, world):
    return slab(n_pairs, 0, world)[2] * world


def allgather_table(full, rank, world, group=None):
    """In-place all-gather of a padded table tensor `full` (flat, or [world * per_rank * texels_per_pair, 4]):
    rank's slab is already in place; afterwards every rank holds every slab."""
    import torch.distributed as dist
    flat = full.view(-1)
    n = flat.numel() // world
    mine = flat[rank * n:(rank + 1) * n]
    if flat.device.type == "cpu":
        mine = mine.clone()          # gloo does not accept an input that aliases the output
    dist.all_gather_into_tensor(flat, mine, group=group)
    return full


class RawDeviceRange:
    """__cuda_array_interface__ view of a raw device range, so that torch can wrap one of the library's tables."""

    def __init__(self, ptr, nbytes):
        self.__cuda_array_interface__ = {"shape": (nbytes // 4,), "typestr": "<f4", "data": (ptr, False), "version": 3,
                                         "strides": None}


def allgather_device_table(ptr, bytes_per_rank, rank, world, stream, group=None):
    """The library's all-gather callback body: `ptr` is the padded device table (world * bytes_per_rank bytes) whose
    slab `rank` was just produced on the CUDA stream `stream`."""
    import torch
    ext = torch.cuda.ExternalStream(stream)
    with torch.cuda.stream(ext):
        full = torch.as_tensor(RawDeviceRange(ptr, bytes_per_rank * world), device="cuda")
        allgather_table(full, rank, world, group)
