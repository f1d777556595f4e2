"""Partition of one atmosphere-LUT build over the GPUs of a box.

Within a table every texel is independent; between tables the next kernel needs the whole previous
table (lookups are scattered), so each rank integrates a contiguous slab of (height, elevation) pairs
and one all-gather per table reassembles it.  The slab arithmetic lives in the library
(atmlut_slab); this module is its host-side mirror plus the collective used by the builder callback.
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
    """In-place all-gather of a padded table tensor `full` ([world * per_rank * texels_per_pair, 4] or flat):
    rank's slab is already in place; afterwards every rank holds every slab."""
    import torch.distributed as dist
    flat = full.view(-1)
    n = flat.numel() // world
    dist.all_gather_into_tensor(flat, flat[rank * n:(rank + 1) * n].clone() if flat.device.type == "cpu"
                                else flat[rank * n:(rank + 1) * n], group=group)
    return full
