"""world_size-2 gloo test of the multi-GPU host logic: slab partition + one all-gather per table
reassembles exactly the table a single rank computes.  Runs on CPU."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from sfsim_b200 import sharding

SHAPE4 = (5, 7, 4, 2)      # 35 pairs: not divisible by 2 -> exercises the padding


def _texel_value(pair, texel):
    return np.array([pair * 1000.0 + texel, pair + 0.5, -texel, 0.0], dtype=np.float32)


def _full_table(world):
    n_pairs = SHAPE4[0] * SHAPE4[1]
    ntex = SHAPE4[2] * SHAPE4[3]
    padded = sharding.padded_pairs(n_pairs, world)
    tab = np.zeros((padded, ntex, 4), dtype=np.float32)
    for p in range(n_pairs):
        for t in range(ntex):
            tab[p, t] = _texel_value(p, t)
    return tab


def _worker(rank, world, port, results):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        n_pairs = SHAPE4[0] * SHAPE4[1]
        ntex = SHAPE4[2] * SHAPE4[3]
        begin, count, per_rank = sharding.slab(n_pairs, rank, world)
        table = torch.zeros(per_rank * world, ntex, 4)
        # "kernel": this rank fills only its slab, in place in the full-size buffer
        for p in range(begin, begin + count):
            for t in range(ntex):
                table[p, t] = torch.from_numpy(_texel_value(p, t))
        for _ in range(3):                               # one all-gather per table, several tables per build
            sharding.allgather_table(table, rank, world)
        results[rank] = table.numpy().copy()
    finally:
        dist.destroy_process_group()


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def test_two_rank_allgather_reassembles_the_table():
    world = 2
    manager = mp.Manager()
    results = manager.dict()
    mp.spawn(_worker, args=(world, _free_port(), results), nprocs=world, join=True)
    want = _full_table(world)
    for rank in range(world):
        np.testing.assert_array_equal(results[rank], want)
    # the padding rows stay zero and the real rows equal the single-rank table
    n_pairs = SHAPE4[0] * SHAPE4[1]
    np.testing.assert_array_equal(results[0][:n_pairs], _full_table(1)[:n_pairs])
    assert float(np.abs(results[0][n_pairs:]).max()) == 0.0
