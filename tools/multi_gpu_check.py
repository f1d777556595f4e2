"""Run under torchrun on N GPUs: the sharded build (slabs + one NCCL all-gather per table) must produce
exactly the bytes of a single-GPU build.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 \
        tools/multi_gpu_check.py [--reduced]
"""
import argparse
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from sfsim_b200 import _lib, atmosphere_lut  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--reduced", action="store_true", help="4-D [5,7,8,2]: pair count not divisible by the world size")
    ap.add_argument("--mode", default="p2p", choices=["p2p", "nccl"])
    ap.add_argument("--repeat", type=int, default=3, help="builds in a row (exercises buffer reuse across runs)")
    args = ap.parse_args()
    rank, local, world = int(os.environ["RANK"]), int(os.environ["LOCAL_RANK"]), int(os.environ["WORLD_SIZE"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    lib = _lib.load()
    _lib.check(lib.atmlut_init(local))
    if args.reduced:
        cfg = _lib.make_config(ray_scatter_shape=(5, 7, 8, 2), transmittance_shape=(8, 15), surface_radiance_shape=(4, 7),
                               ray_steps=20, sphere_steps=8, iterations=2)
    else:
        cfg = _lib.default_config()
    sharded = atmosphere_lut.AtmosphereLutBuilder(cfg=cfg, rank=rank, world=world, mode=args.mode)
    for _ in range(args.repeat):
        sharded.run()
    sharded.sync()
    got = sharded.download()
    gathers = sharded.gathers
    sharded.close()
    dist.barrier()
    ok = True
    if rank == 0:
        single = atmosphere_lut.AtmosphereLutBuilder(cfg=cfg)
        single.run()
        single.sync()
        want = single.download()
        single.close()
        for name, g, w in zip(atmosphere_lut.FILE_NAMES, got, want):
            same = np.array_equal(g, w)
            ok = ok and same
            print("%-26s %s" % (name, "identical" if same else "DIFFERENT (max abs %.3g)" % float(np.abs(g - w).max())))
        print("all-gathers per build: %d" % gathers)
        print("MULTI_GPU_CHECK %s world=%d mode=%s" % ("OK" if ok else "FAILED", world, args.mode))
    flag = torch.tensor([1 if ok else 0], device="cuda")
    dist.broadcast(flag, 0)
    dist.destroy_process_group()
    sys.exit(0 if int(flag.item()) else 1)


if __name__ == "__main__":
    main()
