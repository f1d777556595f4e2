"""Run under torchrun on N GPUs: the sharded build must produce exactly the bytes of a single-GPU build.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 \
        tools/multi_gpu_check.py [--configs reduced shipped stress] [--modes p2p nccl] [--report FILE]

For every (configuration, exchange mode): `--repeat` sharded builds in a row (buffer reuse across runs; the first
captures the CUDA graph, the others replay it), then the timed variant of the build; rank 0 downloads the four file
tables after each kind and compares them byte for byte with the tables of a single-GPU build made in the same process.
One line per check, `MULTI_GPU_CHECK OK` at the end if all were identical; rank 0 appends the lines to --report.
"""
import argparse
import hashlib
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from sfsim_b200 import _lib, atmosphere_lut  # noqa: E402

CONFIGS = {
    # pair count not divisible by the world size, fewer height rows than ranks at 8 GPUs
    "reduced": dict(ray_scatter_shape=(5, 7, 8, 2), transmittance_shape=(8, 15), surface_radiance_shape=(4, 7),
                    ray_steps=20, sphere_steps=8, iterations=2),
    "shipped": dict(),
    "stress": dict(ray_scatter_shape=(64, 253, 64, 16), iterations=10),
}


def digest(tables):
    h = hashlib.sha256()
    for t in tables:
        h.update(np.ascontiguousarray(t).tobytes())
    return h.hexdigest()[:16]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--configs", nargs="+", default=["reduced", "shipped"], choices=sorted(CONFIGS))
    ap.add_argument("--modes", nargs="+", default=["p2p", "nccl"], choices=["p2p", "nccl"])
    ap.add_argument("--reduced", action="store_true", help="same as --configs reduced")
    ap.add_argument("--mode", default=None, help="same as --modes MODE")
    ap.add_argument("--repeat", type=int, default=3, help="builds in a row (exercises buffer reuse across runs)")
    ap.add_argument("--report", default=None, help="append the result lines to this file")
    args = ap.parse_args()
    if args.reduced:
        args.configs = ["reduced"]
    if args.mode:
        args.modes = [args.mode]
    rank, local, world = int(os.environ["RANK"]), int(os.environ["LOCAL_RANK"]), int(os.environ["WORLD_SIZE"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    lib = _lib.load()
    _lib.check(lib.atmlut_init(local))
    lines = []
    ok = True

    def say(line):
        lines.append(line)
        print(line, flush=True)

    for name in args.configs:
        cfg = _lib.make_config(**CONFIGS[name])
        want = None
        if rank == 0:
            single = atmosphere_lut.AtmosphereLutBuilder(cfg=cfg)
            single.run()
            single.sync()
            want = single.download()
            single.close()
        dist.barrier()
        for mode in args.modes:
            sharded = atmosphere_lut.AtmosphereLutBuilder(cfg=cfg, rank=rank, world=world, mode=mode)
            for kind in ("graph x%d" % args.repeat if mode == "p2p" else "eager x%d" % args.repeat, "timed"):
                if kind == "timed":
                    sharded.run_timed()
                else:
                    for _ in range(args.repeat):
                        sharded.run()
                sharded.sync()
                dist.barrier()
                if rank == 0:
                    got = sharded.download()
                    same = all(np.array_equal(g, w) for g, w in zip(got, want))
                    ok = ok and same
                    detail = "" if same else "  " + "; ".join(
                        "%s max abs %.3g" % (n, float(np.abs(g - w).max()))
                        for n, g, w in zip(atmosphere_lut.FILE_NAMES, got, want) if not np.array_equal(g, w))
                    say("world=%d config=%-8s mode=%-4s %-9s sha256=%s single=%s %s%s" %
                        (world, name, mode, kind, digest(got), digest(want), "identical" if same else "DIFFERENT",
                         detail))
                dist.barrier()
            if rank == 0 and mode == "nccl":
                say("world=%d config=%-8s mode=nccl all-gathers per build: %d" %
                    (world, name, sharded.gathers // (args.repeat + 1)))
            sharded.close()
            dist.barrier()
    if rank == 0:
        say("MULTI_GPU_CHECK %s world=%d configs=%s modes=%s" % ("OK" if ok else "FAILED", world, ",".join(args.configs),
                                                                 ",".join(args.modes)))
        if args.report:
            with open(args.report, "a") as f:
                f.write("\n".join(lines) + "\n")
    flag = torch.tensor([1 if ok else 0], device="cuda")
    dist.broadcast(flag, 0)
    dist.destroy_process_group()
    sys.exit(0 if int(flag.item()) else 1)


if __name__ == "__main__":
    main()
