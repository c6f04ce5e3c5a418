#!/usr/bin/env python
"""What the box can carry back to the host: N processes (torchrun), each copying the e2e step's read-back
(125 MB: 2^20 envs x (120 B observations + 4 B reward + 1 B done)) from its GPU to pinned host memory, all
at once.  The figure bench.py's `e2e` is compared with ("nobody knows whether 7.3e8 is 25 % or
95 % of achievable").

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29511 \
        tools/microbench/d2h_ceiling.py [--bind] [--h2d]
"""
import argparse
import json
import os
import sys
import time

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from manytor_b200 import distributed as mtd  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--bind", action="store_true", help="pin each rank to the cores next to its GPU first")
    ap.add_argument("--h2d", action="store_true", help="also run the 16 MB action upload concurrently (full duplex)")
    ap.add_argument("--iters", type=int, default=20)
    args = ap.parse_args()
    rank, world, local = mtd.init_from_env()
    torch.cuda.set_device(local)
    cores = mtd.bind_to_gpu_numa(local) if args.bind else None
    n = 1 << 20
    d2h_bytes, h2d_bytes = n * 125, n * 16
    dbuf = torch.empty(d2h_bytes, dtype=torch.uint8, device="cuda")
    hbuf = torch.empty(d2h_bytes, dtype=torch.uint8).pin_memory()
    ha = torch.empty(h2d_bytes, dtype=torch.uint8).pin_memory()
    da = torch.empty(h2d_bytes, dtype=torch.uint8, device="cuda")
    side = torch.cuda.Stream()

    def once():
        hbuf.copy_(dbuf, non_blocking=True)
        if args.h2d:
            with torch.cuda.stream(side):
                da.copy_(ha, non_blocking=True)

    for _ in range(3):
        once()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    t0 = time.perf_counter()
    for _ in range(args.iters):
        once()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    dt = mtd.max_over_ranks(time.perf_counter() - t0, torch.device("cuda", local)) / args.iters
    if rank == 0:
        print(json.dumps({"gpus": world, "bind": bool(cores), "h2d_concurrent": args.h2d,
                          "d2h_gbs_per_gpu": d2h_bytes / dt / 1e9, "d2h_gbs_total": world * d2h_bytes / dt / 1e9,
                          "env_steps_per_s_ceiling": world * n / dt, "ms_per_step": dt * 1e3}))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
