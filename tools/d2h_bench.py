#!/usr/bin/env python
"""Platform ceiling of device-to-host copies: what a sweep that ships every PSF to the host can reach on this box.

One process per GPU (torchrun, or a single process for --gpus 1); each copies 32 MiB blocks (one 2048^2 fp64 PSF) from HBM
into its own pinned host buffers with plain ``cudaMemcpyAsync`` on `--streams` streams for `--seconds`, after binding to
the GPU's NUMA node like ``bench.py`` does.  Prints per-rank and aggregate GB/s, i.e. PSF/s = GB/s / 0.0336.

    python tools/d2h_bench.py                      # one GPU
    python -m torch.distributed.run --nproc-per-node 8 --master-addr 127.0.0.1 tools/d2h_bench.py
"""
import argparse
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--seconds", type=float, default=3.0)
    ap.add_argument("--streams", type=int, default=4)
    ap.add_argument("--mib", type=int, default=32)
    ap.add_argument("--no-numa", action="store_true")
    args = ap.parse_args()
    import torch
    import torch.distributed as dist

    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("gloo")
    numa = None
    if not args.no_numa and world > 1:
        from paos_b200.sweep import bind_to_gpu_numa

        numa = bind_to_gpu_numa(local)
    n = args.mib << 20
    dev = torch.empty(n, dtype=torch.uint8, device="cuda")
    hosts = [torch.empty(n, dtype=torch.uint8, pin_memory=True) for _ in range(args.streams)]
    streams = [torch.cuda.Stream() for _ in range(args.streams)]
    for h, s in zip(hosts, streams):  # warm-up
        with torch.cuda.stream(s):
            h.copy_(dev, non_blocking=True)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    t0 = time.perf_counter()
    copies = 0
    while time.perf_counter() - t0 < args.seconds:
        for h, s in zip(hosts, streams):
            with torch.cuda.stream(s):
                h.copy_(dev, non_blocking=True)
            copies += 1
        for s in streams:
            s.synchronize()
    dt = time.perf_counter() - t0
    gbs = copies * n / dt / 1e9
    if world > 1:
        allv = [None] * world
        dist.all_gather_object(allv, (gbs, numa))
    else:
        allv = [(gbs, numa)]
    if rank == 0:
        print(json.dumps({"gpus": world, "block_MiB": args.mib, "streams": args.streams, "per_rank_GBs": [round(v[0], 2) for v in allv],
                          "numa_nodes": [v[1] for v in allv], "aggregate_GBs": round(sum(v[0] for v in allv), 2),
                          "aggregate_psf_per_s_2048_fp64": round(sum(v[0] for v in allv) * 1e9 / (2048 * 2048 * 8), 1)}))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
