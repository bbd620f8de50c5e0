#!/usr/bin/env python
"""Platform ceiling of the numpy boundary: copy-only (no kernels), what one env step moves through PCIe per rank --
95 MB device->host (obs 88 + reward 4 + flags 3 B per aircraft at n = 10^6) and 16 MB host->device (actions) -- with
N ranks doing it concurrently.  Run alone (N = 1) or under torchrun (N = 2, 4, 8).  Prints one JSON line (rank 0):
aggregate GB/s and the implied floor of an e2e step, for D2H alone, H2D alone and both directions at once."""
import json
import os
import sys
import time

import torch
import torch.distributed as dist


def main():
    rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    n = 1_000_000
    d2h_bytes, h2d_bytes = n * 95, n * 16
    dsrc = torch.empty(d2h_bytes, dtype=torch.uint8, device=dev)
    hdst = torch.empty(d2h_bytes, dtype=torch.uint8).pin_memory()
    hsrc = torch.empty(h2d_bytes, dtype=torch.uint8).pin_memory()
    ddst = torch.empty(h2d_bytes, dtype=torch.uint8, device=dev)
    s_up, s_dn = torch.cuda.Stream(dev), torch.cuda.Stream(dev)

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    def run(up, down, iters=30):
        for _ in range(3):
            if down:
                with torch.cuda.stream(s_dn):
                    hdst.copy_(dsrc, non_blocking=True)
            if up:
                with torch.cuda.stream(s_up):
                    ddst.copy_(hsrc, non_blocking=True)
        barrier()
        t0 = time.perf_counter()
        for _ in range(iters):
            if down:
                with torch.cuda.stream(s_dn):
                    hdst.copy_(dsrc, non_blocking=True)
            if up:
                with torch.cuda.stream(s_up):
                    ddst.copy_(hsrc, non_blocking=True)
        torch.cuda.synchronize()
        el = time.perf_counter() - t0
        t = torch.tensor([el], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item()) / iters

    out = {"n_gpus": world, "aircraft_per_gpu": n, "d2h_bytes_per_step": d2h_bytes, "h2d_bytes_per_step": h2d_bytes}
    for name, up, down in (("d2h_only", False, True), ("h2d_only", True, False), ("both", True, True)):
        s = run(up, down)
        moved = (d2h_bytes if down else 0) + (h2d_bytes if up else 0)
        out[name] = {"ms_per_step": 1e3 * s, "aggregate_GBps": world * moved / s / 1e9,
                     "aircraft_steps_per_s_ceiling": world * n / s}
    if rank == 0:
        print(json.dumps(out), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    sys.exit(main())
