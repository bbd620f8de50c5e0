#!/usr/bin/env python
"""Role-sharded combat STEP across ranks (torchrun, an even number of ranks, one GPU each):

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 tools/combat_role_check.py

Ranks [0, world/2) hold the egos of env block b, ranks [world/2, world) its opponents (SingleCombatEnv(layout='role')).
Every rank ALSO steps the same env block pair-sharded (both aircraft local) with the same seed and actions; every output
of the role-sharded env -- obs, reward, flags, state, controls, blood, step counts -- must equal the pair-sharded env's
rows of this rank's role bit for bit, for both exchanges (NVLink peer slabs pulled by the pair kernel; NCCL all-gather),
with Crash / Shutdown / env-level resets occurring.  Prints one JSON line (rank 0)."""
import argparse
import json
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from neuralplane_b200 import SingleCombatEnv  # noqa: E402
from neuralplane_b200.combat_exchange import AllGatherExchange, PeerSlabExchange, role_block  # noqa: E402


def run(ex_cls, dev, rank, world, envs_total, steps):
    role, first, E = role_block(envs_total, rank, world)
    pair = SingleCombatEnv(num_envs=E, config="selfplay", random_seed=5, device=dev, index_base=2 * first)
    rolee = SingleCombatEnv(num_envs=E, config="selfplay", random_seed=5, device=dev, layout="role", role=role, first_env=first)
    ex = ex_cls(rolee)
    rolee.connect(ex)
    same = bool(torch.equal(pair.reset()[role::2], rolee.reset()))

    def force_events():   # identical on both ranks of a block: a third of the pairs inside crash / gun range, a few nearly dead
        close = torch.arange(E, device=dev) % 3 == 0
        gap = torch.linspace(30.0, 9000.0, E, device=dev)
        s = pair.model.s
        s[1::2][close, 0] = s[0::2][close, 0] + gap[close]
        s[1::2][close, 1] = s[0::2][close, 1] + 0.03 * gap[close]
        s[1::2][close, 2] = s[0::2][close, 2] + 15.0
        s[0::2][close, 5] = 0.0
        s[1::2][close, 5] = 0.0
        pair.blood[1::18] = 0.4
        pair.blood[6::54] = 0.3
        rolee.model.s.copy_(pair.model.s[role::2])
        rolee.blood.copy_(pair.blood[role::2])
    force_events()
    g = torch.Generator(device=dev).manual_seed(3 + rank % (world // 2))      # the same tape on both ranks of a block
    events = 0
    for k in range(steps):
        a = torch.rand((2 * E, 4), device=dev, generator=g) - 0.5
        rp, rr = pair.step(a), rolee.step(a[role::2].contiguous())
        same = same and all(bool(torch.equal(x[role::2], y)) for x, y in zip(rp[:5], rr[:5]))
        same = same and bool(torch.equal(pair.model.s[role::2], rolee.model.s)) and bool(torch.equal(pair.model.u[role::2], rolee.model.u))
        same = same and bool(torch.equal(pair.blood[role::2], rolee.blood)) and bool(torch.equal(pair.step_count[role::2], rolee.step_count))
        events += int(rp[2].sum()) + int(rp[3].sum())
        if k == steps // 2:
            force_events()
    xms = ex.time_exchange(rolee, 20)
    t = torch.tensor([1 if same else 0, events], device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MIN)
    return {"bit_identical": bool(t[0].item()), "exchange_ms": xms, "link_bytes_per_step_per_rank": ex.link_bytes(rolee)}, int(t[1].item())


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--envs", type=int, default=100_000, help="envs in total")
    ap.add_argument("--steps", type=int, default=20)
    args = ap.parse_args()
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if rank:
        os.dup2(2, 1)
    dist.init_process_group("nccl", device_id=dev)
    out = {"world": world, "envs_total": args.envs, "steps": args.steps}
    out["peer_slabs"], ev = run(PeerSlabExchange, dev, rank, world, args.envs, args.steps)
    out["all_gather"], _ = run(AllGatherExchange, dev, rank, world, args.envs, args.steps)
    out["events"] = ev
    if rank == 0:
        print(json.dumps(out), flush=True)
    dist.destroy_process_group()
    sys.exit(0 if out["peer_slabs"]["bit_identical"] and out["all_gather"]["bit_identical"] else 1)


if __name__ == "__main__":
    main()
