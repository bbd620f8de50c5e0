"""Host-side N>1 logic on CPU: world_size-2 gloo process group (the GPU path uses the same functions over nccl)."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from neuralplane_b200.sharding import gather_rows, max_over_ranks, reduce_counters, shard_range


@pytest.mark.parametrize("n,world", [(1_000_000, 1), (1_000_000, 2), (1_000_000, 8), (1_000_001, 8), (7, 4), (1, 2), (0, 3)])
def test_shard_ranges_partition_the_population(n, world):
    spans = [shard_range(n, r, world) for r in range(world)]
    pos = 0
    for r, (lo, cnt) in enumerate(spans):
        assert lo == pos or cnt == 0
        assert lo % 2 == 0                      # pairs are never split across ranks
        pos = lo + cnt if cnt else pos
    assert pos == n
    sizes = [c for _, c in spans]
    assert max(sizes) - min(sizes) <= 2


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, n_total, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        lo, cnt = shard_range(n_total, rank, world)
        counters = {"overload": 10 * (rank + 1), "resets": cnt, "reached": rank}
        red = reduce_counters(counters)
        ms = max_over_ranks(1.0 + rank)
        rows = torch.arange(lo, lo + cnt, dtype=torch.float32).reshape(-1, 1).repeat(1, 8)   # 8-float records
        allrows = gather_rows(rows)
        q.put((rank, lo, cnt, red, ms, allrows[:, 0].numpy().copy()))
    finally:
        dist.destroy_process_group()


def test_world_size_2_collectives_over_gloo():
    world, n_total = 2, 4096
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, n_total, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted((q.get(timeout=120) for _ in procs), key=lambda x: x[0])
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for rank, lo, cnt, red, ms, col in res:
        assert red == {"overload": 30, "resets": n_total, "reached": 1}
        assert ms == 2.0
        assert np.array_equal(col, np.arange(n_total, dtype=np.float32))      # rank-ordered gather = global order
    assert res[0][1] == 0 and res[0][2] + res[1][2] == n_total


@pytest.mark.parametrize("world", [2, 8])
def test_role_sharded_partner_index(world):
    from neuralplane_b200.combat_exchange import role_sharded_partner_index
    n_envs = 5
    seen = set()
    for r in range(world):
        ego, enm = role_sharded_partner_index(n_envs, r, world)
        assert ego.shape == enm.shape == (n_envs,)
        assert int(ego.max()) < (world // 2) * n_envs <= int(enm.min())       # egos in the first half of the gather
        assert torch.equal(enm - ego, torch.full_like(ego, (world // 2) * n_envs))
        seen.add((int(ego[0]), int(enm[0])))
    assert len(seen) == world // 2        # rank r and rank r + world/2 look at the same pairs


@pytest.mark.parametrize("world,envs", [(2, 500_000), (4, 500_000), (8, 500_000), (8, 100), (2, 6)])
def test_role_blocks_cover_every_env_once_per_role(world, envs):
    """SingleCombatEnv(layout='role'): ranks [0, world/2) hold the egos of contiguous env blocks, rank r + world/2 the
    opponents of the SAME block; global aircraft index of local aircraft i = 2 (first_env + i) + role."""
    from neuralplane_b200.combat_exchange import partner_rank, role_block
    cover = {0: [], 1: []}
    for r in range(world):
        role, first, n = role_block(envs, r, world)
        assert role == (0 if r < world // 2 else 1) and n % 2 == 0
        assert role_block(envs, partner_rank(r, world), world)[1:] == (first, n)       # the partner rank holds the same envs
        assert partner_rank(partner_rank(r, world), world) == r
        cover[role] += list(range(first, first + n))
    assert cover[0] == list(range(envs)) and cover[1] == list(range(envs))


def _role_worker(rank, world, port, envs, q):
    """The all-gather form of the role-sharded exchange over gloo: every rank contributes its [n, 28] record slab; the block
    the pair kernel reads (gathered[partner * n:]) must be the partner rank's slab, i.e. the other aircraft of MY envs."""
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from neuralplane_b200.combat_exchange import RECORD_FLOATS, partner_rank, role_block
        role, first, n = role_block(envs, rank, world)
        gidx = 2 * (first + torch.arange(n)) + role                       # global aircraft index = index_base + 2 i
        slab = gidx.to(torch.float32).reshape(-1, 1).repeat(1, RECORD_FLOATS)
        gathered = torch.empty((world * n, RECORD_FLOATS))
        dist.all_gather_into_tensor(gathered, slab)
        part = gathered[partner_rank(rank, world) * n:][:n, 0].to(torch.int64)
        q.put((rank, bool(torch.equal(part, gidx + (1 - 2 * role)))))       # ego 2e <-> opponent 2e + 1
    finally:
        dist.destroy_process_group()


def test_role_sharded_all_gather_addresses_the_partner_over_gloo():
    world, envs = 2, 4096
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_role_worker, args=(r, world, port, envs, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in procs)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert res == [(0, True), (1, True)]
