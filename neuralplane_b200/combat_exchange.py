"""The one real exchange step of the path: relative geometry for combat when a pair's two aircraft do NOT
live on the same rank (role-sharded layout: ego population on some ranks, opponents on others -- natural when the two
policies run on different GPUs).  Pair-sharded layouts (the default; `shard_range` never splits a pair) need none of
this: the step kernel holds both aircraft of a pair in one thread.

    records = local_records(env)                         # [n_local, 8] written by combat_records_kernel
    allrec  = gather_records(records, group)             # NCCL all-gather over NVLink: [world * n_local, 8]
    geo     = relative_geometry(allrec, ego_idx, enm_idx) # [m, 8]: AO TA R AO2 TA2 R2 side dvx

Record = position (3), inertial velocity xdot[0:3] (3), body-axis vx, blood: 32 B per aircraft, so 10^6 aircraft gather
32 MB per rank per step -- tens of microseconds on NVLink 5, negligible next to the step kernel.

`PeerRecordExchange` is the fused form of the same exchange: the record slabs live in peer-mapped (symmetric) memory and
np_combat_relgeo_peers pulls the partner records over NVLink with its own loads -- no gathered array, no NCCL call, only
the 32 B per pair this rank needs cross the link, and the transfer overlaps the geometry arithmetic.
"""
import torch

from . import _native as nv
from .sharding import gather_rows

RECORD_WIDTH = 8
GEO_COLUMNS = ("AO", "TA", "R", "AO2", "TA2", "R2", "side", "dvx")


def local_records(env, out=None):
    """[n, 8] float32 records of this rank's aircraft (device tensor; reuses `out` as the send slab if given)."""
    if out is None:
        out = torch.empty((env.n, RECORD_WIDTH), dtype=torch.float32, device=env.device)
    nv.check(nv.lib().np_env_combat_records(env._handle, out.data_ptr(), env._stream()), "np_env_combat_records")
    return out


def gather_records(records, group=None):
    """All-gather the per-rank slabs in rank order: [world * n_local, 8] (identity without a process group)."""
    return gather_rows(records, group)


def relative_geometry(all_records, ego_idx, enm_idx):
    """Pairwise terms for index pairs into the gathered records: [m, 8] = AO, TA, R, AO2, TA2, R2, side, dvx."""
    ego_idx = ego_idx.to(device=all_records.device, dtype=torch.int32).contiguous()
    enm_idx = enm_idx.to(device=all_records.device, dtype=torch.int32).contiguous()
    m = ego_idx.numel()
    out = torch.empty((m, 8), dtype=torch.float32, device=all_records.device)
    st = nv.lib().np_combat_relgeo(all_records.contiguous().data_ptr(), ego_idx.data_ptr(), enm_idx.data_ptr(), out.data_ptr(), m,
                                   torch.cuda.current_stream(all_records.device).cuda_stream)
    nv.check(st, "np_combat_relgeo")
    return out


def role_sharded_partner_index(n_local_envs, rank, world):
    """Role-sharded layout: ranks [0, world/2) hold the egos of env block r, ranks [world/2, world) the opponents of env
    block r - world/2.  Returns (ego_idx, enm_idx) into the rank-ordered gathered array for THIS rank's env block."""
    if world % 2:
        raise ValueError("role sharding needs an even world size")
    half = world // 2
    block = rank % half
    e = torch.arange(n_local_envs, dtype=torch.int64)
    return block * n_local_envs + e, (half + block) * n_local_envs + e


class PeerRecordExchange:
    """Role-sharded exchange over NVLink peer memory.  One symmetric [n_local, 8] record slab per rank:

        ex  = PeerRecordExchange(n_local, device)            # collective: allocates + rendezvous
        geo = ex.geometry(records_or_env, ego_idx, enm_idx)  # records -> slab, barrier, fused pull + geometry, barrier

    Needs an initialised NCCL process group and P2P access between the ranks' GPUs (one NVSwitch domain)."""

    def __init__(self, n_local, device, group=None):
        import torch.distributed as dist
        import torch.distributed._symmetric_memory as symm
        self.group = group if group is not None else dist.group.WORLD
        self.n_local = int(n_local)
        self.slab = symm.empty((self.n_local, RECORD_WIDTH), dtype=torch.float32, device=device)
        self.handle = symm.rendezvous(self.slab, self.group)
        self.world = self.handle.world_size

    def geometry(self, records, ego_idx, enm_idx):
        """records: this rank's [n_local, 8] records (copied into the slab), a combat env (its records kernel writes the
        slab directly) or None (the caller already wrote `self.slab`).  ego_idx / enm_idx index the virtual rank-ordered gathered array.  Returns [m, 8]."""
        dev = self.slab.device
        if torch.is_tensor(records):
            self.slab.copy_(records)
        elif records is not None:                       # None: the slab already holds this step's records
            local_records(records, out=self.slab)
        self.handle.barrier(channel=0)                  # every rank's slab is written and visible to its peers
        ego_idx = ego_idx.to(device=dev, dtype=torch.int32).contiguous()
        enm_idx = enm_idx.to(device=dev, dtype=torch.int32).contiguous()
        m = ego_idx.numel()
        out = torch.empty((m, 8), dtype=torch.float32, device=dev)
        st = nv.lib().np_combat_relgeo_peers(self.handle.buffer_ptrs_dev, self.world, self.n_local, ego_idx.data_ptr(),
                                             enm_idx.data_ptr(), out.data_ptr(), m, torch.cuda.current_stream(dev).cuda_stream)
        nv.check(st, "np_combat_relgeo_peers")
        self.handle.barrier(channel=1)                  # peers have read this slab: it may be overwritten by the next step
        return out


# ---------------------------------------------------------------------------------------------------------------------------
# Role-sharded combat STEP (SingleCombatEnv(layout='role')): the objects that make the partner rank's 28-float records
# readable between the local half and the pair half of a step (include/nplane.h: np_env_combat_role_local / _pair).
# ---------------------------------------------------------------------------------------------------------------------------
RECORD_FLOATS = nv.COMBAT_RECORD_FLOATS


def role_block(num_envs_total, rank, world):
    """Role-sharded layout over an even world: ranks [0, world/2) hold the egos of env block b = rank, ranks
    [world/2, world) the opponents of block b = rank - world/2.  Returns (role, first_env, n_envs) of this rank's block;
    blocks are contiguous env ranges of even size (the step kernel moves aircraft in pairs)."""
    if world < 2 or world % 2:
        raise ValueError("role sharding needs an even world size >= 2")
    half = world // 2
    role, block = rank // half, rank % half
    from .sharding import shard_range
    first, n = shard_range(num_envs_total, block, half)
    if n % 2:
        raise ValueError("role sharding needs an even number of envs per block")
    return role, first, n


def partner_rank(rank, world):
    return (rank + world // 2) % world


class _Timed:
    def time_exchange(self, env, iters):
        """Milliseconds per exchange (barrier / all-gather alone), device-timed on the env's stream."""
        torch.cuda.synchronize(env.device)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(iters):
            self.sync(env)
        e1.record()
        torch.cuda.synchronize(env.device)
        return e0.elapsed_time(e1) / iters


class LocalPairExchange(_Timed):
    """Both role envs live in ONE process on one device (tests, single-GPU debugging): the 'partner slab' is simply the
    other env's slab, and stream order is the barrier.  Drive the two envs with `step_both` / `reset_both`."""
    peer = False

    def __init__(self, env_ego, env_enm):
        assert env_ego.role == 0 and env_enm.role == 1 and env_ego.n == env_enm.n and env_ego.device == env_enm.device
        self.slabs = [torch.zeros((e.n, RECORD_FLOATS), dtype=torch.float32, device=e.device) for e in (env_ego, env_enm)]
        self.envs = (env_ego, env_enm)
        env_ego.connect(self)
        env_enm.connect(self)

    def own_slab(self, env):
        return self.slabs[env.role]

    def partner_ptr(self, env):
        return self.slabs[1 - env.role].data_ptr()

    def sync(self, env):
        pass

    def advance(self, env):
        pass

    def link_bytes(self, env):
        return 0

    def reset_both(self, draws_ego=None, draws_enm=None):
        e0, e1 = self.envs
        for e in (e0, e1):
            e._flags.fill_(1)
            e._pair_reset.fill_(1)
        e0.step_local(None, 0, draws_ego)
        e1.step_local(None, 0, draws_enm)
        e0.step_pair(0)
        e1.step_pair(0)
        return e0.last_obs, e1.last_obs

    def step_both(self, a_ego, a_enm, draws_ego=None, draws_enm=None):
        e0, e1 = self.envs
        e0.step_local(e0._normalise_action(a_ego), e0.n_substeps, draws_ego)
        e1.step_local(e1._normalise_action(a_enm), e1.n_substeps, draws_enm)
        e0.step_pair(e0.n_substeps)
        e1.step_pair(e1.n_substeps)
        return [(e.last_obs, e.last_reward, e.is_done, e.bad_done, e.exceed_time_limit) for e in (e0, e1)]


class PeerSlabExchange(_Timed):
    """Record slabs in NVLink peer-mapped (symmetric) memory, double-buffered: the pair kernel pulls the partner's records
    with its own loads -- only the 112 B per env this rank needs cross the link, no gathered array, no NCCL call; ONE
    cross-device barrier per step (the slab written at step k is next overwritten at step k + 2, after barrier k + 1)."""
    peer = True

    def __init__(self, env, group=None):
        import torch.distributed as dist
        import torch.distributed._symmetric_memory as symm
        self.group = group if group is not None else dist.group.WORLD
        self.rank, self.world = dist.get_rank(self.group), dist.get_world_size(self.group)
        self.buf = symm.empty((2, env.n, RECORD_FLOATS), dtype=torch.float32, device=env.device)
        self.buf.zero_()
        self.handle = symm.rendezvous(self.buf, self.group)
        self.partner = partner_rank(self.rank, self.world)
        self.partner_base = int(self.handle.buffer_ptrs[self.partner])
        self.k = 0
        torch.cuda.synchronize(env.device)
        self.handle.barrier(channel=0)

    def own_slab(self, env):
        return self.buf[self.k]

    def partner_ptr(self, env):
        return self.partner_base + self.k * env.n * RECORD_FLOATS * 4

    def sync(self, env):
        self.handle.barrier(channel=0)

    def advance(self, env):
        self.k ^= 1

    def link_bytes(self, env):
        return env.n * RECORD_FLOATS * 4


class AllGatherExchange(_Timed):
    """The exchange BASELINE configs[4] names: an NCCL all-gather of every rank's record slab; the pair kernel reads the
    partner block of the gathered array.  Moves world x n x 112 B to every rank where the peer-slab form moves n x 112 B."""
    peer = False

    def __init__(self, env, group=None):
        import torch.distributed as dist
        self.dist = dist
        self.group = group if group is not None else dist.group.WORLD
        self.rank, self.world = dist.get_rank(self.group), dist.get_world_size(self.group)
        self.slab = torch.zeros((env.n, RECORD_FLOATS), dtype=torch.float32, device=env.device)
        self.gathered = torch.zeros((self.world * env.n, RECORD_FLOATS), dtype=torch.float32, device=env.device)
        self.partner = partner_rank(self.rank, self.world)

    def own_slab(self, env):
        return self.slab

    def partner_ptr(self, env):
        return self.gathered[self.partner * env.n:].data_ptr()

    def sync(self, env):
        self.dist.all_gather_into_tensor(self.gathered, self.slab, group=self.group)

    def advance(self, env):
        pass

    def link_bytes(self, env):
        return (self.world - 1) * env.n * RECORD_FLOATS * 4
