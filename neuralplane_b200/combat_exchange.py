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
