"""Multi-GPU host logic: independent worlds are sharded over ranks (no data-path collective); the only exchange is
the all-gather of the per-step replication snapshot (36-byte CommandFrameObject records, netphys_common/common.h:120-127)
that the networking layer sends to clients.  torch.distributed is the plumbing (nccl on GPUs, gloo in the CPU tests); on GPUs the
exchange itself is the library's (init_world_comm + World.set_snapshot_gather: every np_world_step pushes its slot into the peers'
buffers with the copy engines -- peers mapped through CUDA IPC, ncclAllGather as the fall-back -- see DESIGN.md section 6)."""
import numpy as np


def shard_range(n_items, rank, world_size):
    """Contiguous block partition: rank r owns items [lo, hi).  Sizes differ by at most one."""
    base, rem = divmod(n_items, world_size)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def guid_base(bodies_per_rank, rank):
    """Snapshot guids are global body indices: rank r's bodies follow those of ranks < r."""
    return int(np.sum(bodies_per_rank[:rank]))


class SnapshotExchange:
    """All-gather of equally sized per-rank snapshot buffers into one global snapshot on every rank.

    send / recv are torch uint8 tensors (CUDA for nccl, CPU for gloo).  With world_size == 1 it is a copy-free no-op.
    """

    def __init__(self, bytes_per_rank, device, world_size, rank):
        import torch
        self.torch = torch
        self.world_size, self.rank = world_size, rank
        self.send = [torch.zeros(bytes_per_rank, dtype=torch.uint8, device=device) for _ in range(2)]   # double buffered
        self.recv = torch.zeros(bytes_per_rank * world_size, dtype=torch.uint8, device=device) if world_size > 1 else None
        self.flip = 0

    def next_send_buffer(self):
        self.flip ^= 1
        return self.send[self.flip]

    def exchange(self, buf):
        if self.world_size == 1:
            return buf
        self.torch.distributed.all_gather_into_tensor(self.recv, buf)
        return self.recv


def padded_shard(n_bodies, rank, world_size):
    """Equal-length index ranges for the sharded pair search of ONE world: rank r searches bodies [first, first + count) where
    every rank's slot is `slot` = ceil(n / world_size) long (the all-gather needs equal pieces); the last ranks may own fewer
    (or no) bodies.  Returns (first, count, slot)."""
    slot = max(1, -(-n_bodies // world_size))
    first = min(n_bodies, rank * slot)
    return first, min(n_bodies, first + slot) - first, slot


class CollisionExchange:
    """The one exchange of a sharded world step (np_world_set_shard / step_begin / step_end): all-gather, in place, of the per-body
    collision records -- hit_j int32 (partner index, CollisionData::success in bit 30) and contact float32[4] (penetration direction,
    depth): physics.h:26-31 Collision, 20 bytes per body -- each rank contributing the slot of bodies it searched.  `flags` is scratch
    of the searching rank and does not travel.  Tensors are CUDA for nccl, CPU for gloo."""

    def __init__(self, n_bodies, device, world_size, rank):
        import torch
        self.torch = torch
        self.world_size, self.rank = world_size, rank
        self.first, self.count, self.slot = padded_shard(n_bodies, rank, world_size)
        total = self.slot * world_size
        self.hit_j = torch.full((total,), -1, dtype=torch.int32, device=device)
        self.flags = torch.zeros(total, dtype=torch.int32, device=device)
        self.contact = torch.zeros((total, 4), dtype=torch.float32, device=device)

    def pointers(self):
        return self.hit_j.data_ptr(), self.flags.data_ptr(), self.contact.data_ptr()

    def exchange(self):
        if self.world_size == 1:
            return
        lo, hi = self.rank * self.slot, (self.rank + 1) * self.slot
        for t in (self.hit_j, self.contact):
            piece = t[lo:hi]
            if not t.is_cuda:
                piece = piece.clone()                   # gloo: no aliasing of input and output
            self.torch.distributed.all_gather_into_tensor(t, piece)


def init_world_comm(world, rank, world_size):
    """Give `world` its own NCCL communicator (np_comm_init, the C-ABI path a C++ server uses): rank 0 creates the 128-byte unique id,
    torch.distributed only carries it to the other ranks."""
    import torch.distributed as dist
    box = [world.comm_unique_id() if rank == 0 else None]
    dist.broadcast_object_list(box, src=0)
    world.comm_init(world_size, rank, box[0])


class ShardedWorld:
    """One world whose pair search is split over the ranks (BASELINE configs[4]).  Every rank loads the same scene.
    lib_comm=True (default on GPUs): the library's own NCCL path -- np_world_shard_over_comm + np_world_step_sharded = step_begin ->
    one grouped ncclAllGather of hit_j and contact on the world's stream -> step_end.  lib_comm=False: the same exchange through
    torch.distributed (CollisionExchange), kept for the gloo tests and as a cross-check."""

    def __init__(self, world, device, world_size, rank, lib_comm=True, halo=0):
        self.w = world; self.lib_comm = lib_comm
        if lib_comm:
            init_world_comm(world, rank, world_size)
            world.shard_over_comm()
            if halo > 0:
                world.set_shard_halo(halo)      # partitioned steps: own bodies only, ncclSend/Recv of the halos between neighbours instead of the all-gather
            return
        assert halo == 0, "partitioned steps go through the library's own NCCL path"
        import torch
        self.torch = torch
        self.x = CollisionExchange(world.n, device, world_size, rank)
        world.set_shard(self.x.first, self.x.count, *self.x.pointers())
        self.stream = torch.cuda.ExternalStream(world.L.np_world_stream(world.w), device=device)

    def step(self, dt):
        if self.lib_comm:
            self.w.step_sharded(dt); return
        self.w.step_begin(dt)
        with self.torch.cuda.stream(self.stream):       # the collective is ordered after the search and before the response
            self.x.exchange()
        self.w.step_end()
