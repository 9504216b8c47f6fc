"""Multi-GPU host logic: independent worlds are sharded over ranks (no data-path collective); the only exchange is
the all-gather of the per-step replication snapshot (36-byte CommandFrameObject records, netphys_common/common.h:120-127)
that the networking layer sends to clients.  torch.distributed is the plumbing (nccl on GPUs, gloo in the CPU tests)."""
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
