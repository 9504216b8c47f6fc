"""CPU suite: the N>1 host logic (sharding of independent worlds, all-gather of the replication snapshot) with two gloo ranks."""
import os
import socket
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_shard_range_partitions_exactly():
    from netphys_b200.dist import shard_range, guid_base
    for n in (0, 1, 7, 4096, 4097):
        for ws in (1, 2, 3, 8):
            spans = [shard_range(n, r, ws) for r in range(ws)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(spans[r][1] == spans[r + 1][0] for r in range(ws - 1))
            sizes = [hi - lo for lo, hi in spans]
            assert max(sizes) - min(sizes) <= 1
    assert guid_base(np.array([10, 12, 9]), 2) == 22


def _worker(rank, world_size, port, q):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    import torch
    import torch.distributed as dist
    from netphys_b200.dist import SnapshotExchange, shard_range
    from netphys_b200.lib import SNAPSHOT_DTYPE
    dist.init_process_group("gloo", rank=rank, world_size=world_size)
    try:
        worlds = 6; bodies = 5
        lo, hi = shard_range(worlds, rank, world_size)                      # 4096 worlds over 8 GPUs in miniature
        n = (hi - lo) * bodies
        x = SnapshotExchange(36 * n, torch.device("cpu"), world_size, rank)
        ok = True
        for step in range(3):
            snap = np.zeros(n, SNAPSHOT_DTYPE)
            snap["guid"] = lo * bodies + np.arange(n); snap["pos"][:, 0] = snap["guid"] + 0.5 * step; snap["rot"][:, 0] = 1.0; snap["is_enabled"] = 1
            buf = x.next_send_buffer(); buf.copy_(torch.from_numpy(snap.view(np.uint8).copy()))
            full = x.exchange(buf).numpy().view(SNAPSHOT_DTYPE)
            ok &= len(full) == worlds * bodies and np.array_equal(full["guid"], np.arange(worlds * bodies))
            ok &= np.array_equal(full["pos"][:, 0], np.arange(worlds * bodies) + 0.5 * step)
        q.put((rank, bool(ok)))
    finally:
        dist.destroy_process_group()


def test_snapshot_allgather_two_ranks_gloo():
    import torch.multiprocessing as mp
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    ctx = mp.get_context("spawn"); q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs: p.start()
    res = sorted(q.get(timeout=180) for _ in procs)
    for p in procs: p.join(timeout=60)
    assert res == [(0, True), (1, True)]


def test_padded_shard_covers_world_with_equal_slots():
    from netphys_b200.dist import padded_shard
    for n in (0, 1, 5, 4096, 4097, 1_000_003):
        for ws in (1, 2, 3, 8):
            parts = [padded_shard(n, r, ws) for r in range(ws)]
            assert len({p[2] for p in parts}) == 1 and parts[0][2] * ws >= n
            assert parts[0][0] == 0 and sum(p[1] for p in parts) == n
            assert all(parts[r][0] + parts[r][1] == parts[r + 1][0] for r in range(ws - 1))
            assert all(p[1] == 0 or p[0] == r * p[2] for r, p in enumerate(parts)), "a non-empty shard starts at its all-gather slot"


def _collision_worker(rank, world_size, port, q):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    import torch
    import torch.distributed as dist
    from netphys_b200.dist import CollisionExchange
    dist.init_process_group("gloo", rank=rank, world_size=world_size)
    try:
        n = 11                                                              # odd on purpose: the last slot is padded
        x = CollisionExchange(n, torch.device("cpu"), world_size, rank)
        ok = x.slot == 6 and x.first == 6 * rank and x.count == (6 if rank == 0 else 5)
        for step in range(3):
            i = torch.arange(x.first, x.first + x.count)
            x.hit_j[x.first:x.first + x.count] = (i + 1 + step).to(torch.int32)       # what the owned search would write
            x.flags[x.first:x.first + x.count] = 3
            x.contact[x.first:x.first + x.count] = (i[:, None] * 4 + torch.arange(4)[None, :]).to(torch.float32) + step
            x.exchange()
            a = torch.arange(n)
            ok &= bool(torch.equal(x.hit_j[:n], (a + 1 + step).to(torch.int32)))
            ok &= bool(torch.equal(x.contact[:n], (a[:, None] * 4 + torch.arange(4)[None, :]).to(torch.float32) + step))
        q.put((rank, bool(ok)))
    finally:
        dist.destroy_process_group()


def test_collision_record_allgather_two_ranks_gloo():
    import torch.multiprocessing as mp
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    ctx = mp.get_context("spawn"); q = ctx.Queue()
    procs = [ctx.Process(target=_collision_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs: p.start()
    res = sorted(q.get(timeout=180) for _ in procs)
    for p in procs: p.join(timeout=60)
    assert res == [(0, True), (1, True)]
