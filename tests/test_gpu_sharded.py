"""GPU suite: ONE world whose pair search is sharded over R processes (np_world_set_shard / step_begin / step_end,
SURVEY 8(e)-2).  R worlds on one GPU stand in for the R ranks and the host plays the all-gather; the NCCL version of the same
exchange is netphys_b200.dist.ShardedWorld (exercised by bench.py under torchrun, and by tests/test_dist.py on gloo)."""
import numpy as np
import pytest

from netphys_b200 import scenes
from netphys_b200.dist import padded_shard
from conftest import assert_biteq

pytestmark = pytest.mark.gpu


class _Rank:
    def __init__(self, npb, sc, n_ranks, rank, mode):
        self.w = w = npb.World(0)
        w.set_option(w.OPT_BROADPHASE, mode); w.set_option(w.OPT_COUNTERS, 1)
        w.load_scene(sc)
        self.first, self.count, self.slot = padded_shard(w.n, rank, n_ranks)
        self.total = self.slot * n_ranks
        self.d = [w.dev_alloc(4 * self.total), w.dev_alloc(4 * self.total), w.dev_alloc(16 * self.total)]
        w.set_shard(self.first, self.count, *self.d)

    def my_records(self):
        out = [np.zeros(self.slot, np.int32), np.zeros(self.slot, np.int32), np.zeros((self.slot, 4), np.float32)]
        for a, p, width in zip(out, self.d, (4, 4, 16)):
            self.w.dev_download(a, p + width * self.first if self.count else p)      # an empty tail rank contributes padding
        return out

    def receive(self, gathered):
        for a, p in zip(gathered, self.d):
            if a is not None: self.w.dev_upload(p, a)


def _sharded_steps(npb, sc, n_ranks, mode, steps, check):
    ranks = [_Rank(npb, sc, n_ranks, r, mode) for r in range(n_ranks)]
    try:
        for k in range(steps):
            for r in ranks: r.w.step_begin(sc["dt"])
            pieces = [r.my_records() for r in ranks]
            gathered = [np.concatenate([p[c] for p in pieces]) for c in range(3)]
            gathered[1] = None                       # the flags array is scratch of the searching rank: only hit_j (success in bit 30) and contact travel
            for r in ranks: r.receive(gathered)
            for r in ranks: r.w.step_end()
            check(k, ranks)
        return [r.w.counters() for r in ranks]
    finally:
        for r in ranks: r.w.close()


@pytest.mark.parametrize("n,n_ranks", [(700, 2), (1001, 3), (5, 8)])
def test_sharded_exact_world_equals_reference_on_every_rank(npb, oracle, n, n_ranks):
    sc = scenes.lattice_scene(n, spacing=1.1, seed=n, mix=(0.5, 0.3, 0.2))
    ow = oracle.world(scenes.resolve(sc, oracle))

    def check(k, ranks):
        ao, bo, do = ow.step(sc["dt"], True); so = ow.state()
        for r in ranks:
            a, b, d = r.w.collisions_as_arrays()
            assert np.array_equal(a, ao) and np.array_equal(b, bo), "collision list, step %d" % k
            assert_biteq(d, do, "collision data, step %d" % k); assert_biteq(r.w.states(), so, "replicated state, step %d" % k)

    ctrs = _sharded_steps(npb, sc, n_ranks, 0, 8, check)
    pt, hits, _ = ow.counters()
    assert sum(c["pair_tests"] for c in ctrs) == pt and sum(c["hits"] for c in ctrs) == hits, "the ranks together ran the reference's DetectCollision calls, once each"
    if n >= 100:
        assert max(c["pair_tests"] for c in ctrs) < 0.75 * pt, "the search is not actually split"


def test_sharded_grid_world_equals_unsharded(npb):
    sc = scenes.lattice_scene(1500, spacing=0.9, seed=3, mix=(0.5, 0.2, 0.3))
    with npb.World(0) as one:
        one.set_option(one.OPT_BROADPHASE, one.GRID); one.load_scene(sc)

        def check(k, ranks):
            one.step(sc["dt"]); a1, b1, d1 = one.collisions_as_arrays(); s1 = one.states()
            assert len(a1) > 50
            for r in ranks:
                a, b, d = r.w.collisions_as_arrays()
                assert np.array_equal(a, a1) and np.array_equal(b, b1); assert_biteq(d, d1, "collision data"); assert_biteq(r.w.states(), s1, "state")

        _sharded_steps(npb, sc, 3, one.GRID, 4, check)


def test_shard_api_errors(npb):
    sc = scenes.lattice_scene(64, spacing=1.2, seed=1)
    with npb.World(0) as w:
        w.load_scene(sc)
        with pytest.raises(npb.NetphysError): w.step_end()                                   # no step in flight
        d = [w.dev_alloc(4 * 64), w.dev_alloc(4 * 64), w.dev_alloc(16 * 64)]
        with pytest.raises(npb.NetphysError): w.set_shard(0, 32, d[0], 0, d[2])             # missing array
        w.dev_upload(d[0], np.full(64, -1, np.int32)); w.dev_upload(d[1], np.zeros(64, np.int32)); w.dev_upload(d[2], np.zeros((64, 4), np.float32))
        w.set_shard(0, 32, *d)
        with pytest.raises(npb.NetphysError): w.step(sc["dt"])                               # a sharded world cannot run the monolithic step
        w.step_begin(sc["dt"])
        with pytest.raises(npb.NetphysError): w.step_begin(sc["dt"])
        with pytest.raises(npb.NetphysError): w.set_shard(0, -1)
        w.step_end()
        w.set_shard(0, -1)                                                            # back to a whole-world search
        w.step(sc["dt"]); w.synchronize()


def test_shard_changes_between_steps_and_bodies_added(npb, oracle):
    """The searched range may change between steps and bodies may be added: the search order is rebuilt for the range, results stay
    the reference's.  One world plays both ranks in turn (two half searches per step into the same arrays)."""
    sc = scenes.lattice_scene(1200, spacing=1.1, seed=9, mix=(0.5, 0.3, 0.2))
    ow = oracle.world(scenes.resolve(sc, oracle))
    with npb.World(0) as w, npb.World(0) as w2:
        for x in (w, w2): x.load_scene(sc)
        n = w.n
        d = [w.dev_alloc(4 * 4096), w.dev_alloc(4 * 4096), w.dev_alloc(16 * 4096)]
        d2 = [w2.dev_alloc(4 * 4096), w2.dev_alloc(4 * 4096), w2.dev_alloc(16 * 4096)]
        for k in range(6):
            cut = [n // 2, n // 3, 1, n - 1, n // 2, 700][k]                   # a different split every step
            w.set_shard(0, cut, *d); w2.set_shard(cut, n - cut, *d2)
            w.step_begin(sc["dt"]); w2.step_begin(sc["dt"])
            for width, p, p2 in zip((4, 16), (d[0], d[2]), (d2[0], d2[2])):    # the host plays the all-gather (hit_j and contact; flags stay local)
                a = np.zeros(width // 4 * cut, np.int32); b = np.zeros(width // 4 * (n - cut), np.int32)
                w.dev_download(a, p); w2.dev_download(b, p2 + width * cut)
                w.dev_upload(p + width * cut, b); w2.dev_upload(p2, a)
            w.step_end(); w2.step_end()
            ao, bo, do = ow.step(sc["dt"], True)
            for x in (w, w2):
                a, b, dd = x.collisions_as_arrays()
                assert np.array_equal(a, ao) and np.array_equal(b, bo); assert_biteq(dd, do, "collision data"); assert_biteq(x.states(), ow.state(), "state step %d" % k)


# ---- partitioned steps: every rank integrates / responds for its own bodies only, neighbours exchange halos ---------------------------
def _partitioned_steps(npb, sc, n_ranks, halo, steps, check):
    ranks = [_Rank(npb, sc, n_ranks, r, 0) for r in range(n_ranks)]
    try:
        for r in ranks: r.w.set_shard_halo(halo)
        n = ranks[0].w.n
        xf = [r.w.transforms_device() for r in ranks]
        for k in range(steps):
            for r in ranks: r.w.step_integrate(sc["dt"])
            for q, r in enumerate(ranks[:-1]):                       # the next rank's first `halo` transform records (64 B each) -> right after the own range
                nx = ranks[q + 1]; h = min(halo, nx.count)
                if h > 0:
                    buf = np.zeros((h, 16), np.float32); nx.w.dev_download(buf, xf[q + 1] + 64 * nx.first); r.w.dev_upload(xf[q] + 64 * nx.first, buf)
            for r in ranks: r.w.step_search()
            for q, r in enumerate(ranks[1:], 1):                     # the previous rank's last `halo` records -> right before the own range
                pv = ranks[q - 1]; h = min(halo, pv.count)
                if h > 0 and r.count > 0:
                    for p_src, p_dst, width, dt_ in ((pv.d[0], r.d[0], 4, np.int32), (pv.d[2], r.d[2], 16, np.float32)):
                        buf = np.zeros(h * width // 4, dt_); pv.w.dev_download(buf, p_src + width * (r.first - h)); r.w.dev_upload(p_dst + width * (r.first - h), buf)
            for r in ranks: r.w.step_end()
            check(k, ranks)
        return [r.w.counters() for r in ranks]
    finally:
        for r in ranks: r.w.close()


@pytest.mark.parametrize("n,n_ranks,halo", [(700, 2, 64), (1001, 3, 40), (2000, 4, 500), (49, 8, 7), (24, 3, 8)])
def test_partitioned_world_equals_reference_on_the_owned_bodies(npb, oracle, n, n_ranks, halo):
    """np_world_set_shard_halo: each rank steps its own index range (integrate, search, respond) and sees `halo` bodies past it; the host
    plays the two neighbour exchanges.  Own bodies' state and collisions == the oracle's single world, every step; everything else on a
    rank is allowed to be stale."""
    sc = scenes.lattice_scene(n, spacing=1.1, seed=n, mix=(0.5, 0.3, 0.2))
    ow = oracle.world(scenes.resolve(sc, oracle))

    def check(k, ranks):
        ao, bo, do = ow.step(sc["dt"], True); so = ow.state()
        got_a = []; got_b = []; got_d = []
        for r in ranks:
            r.w.synchronize()
            a, b, d = r.w.collisions_as_arrays()
            assert len(a) == 0 or (a.min() >= r.first and a.max() < r.first + r.count)
            got_a.append(a); got_b.append(b); got_d.append(d)
            assert_biteq(r.w.states()[r.first:r.first + r.count], so[r.first:r.first + r.count], "own bodies' state, step %d rank range %d+%d" % (k, r.first, r.count))
        assert np.array_equal(np.concatenate(got_a), ao) and np.array_equal(np.concatenate(got_b), bo), "collision list, step %d" % k
        assert_biteq(np.concatenate(got_d), do, "collision data, step %d" % k)

    ctrs = _partitioned_steps(npb, sc, n_ranks, halo, 8, check)
    pt, hits, _ = ow.counters()
    assert sum(c["pair_tests"] for c in ctrs) == pt and sum(c["hits"] for c in ctrs) == hits


def test_partitioned_world_reports_a_halo_that_is_too_small(npb):
    """a body whose first hit lies further ahead than the halo: the step must say so instead of returning a different collision list"""
    sc = scenes.lattice_scene(600, spacing=1.1, seed=600, mix=(0.5, 0.3, 0.2))
    with npb.World(0) as one:
        one.load_scene(sc); one.step(sc["dt"]); a, b, _ = one.collisions_as_arrays()
    reach = int((b - a).max())
    assert reach > 1, "workload no longer has a first hit beyond the next body"
    for halo, ok in ((reach, True), (reach - 1, False)):
        w = npb.World(0)
        try:
            w.load_scene(sc)
            d = [w.dev_alloc(4 * 600), w.dev_alloc(4 * 600), w.dev_alloc(16 * 600)]
            w.set_shard(0, 600, *d); w.set_shard_halo(halo)
            w.step_integrate(sc["dt"]); w.step_search(); w.step_end()
            if ok:
                w.synchronize(); a2, b2, _ = w.collisions_as_arrays()
                assert np.array_equal(a2, a) and np.array_equal(b2, b)
            else:
                with pytest.raises(npb.NetphysError):
                    w.synchronize()
        finally:
            w.close()
