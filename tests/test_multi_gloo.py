"""CPU, world_size 2 over gloo: the host-side logic of the multi-GPU modes (query sharding + result gather,
first-solution termination + path broadcast) with a stand-in planner.  The stand-in is a TEST DOUBLE, not a CPU path of
the product: it only imitates the call surface of cudasbmp_b200.KGMT so the collectives can be exercised without GPUs."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close(); return p


class FakePlanner:
    """Deterministic stand-in: a 'plan' solves after (seed % 5) + 2 iterations with cost seed * 0.5 + 1."""

    def __init__(self):
        self.seed, self.it, self.need = 0, 0, 0

    def set_seed(self, s):
        self.seed = s

    def plan(self, init, goal):
        return dict(stop=1, iterations=self.seed % 5 + 2, tree_size=100 + self.seed, cost_to_goal=self.seed * 0.5 + 1,
                    expansions=1000 * (self.seed + 1), device_ms=0.1)

    def begin(self, init, goal):
        self.it, self.need = 0, self.seed % 5 + 2

    def iterate_many(self, k):
        self.it = min(self.it + k, self.need)
        done = self.it >= self.need
        return dict(stop=1 if done else 0, cost_to_goal=self.seed * 0.5 + 1 if done else 0.0, iteration=self.it)

    def extract_path(self):
        return np.full((3, 7), float(self.seed), dtype=np.float32)


def _worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from cudasbmp_b200 import multi
    try:
        Q = 7
        inits = np.zeros((Q, 7), np.float32); goals = np.ones((Q, 7), np.float32)
        table = multi.plan_batch(FakePlanner(), inits, goals, seeds=list(range(10, 10 + Q)))
        port_res = multi.plan_portfolio(FakePlanner(), inits[0], goals[0], base_seed=3, check_every=1)
        out[rank] = (table, port_res)
    finally:
        dist.destroy_process_group()


def test_shard_range_is_a_partition():
    from cudasbmp_b200.multi import shard_range
    for n in (0, 1, 7, 8, 1024):
        for world in (1, 2, 3, 8):
            parts = [shard_range(n, r, world) for r in range(world)]
            assert parts[0][0] == 0 and parts[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(parts, parts[1:]))
            sizes = [b - a for a, b in parts]
            assert max(sizes) - min(sizes) <= 1


def test_batch_and_portfolio_over_gloo_world2():
    world, port = 2, _free_port()
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_worker, args=(world, port, out), nprocs=world, join=True)
    t0, p0 = out[0]
    t1, p1 = out[1]
    np.testing.assert_array_equal(t0, t1)                       # every rank holds the whole table
    assert list(t0[:, 0]) == list(range(7))
    assert list(t0[:, 1]) == [0, 0, 0, 0, 1, 1, 1]              # contiguous shards 4 + 3
    np.testing.assert_allclose(t0[:, 5], np.arange(10, 17) * 0.5 + 1)
    # portfolio: rank 0 has seed 3 (needs 5 iterations), rank 1 seed 4 (needs 6) -> rank 0 wins at check 5
    for p in (p0, p1):
        assert p["winner"] == 0 and p["checks"] == 5 and abs(p["cost"] - 2.5) < 1e-6
        assert p["path"].shape == (3, 7) and (p["path"] == 3.0).all()
    assert p1["stop"] == 0                                       # the loser was stopped early


def test_single_process_paths():
    from cudasbmp_b200 import multi
    t = multi.plan_batch(FakePlanner(), np.zeros((3, 7)), np.zeros((3, 7)), seeds=[1, 2, 3])
    assert t.shape == (3, 8) and list(t[:, 3]) == [3, 4, 5]
    p = multi.plan_portfolio(FakePlanner(), np.zeros(7), np.zeros(7), base_seed=1, check_every=2)
    assert p["winner"] == 0 and p["checks"] == 2


# ------------------------------------------------------------------------------------------ sharded expansion
class FakeShardPlanner:
    """TEST DOUBLE with the call surface of KGMT.shard_*: candidate slot s of an iteration is 'accepted' iff s % 3 == 0;
    every candidate increments delta[s % 16].  Buffers are CPU tensors addressed by data_ptr(), as on the GPU."""
    BLOCK = 8192

    def __init__(self):
        self.itr, self.tree, self.lo, self.hi = 1, [], 0, 0

    def M(self):
        return 20000 * self.itr + 37

    def shard_delta_ints(self):
        return 16

    @staticmethod
    def _view(ptr, n, ct=np.int32):
        import ctypes as C
        return np.ctypeslib.as_array((C.c_byte * (n * np.dtype(ct).itemsize)).from_address(ptr)).view(ct)

    def shard_expand(self, rank, world, delta_ptr):
        from cudasbmp_b200.multi import shard_range
        nb = (self.M() + self.BLOCK - 1) // self.BLOCK
        b0, b1 = shard_range(nb, rank, world)
        self.lo, self.hi = b0 * self.BLOCK, min(b1 * self.BLOCK, self.M())
        d = self._view(delta_ptr, 16)
        for s in range(self.lo, self.hi):
            d[s % 16] += 1
        self.acc = [s for s in range(self.lo, self.hi) if s % 3 == 0]
        return dict(iteration=self.itr, candidates=self.M(), accepted_local=len(self.acc), stop=0)

    def shard_pack(self, send_ptr, cap):
        buf = self._view(send_ptr, cap * 9)
        buf[cap * 8: cap * 8 + len(self.acc)] = self.acc          # the slot section starts at byte 32 * cap

    def shard_commit(self, recv_ptr, cap, counts, delta_ptr):
        buf = self._view(recv_ptr, len(counts) * cap * 9)
        for g, c in enumerate(counts):
            seg = buf[g * cap * 9:(g + 1) * cap * 9]
            self.tree.extend(int(v) for v in seg[cap * 8: cap * 8 + c])
        d = self._view(delta_ptr, 16)
        assert int(d.sum()) == self.M(), (int(d.sum()), self.M())
        d[:] = 0
        self.itr += 1
        return dict(iteration=self.itr - 1, accepted=sum(counts), tree_size=len(self.tree), stop=1 if self.itr > 3 else 0)


def _shard_worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from cudasbmp_b200.sharded import ShardedExpander
    try:
        p = FakeShardPlanner()
        hist = ShardedExpander(p, device=torch.device("cpu")).run()
        out[rank] = (p.tree, [(h["accepted"], h["accepted_local"], h["cap_rows"]) for h in hist])
    finally:
        dist.destroy_process_group()


def test_sharded_expansion_over_gloo_world2():
    world, port = 2, _free_port()
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_shard_worker, args=(world, port, out), nprocs=world, join=True)
    (t0, h0), (t1, h1) = out[0], out[1]
    assert t0 == t1                                              # replicated trees stay identical
    want = []
    for itr in (1, 2, 3):
        want += [s for s in range(20000 * itr + 37) if s % 3 == 0]
    assert t0 == want                                            # rank-major order == global candidate order
    assert [h[0] for h in h0] == [h[0] for h in h1] and len(h0) == 3
    assert all(a[1] + b[1] == a[0] for a, b in zip(h0, h1))     # local counts add up
    assert all(a[2] == b[2] and a[2] % 4 == 0 for a, b in zip(h0, h1))
