"""Sharded expansion (BASELINE config 5): the candidates of ONE iteration are split over the ranks of a
torch.distributed group (NCCL over NVLink/NVSwitch on GPUs; gloo in the CPU tests with a stand-in planner).

Tree and region maps are replicated on every GPU and stay bit-identical to a single-GPU run: rank g owns a contiguous
range of candidate slots, so rank-major order is global candidate order and every rank inserts the gathered rows in
exactly the order the single-GPU scan would.  Per iteration:

    planner.shard_expand      stages 2-5a on the rank's slots; counter increments into a zeroed delta slab
    all_gather(counts)        4 bytes per rank (sizes the row exchange)
    planner.shard_pack        accepted rows -> send buffer, 36 B per row (float4 state | float4 ctrl+cost | int32 slot)
    all_gather_into_tensor    36 * cap bytes per rank, cap = max count rounded up to 4 rows
    all_reduce(SUM)           the delta slab: 4*N^2 + 4*N^2*n^2 int32
    planner.shard_commit      ordered insertion of every rank's rows, maps += deltas, goal test, next scores

The communication time is measured separately from the compute time (CUDA events on the shared stream) and reported as
it is: at K = 5 obstacles the exchange dominates, at K = 1 000 the expansion does (DESIGN.md "Multi-GPU").
"""
import numpy as np
import torch
import torch.distributed as dist

ROW_BYTES = 36


def round_up4(n):
    return max(4, (int(n) + 3) & ~3)


class ShardedExpander:
    """Drives kgmt_shard_expand / pack / commit around the three collectives.

    planner: cudasbmp_b200.KGMT after begin() / seed_frontier() (or any object with shard_delta_ints, shard_expand,
    shard_pack, shard_commit taking tensors' data_ptr()).  device: torch device of the exchange buffers
    (cuda:<local rank> with NCCL, cpu with gloo)."""

    def __init__(self, planner, group=None, device=None, timing=False):
        self.p = planner
        self.group = group
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        if device is None:
            nccl = dist.is_initialized() and dist.get_backend(group) == "nccl"
            device = torch.device("cuda", torch.cuda.current_device()) if (nccl or torch.cuda.is_available()) else torch.device("cpu")
        self.dev = device
        self.delta = torch.zeros(planner.shard_delta_ints(), dtype=torch.int32, device=device)
        self.counts = torch.zeros(self.world, dtype=torch.int32, device=device)
        self.mine = torch.zeros(1, dtype=torch.int32, device=device)
        self.send = self.recv = None
        self.cap = 0
        self.timing = timing and device.type == "cuda"
        self.last = {}
        if device.type == "cuda" and hasattr(planner, "set_stream"):
            # planner kernels and the collectives are ordered by ONE stream: torch's current one (handle 0 is the
            # legacy default stream, spelled cudaStreamLegacy = 1 when passed explicitly)
            planner.set_stream(torch.cuda.current_stream(device).cuda_stream or 1)

    def _buffers(self, cap):
        if cap > self.cap:
            self.cap = max(cap, 2 * self.cap)
            self.send = torch.empty(self.cap * ROW_BYTES, dtype=torch.uint8, device=self.dev)
            self.recv = torch.empty(self.world * self.cap * ROW_BYTES, dtype=torch.uint8, device=self.dev)

    def _ev(self):
        if not self.timing:
            return None
        e = torch.cuda.Event(enable_timing=True)
        e.record()
        return e

    def iterate(self):
        """One expansion iteration; returns the iteration stats dict (identical on every rank) with
        'accepted_local', 'cap_rows', and — with timing — 'compute_ms', 'comm_ms', 'comm_bytes'."""
        e0 = self._ev()
        info = self.p.shard_expand(self.rank, self.world, self.delta.data_ptr())
        e1 = self._ev()
        self.mine[0] = info["accepted_local"]
        if self.world > 1:
            dist.all_gather_into_tensor(self.counts, self.mine, group=self.group)
        else:
            self.counts.copy_(self.mine)
        counts = [int(c) for c in self.counts.tolist()]
        cap = round_up4(max(counts))
        self._buffers(cap)
        e2 = self._ev()
        send = self.send[:cap * ROW_BYTES]
        self.p.shard_pack(send.data_ptr(), cap)
        e3 = self._ev()
        if self.world > 1:
            recv = self.recv[:self.world * cap * ROW_BYTES]
            dist.all_gather_into_tensor(recv, send, group=self.group)
            dist.all_reduce(self.delta, op=dist.ReduceOp.SUM, group=self.group)
        else:
            recv = send
        e4 = self._ev()
        st = self.p.shard_commit(recv.data_ptr(), cap, counts, self.delta.data_ptr())
        e5 = self._ev()
        st["accepted_local"], st["cap_rows"] = info["accepted_local"], cap
        if self.timing:
            torch.cuda.synchronize()
            st["expand_ms"], st["pack_ms"], st["commit_ms"] = e0.elapsed_time(e1), e2.elapsed_time(e3), e4.elapsed_time(e5)
            st["compute_ms"] = st["expand_ms"] + st["pack_ms"] + st["commit_ms"]
            st["comm_ms"] = e1.elapsed_time(e2) + e3.elapsed_time(e4)
            st["comm_bytes"] = (self.world * cap * ROW_BYTES + self.delta.numel() * 4 + 4 * self.world) if self.world > 1 else 0
        self.last = st
        return st

    def run(self, max_iterations=1 << 30):
        """Iterate until the planner stops; returns the list of per-iteration stats."""
        out = []
        while len(out) < max_iterations:
            st = self.iterate()
            out.append(st)
            if st["stop"] != 0:
                break
        return out


class PeerExpander:
    """The same sharded expansion with the exchange done by the library's own kernels over peer memory (NVLink /
    NVSwitch): torch.distributed is used ONCE, to gather the cudaIpc handles; no collective on the data path.

    planner: cudasbmp_b200.KGMT (one per process / GPU).  Every rank calls iterate() for the same iteration."""

    def __init__(self, planner, group=None, timing=False):
        self.p = planner
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        mine = planner.peer_export()
        if self.world > 1:
            dev = torch.device("cuda", torch.cuda.current_device()) if dist.get_backend(group) == "nccl" else torch.device("cpu")
            t = torch.tensor(list(mine), dtype=torch.uint8, device=dev)
            allh = torch.empty(self.world * t.numel(), dtype=torch.uint8, device=dev)
            dist.all_gather_into_tensor(allh, t, group=group)
            handles = bytes(allh.cpu().tolist())
        else:
            handles = mine
        planner.peer_attach(self.rank, self.world, handles)
        if self.world > 1:
            dist.barrier(group=group)          # every rank has mapped every other before the first exchange
        self.timing = timing
        if timing:
            planner.set_stream(torch.cuda.current_stream().cuda_stream or 1)

    def iterate(self):
        if self.timing:
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
        self.p.peer_expand_begin()
        if self.timing:
            e1.record()
        st = self.p.peer_expand_end()
        if self.timing:
            torch.cuda.synchronize()
            st["total_ms"] = e0.elapsed_time(e1)
        return st

    def run(self, max_iterations=1 << 30):
        out = []
        while len(out) < max_iterations:
            st = self.iterate()
            out.append(st)
            if st["stop"] != 0:
                break
        return out

    def close(self):
        self.p.peer_detach()
