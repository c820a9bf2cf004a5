"""Distro sharding across the GPUs of one box.

Distros are independent in both the planner (one amboy job per distro,
units/crons.go:303-332) and the allocator (units/crons.go:274-301), so the path
shards by whole distros with no data-path exchange.  The only collective is one
all-gather of the per-distro evg_alloc_result vector (16 B per distro) so every
rank ends the tick holding every distro's (new_hosts, free_hosts, deficit).

`torch.distributed` is plumbing here: NCCL over NVLink on the GPU box, gloo in
the CPU tests.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import List, Sequence

import numpy as np

RESULT_BYTES = 16  # sizeof(evg_alloc_result)


@dataclass
class Shards:
    world: int
    owner: np.ndarray          # [D] rank owning each distro
    slot: np.ndarray           # [D] position of the distro inside its rank's shard
    members: List[np.ndarray]  # per rank: global distro ids in shard order
    load: np.ndarray           # per rank: summed weight

    @property
    def max_shard(self) -> int:
        return max((len(m) for m in self.members), default=0)


def lpt_partition(weights: Sequence[int], world: int) -> Shards:
    """Longest-processing-time bin packing of whole distros onto `world` ranks.
    weight = tasks (+ hosts) of the distro; a distro never spans ranks because its
    sort must not cross GPUs.  Deterministic: ties go to the lower distro id / rank."""
    w = np.asarray(weights, dtype=np.int64)
    D = int(w.shape[0])
    order = np.lexsort((np.arange(D), -w))  # heaviest first, stable on id
    load = np.zeros(world, dtype=np.int64)
    owner = np.zeros(D, dtype=np.int64)
    if D >= 4 * world and w.size and w.max() * world <= max(1, int(w.sum())) // 8:
        # many comparable distros: round-robin over the sorted list is LPT-equivalent and O(D)
        owner[order] = np.arange(D) % world
        np.add.at(load, owner, w)
    else:
        import heapq
        heap = [(0, r) for r in range(world)]
        for d in order.tolist():
            l, r = heapq.heappop(heap)
            owner[d] = r
            load[r] = l + int(w[d])
            heapq.heappush(heap, (int(load[r]), r))
    members = [np.nonzero(owner == r)[0] for r in range(world)]
    slot = np.zeros(D, dtype=np.int64)
    for m in members:
        slot[m] = np.arange(len(m))
    return Shards(world, owner, slot, members, load)


class ResultGather:
    """The per-tick collective: one all-gather of the padded per-rank result vectors.

    `send` is a uint8 tensor [max_shard * 16] on the rank's device -- bind it with
    Engine.bind_result_buffer so the allocator kernel writes into it directly (rows past
    this rank's shard are padding).  gather() returns a uint8 tensor [D * 16] in GLOBAL
    distro order on the same device; all buffers are allocated once."""

    def __init__(self, shards: Shards, device):
        import torch
        self.shards = shards
        self.pad = shards.max_shard * RESULT_BYTES
        self.send = torch.zeros(max(self.pad, 1), dtype=torch.uint8, device=device)[:self.pad]
        self.gathered = torch.empty(shards.world * self.pad, dtype=torch.uint8, device=device)
        self.index = torch.as_tensor(shards.owner * shards.max_shard + shards.slot, device=device)
        # the reordered result lives in a buffer of this slot, not in a fresh allocation per tick: a block freed back to
        # the communication stream's pool could be rewritten by the next gather while another stream still reads it
        self.result = torch.empty((int(self.index.shape[0]), RESULT_BYTES), dtype=torch.uint8, device=device)

    def gather(self):
        import torch.distributed as dist
        if self.shards.world > 1:
            dist.all_gather_into_tensor(self.gathered, self.send)
            rows = self.gathered.view(self.shards.world * self.shards.max_shard, RESULT_BYTES)
        else:
            rows = self.send.view(self.shards.max_shard, RESULT_BYTES)
        import torch
        torch.index_select(rows, 0, self.index, out=self.result)
        return self.result.reshape(-1)


class PipelinedGather:
    """ResultGather with `depth` buffer sets and the collective on its own CUDA stream: tick k's all-gather (and
    the reorder to global distro order) runs while tick k+1's planner already occupies the SMs -- the gather is
    16 B per distro, all launch latency, and nothing in tick k+1 depends on it.

    Per tick k:  before_tick(k, main)  -> bind send(k) as the allocator's result buffer -> run the tick on `main`
                 -> launch(k, main).   drain(main) makes `main` wait for every gather still in flight
    (call it before the closing timing event).  On CPU tensors (the gloo tests) everything is synchronous."""

    def __init__(self, shards: Shards, device, depth: int = 2):
        import torch
        self.torch = torch
        self.depth = depth
        self.slots = [ResultGather(shards, device) for _ in range(depth)]
        self.cuda = torch.device(device).type == "cuda"
        self.out = [None] * depth
        self.done = [None] * depth
        if self.cuda:
            self.comm = torch.cuda.Stream(device)
            self.ready = [torch.cuda.Event() for _ in range(depth)]

    def send(self, k: int):
        return self.slots[k % self.depth].send

    def before_tick(self, k: int, main) -> None:
        """Tick k overwrites send(k) and, later, result(k): the gather that last used that slot (tick k - depth) must
        be done before `main` goes on, and whatever `main` has queued so far (readers of that slot's result) must be done
        before the communication stream rewrites it."""
        d = self.done[k % self.depth]
        if self.cuda and d is not None:
            main.wait_event(d)

    def launch(self, k: int, main) -> None:
        i = k % self.depth
        if not self.cuda:
            self.out[i] = self.slots[i].gather()
            return
        torch = self.torch
        self.ready[i].record(main)  # orders the gather behind tick k's kernels AND behind every earlier reader of slot i on `main`
        self.comm.wait_event(self.ready[i])
        with torch.cuda.stream(self.comm):
            self.out[i] = self.slots[i].gather()
            done = torch.cuda.Event()
            done.record(self.comm)
        self.done[i] = done

    def drain(self, main) -> None:
        if self.cuda:
            for d in self.done:
                if d is not None:
                    main.wait_event(d)

    def result(self, k: int):
        """Gathered rows of tick k, global distro order (valid once the stream that waits on drain() has caught up)."""
        return self.out[k % self.depth]


def all_gather_results(local, shards: Shards, rank: int):
    """One-shot form of ResultGather (allocates; used by the CPU tests)."""
    g = ResultGather(shards, local.device)
    g.send.copy_(local)
    return g.gather()


def decode_results(buf) -> np.ndarray:
    """uint8 tensor [D*16] -> numpy structured array (new_hosts, free_hosts, deficit_ns)."""
    from . import _lib as L
    return np.frombuffer(buf.cpu().numpy().tobytes(), dtype=L.ALLOC_RESULT_DTYPE)
