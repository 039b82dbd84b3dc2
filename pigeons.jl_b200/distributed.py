"""Sharding of the chain ladder across GPUs (one process per GPU).

Replaces the role of src/mpi_utils/ (LoadBalance, Entangler) for the scan path:
chains are split into contiguous blocks by the reference's own `LoadBalance`
rule, each process drives one engine handle on its GPU, and the only data-path
exchange — the single boundary pair per shard per swap phase — happens inside
the scan kernel through peer-mapped neighbour mailboxes (CUDA IPC over NVLink).
`torch.distributed` is plumbing only: IPC-handle exchange at start-up and the
once-per-round gather of the per-chain statistics (the role of
`all_reduce_deterministically`, src/mpi_utils/Entangler.jl:286-297).
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import List

import numpy as np


@dataclass(frozen=True)
class LoadBalance:
    """src/mpi_utils/LoadBalance.jl:36-128 (1-based process and global indices)."""
    my_process_index: int
    n_processes: int
    n_global_indices: int

    def __post_init__(self):
        assert 1 <= self.my_process_index <= self.n_processes <= self.n_global_indices

    def basic_load(self) -> int:
        return self.n_global_indices // self.n_processes

    def n_extras(self) -> int:
        return self.n_global_indices % self.n_processes

    def my_load(self) -> int:
        return self.basic_load() + (1 if self.my_process_index <= self.n_extras() else 0)

    def my_first_global_idx(self) -> int:
        before = self.my_process_index - 1
        with_extra = min(before, self.n_extras())
        return 1 + (before - with_extra) * self.basic_load() + with_extra * (self.basic_load() + 1)

    def my_global_indices(self) -> range:
        s = self.my_first_global_idx()
        return range(s, s + self.my_load())

    def find_process(self, global_idx: int) -> int:
        basic = self.basic_load()
        first_block = self.n_extras() * (basic + 1)
        if global_idx <= first_block:
            return 1 + (global_idx - 1) // (basic + 1)
        return 1 + self.n_extras() + (global_idx - first_block - 1) // basic


def shard_layout(n_chains: int, world_size: int, n_chains_variational: int = 0):
    """[(first_chain, n_local)] per rank as the engine lays a ladder out (`pgn_local_range`): LoadBalance's balanced
    blocks, except that a two-leg ladder keeps its two target chains (n_var, n_var + 1) on one shard — when a block
    boundary falls between them, chain n_var + 1 moves to the lower shard."""
    blocks = []
    for r in range(1, world_size + 1):
        lb = LoadBalance(r, world_size, n_chains)
        blocks.append([lb.my_first_global_idx(), lb.my_load()])
    nv = n_chains_variational
    if 0 < nv < n_chains:
        for r in range(world_size - 1):
            if blocks[r][0] + blocks[r][1] - 1 == nv:
                if blocks[r + 1][1] < 2:
                    raise ValueError("two legs: too few chains per shard to keep both target chains on one shard")
                blocks[r][1] += 1
                blocks[r + 1][0] += 1
                blocks[r + 1][1] -= 1
                break
    return [tuple(b) for b in blocks]


class Communicator:
    """Minimal collective surface the host driver needs."""
    rank: int = 0
    world_size: int = 1

    def all_gather_bytes(self, payload: bytes) -> List[bytes]:
        raise NotImplementedError

    def all_gather_array(self, a: np.ndarray) -> List[np.ndarray]:
        raise NotImplementedError

    def connect_neighbours(self, engine) -> None:
        """Exchange mailbox IPC handles with the left/right shard and attach them."""
        if self.world_size == 1:
            return
        handles = self.all_gather_bytes(engine.ipc_export())
        if self.rank > 0:
            engine.ipc_attach(0, handles[self.rank - 1])
        if self.rank < self.world_size - 1:
            engine.ipc_attach(1, handles[self.rank + 1])
        self.barrier()

    def barrier(self) -> None:
        pass


class SingleProcess(Communicator):
    def all_gather_bytes(self, payload):
        return [payload]

    def all_gather_array(self, a):
        return [a]


class TorchDistributed(Communicator):
    """torch.distributed (NCCL on GPUs, gloo on CPU) as the control-plane transport."""

    def __init__(self, device=None):
        import torch.distributed as dist
        assert dist.is_initialized(), "call torch.distributed.init_process_group first"
        self.dist = dist
        self.rank = dist.get_rank()
        self.world_size = dist.get_world_size()
        self.device = device    # torch.device for NCCL tensors, None for gloo/CPU

    def _tensor(self, a: np.ndarray):
        import torch
        t = torch.from_numpy(np.ascontiguousarray(a))
        return t.to(self.device) if self.device is not None else t

    def all_gather_array(self, a: np.ndarray) -> List[np.ndarray]:
        """Variable-length gather (shards may differ by one chain)."""
        import torch
        a = np.ascontiguousarray(a)
        flat = a.reshape(-1).view(np.uint8)
        sizes = [torch.zeros(1, dtype=torch.int64, device=self.device) for _ in range(self.world_size)]
        self.dist.all_gather(sizes, torch.tensor([flat.size], dtype=torch.int64, device=self.device))
        sizes = [int(s.item()) for s in sizes]
        mx = max(sizes)
        pad = np.zeros(mx, dtype=np.uint8)
        pad[: flat.size] = flat
        bufs = [torch.zeros(mx, dtype=torch.uint8, device=self.device) for _ in range(self.world_size)]
        self.dist.all_gather(bufs, self._tensor(pad))
        out = []
        for b, s in zip(bufs, sizes):
            raw = b.cpu().numpy()[:s]
            out.append(raw.view(a.dtype).copy())
        return out

    def all_gather_bytes(self, payload: bytes) -> List[bytes]:
        arrs = self.all_gather_array(np.frombuffer(payload, dtype=np.uint8))
        return [x.tobytes() for x in arrs]

    def barrier(self):
        self.dist.barrier()


class ThreadGroup:
    """Several shards driven from ONE process, one host thread per shard (the counterpart of a
    multi-threaded host that owns several GPUs, or several handles on one GPU).  Neighbouring
    handles are connected with `pgn_peer_attach` — the mailboxes are ordinary device pointers of
    the same process, so no CUDA IPC is involved — and the per-round gathers are plain copies
    between the threads."""

    def __init__(self, world_size: int):
        import threading
        self.world_size = int(world_size)
        self.barrier = threading.Barrier(self.world_size)
        self.slots = [None] * self.world_size
        self.engines = [None] * self.world_size

    def comm(self, rank: int) -> "ThreadComm":
        return ThreadComm(self, rank)

    def run(self, fn):
        """Run `fn(comm)` on world_size threads; returns the list of results in rank order.  The first
        exception aborts the barrier (so no thread is left waiting) and is re-raised."""
        import threading
        results, errors = [None] * self.world_size, [None] * self.world_size

        def work(r):
            try:
                results[r] = fn(self.comm(r))
            except BaseException as e:      # noqa: BLE001 - re-raised below
                errors[r] = e
                self.barrier.abort()

        threads = [threading.Thread(target=work, args=(r,)) for r in range(self.world_size)]
        for t in threads:
            t.start()
        for t in threads:
            t.join()
        import threading as _t
        first = [e for e in errors if e is not None and not isinstance(e, _t.BrokenBarrierError)]
        if first:
            raise first[0]
        if any(e is not None for e in errors):
            raise [e for e in errors if e is not None][0]
        return results


class ThreadComm(Communicator):
    def __init__(self, group: ThreadGroup, rank: int):
        self.group, self.rank, self.world_size = group, int(rank), group.world_size

    def all_gather_array(self, a: np.ndarray) -> List[np.ndarray]:
        g = self.group
        g.slots[self.rank] = np.array(a, copy=True)
        g.barrier.wait()
        out = [np.array(s, copy=True) for s in g.slots]
        g.barrier.wait()
        return out

    def all_gather_bytes(self, payload: bytes) -> List[bytes]:
        return [x.tobytes() for x in self.all_gather_array(np.frombuffer(payload, dtype=np.uint8))]

    def connect_neighbours(self, engine) -> None:
        g = self.group
        g.engines[self.rank] = engine
        g.barrier.wait()
        if self.rank > 0:
            engine.peer_attach(0, g.engines[self.rank - 1])
        if self.rank < self.world_size - 1:
            engine.peer_attach(1, g.engines[self.rank + 1])
        g.barrier.wait()

    def barrier(self) -> None:
        self.group.barrier.wait()
