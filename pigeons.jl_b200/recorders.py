"""Round statistics after `reduce_recorders!` (src/recorders/recorders.jl:88-120).

The engine returns fixed-layout arrays per shard (per local chain / per local
pair).  Because statistics are keyed by chain (or by pair, stored at the lower
chain) and every chain lives on exactly one shard, the cross-shard reduction is
a concatenation in chain order — no floating-point reduction crosses GPUs, so
the result is identical for any number of shards (the invariance guarantee of
docs/src/distributed.md:37-55).
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Optional

import numpy as np


@dataclass
class ReducedRecorders:
    n_scans: int
    swap_n: np.ndarray            # [N] pair (i,i+1) stored at i (last entry unused)
    swap_mean: np.ndarray         # swap_acceptance_pr
    logsum_fwd: np.ndarray        # log_sum_ratio[(i,i+1)]
    logsum_bwd: np.ndarray        # log_sum_ratio[(i+1,i)]
    expl_acc_n: np.ndarray
    expl_acc_mean: np.ndarray     # explorer_acceptance_pr
    expl_n_steps: np.ndarray      # explorer_n_steps
    am_n: np.ndarray
    am_mean: np.ndarray           # am_factors
    rev_n: np.ndarray
    rev_mean: np.ndarray          # reversibility_rate
    n_tempered_restarts: int
    n_round_trips: int
    online_n: int
    online_mean: np.ndarray       # [d] target chain
    online_var: np.ndarray
    index_process: Optional[np.ndarray]   # [n_scans, N]
    swap_lr: Optional[np.ndarray]
    swap_u: Optional[np.ndarray]
    swap_accept: Optional[np.ndarray]
    target_trace: Optional[np.ndarray]    # [n_scans, d]; two legs: [n_scans, 2, d] (chains n_var, n_var + 1)
    n_density_points: int
    n_ref_equiv_evals: int
    kernel_ms: float
    wall_s: float

    @property
    def has_swap_stats(self) -> bool:
        return bool(np.any(self.swap_n > 0))


_PER_CHAIN = ["swap_n", "swap_mean", "logsum_fwd", "logsum_bwd", "expl_acc_n", "expl_acc_mean", "expl_n_steps",
              "am_n", "am_mean", "rev_n", "rev_mean"]
_PER_SCAN_CHAIN = ["index_process", "swap_lr", "swap_u", "swap_accept"]


def merge_round_results(comm, res, n_chains: int, dim: int, n_chains_variational: int = 0) -> ReducedRecorders:
    """Concatenate the shards' arrays in chain order (rank order == chain order).  The target-chain recorders come from
    the shard owning the target chain(s): chain N, or with two legs chains n_var and n_var + 1 (one shard, engine rule)."""
    two_legs = 0 < n_chains_variational < n_chains
    if comm.world_size == 1:
        g = {k: getattr(res, k) for k in _PER_CHAIN + _PER_SCAN_CHAIN}
        restarts, trips = res.n_tempered_restarts, res.n_round_trips
        online_n, online_mean, online_var, trace = res.online_n, res.online_mean, res.online_var, res.target_trace
        pts, evals, kms, wall = res.n_density_points, res.n_ref_equiv_evals, res.kernel_ms, res.wall_s
    else:
        g = {}
        for k in _PER_CHAIN:
            g[k] = np.concatenate(comm.all_gather_array(getattr(res, k)))
        for k in _PER_SCAN_CHAIN:
            a = getattr(res, k)
            if a is None:
                g[k] = None
            else:
                parts = comm.all_gather_array(a)
                parts = [p.reshape(res.n_scans, -1) for p in parts]
                g[k] = np.concatenate(parts, axis=1)
        scal = np.array([res.n_tempered_restarts, res.n_round_trips, res.n_density_points, res.n_ref_equiv_evals],
                        dtype=np.int64)
        tot = np.sum(np.stack(comm.all_gather_array(scal)), axis=0)
        restarts, trips, pts, evals = (int(v) for v in tot)
        times = np.stack(comm.all_gather_array(np.array([res.kernel_ms, res.wall_s])))
        kms, wall = float(times[:, 0].max()), float(times[:, 1].max())
        from .distributed import shard_layout
        t_chain = n_chains_variational if two_legs else n_chains
        last = next(r for r, (f, n) in enumerate(shard_layout(n_chains, comm.world_size, n_chains_variational if two_legs else 0))
                    if f <= t_chain < f + n)
        online_n = int(comm.all_gather_array(np.array([res.online_n], dtype=np.int64))[last][0])
        online_mean = comm.all_gather_array(res.online_mean)[last]
        online_var = comm.all_gather_array(res.online_var)[last]
        trace = None
        if res.target_trace is not None:
            trace = comm.all_gather_array(res.target_trace)[last].reshape((res.n_scans, 2, dim) if two_legs else (res.n_scans, dim))
    for k in _PER_CHAIN:
        assert g[k].shape[0] == n_chains, (k, g[k].shape, n_chains)
    return ReducedRecorders(
        n_scans=res.n_scans, n_tempered_restarts=restarts, n_round_trips=trips, online_n=online_n,
        online_mean=online_mean, online_var=online_var, target_trace=trace, n_density_points=pts,
        n_ref_equiv_evals=evals, kernel_ms=kms, wall_s=wall, **g)
