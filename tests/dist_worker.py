"""Worker for tests/test_distributed_gloo.py: one rank of a world_size-N gloo job.

Every rank drives the host PT driver (`pigeons()` with adaptation on the merged
statistics) over `TorchDistributed`; the engine behind it is a *sharded view* of
the single-process CPU oracle (each rank exposes only its LoadBalance block of
chains), so the test exercises the N>1 host path — shard geometry, neighbour
hand-shake plumbing, per-round gather/concatenation, identical adaptation on all
ranks — and checks the invariance guarantee: the result equals the 1-process run.
"""
import json
import os
import sys

import numpy as np
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import pigeons_jl_b200 as pg              # noqa: E402
from oracle_adapter import load_oracle    # noqa: E402


class ShardedOracleView:
    """Engine-like object: full oracle ladder inside, one shard outside."""

    def __init__(self, lib, rank, world_size, n_chains, **kw):
        self.full = pg.Engine(lib, n_chains=n_chains, rank=0, world_size=1, **kw)
        nv = kw.get("n_chains_variational", 0)
        self.two_legs = 0 < nv < n_chains
        # the engine's layout rule (pgn_local_range): balanced blocks, two target chains kept on one shard
        self.layout = pg.shard_layout(n_chains, world_size, nv if self.two_legs else 0)
        self.first_chain, self.n_local = self.layout[rank]
        self.target_chain = nv if self.two_legs else n_chains
        self.dim = self.full.dim
        self.rank, self.world_size, self.n_chains = rank, world_size, n_chains
        self.attached = {}

    def __getattr__(self, name):          # set_schedule, set_explorer, init_replicas, close ...
        return getattr(self.full, name)

    def get_state(self):
        st = self.full.get_state()
        lo, hi = self.first_chain - 1, self.first_chain - 1 + self.n_local
        return {k: np.ascontiguousarray(v[lo:hi]) for k, v in st.items()}

    def ipc_export(self):
        return (b"mailbox-of-rank-%d" % self.rank).ljust(64, b"\0")

    def ipc_attach(self, side, handle):
        self.attached[side] = handle

    def run_round(self, n_scans, **logs):
        res = self.full.run_round(n_scans, **logs)
        lo, hi = self.first_chain - 1, self.first_chain - 1 + self.n_local
        for k in ("swap_n", "swap_mean", "logsum_fwd", "logsum_bwd", "expl_acc_n", "expl_acc_mean", "expl_n_steps",
                  "am_n", "am_mean", "rev_n", "rev_mean"):
            setattr(res, k, np.ascontiguousarray(getattr(res, k)[lo:hi]))
        for k in ("index_process", "swap_lr", "swap_u", "swap_accept"):
            a = getattr(res, k)
            if a is not None:
                setattr(res, k, np.ascontiguousarray(a[:, lo:hi]))
        owner_of_target = self.first_chain <= self.target_chain < self.first_chain + self.n_local
        if self.rank != 0:   # global counters are reported once (by the first shard)
            res.n_tempered_restarts = res.n_round_trips = 0
            res.n_density_points = res.n_ref_equiv_evals = 0
        if not owner_of_target:
            res.online_n = 0
            res.online_mean = np.zeros_like(res.online_mean)
            res.online_var = np.zeros_like(res.online_var)
            if res.target_trace is not None:
                res.target_trace = np.zeros_like(res.target_trace)
        return res


def main():
    dist.init_process_group("gloo")
    rank, world = dist.get_rank(), dist.get_world_size()
    out_path = sys.argv[1]
    lib = load_oracle()
    comm = pg.TorchDistributed()
    cases = {
        "toy_slice": dict(target=pg.toy_mvn_target(2), explorer=pg.SliceSampler(), n_chains=11, n_rounds=6, seed=1),
        "funnel_automala": dict(target=pg.Funnel(8), explorer=pg.AutoMALA(), n_chains=6, n_rounds=5, seed=2),
        # two legs of equal length: the balanced split separates the target chains, the layout rule re-unites them
        "two_legs_gmm_gaussian": dict(target=pg.eight_mode_mixture(6, 3.0), explorer=pg.AutoMALA(), n_chains=4, n_chains_variational=4,
                                      variational=pg.GaussianReference(first_tuning_round=2), n_rounds=5, seed=3),
        "two_legs_unid_slice": dict(target=pg.UnidentifiableProduct(100, 50), n_chains=4, n_chains_variational=3, n_rounds=5, seed=4),
    }
    report = {}
    for name, kw in cases.items():
        views = []

        def factory(**cfg):
            v = ShardedOracleView(lib, **{k: cfg.pop(k) for k in ("rank", "world_size", "n_chains")},
                                  **{k: v for k, v in cfg.items() if k != "device"})
            views.append(v)
            return v
        rec = [pg.index_process, pg.swap_trace, pg.traces]
        # checked_round: after round 3 rank 0 re-runs rounds 1..3 in one process and compares (src/pt/checks.jl:36-78)
        pt = pg.pigeons(engine_factory=factory, engine_lib=lib, comm=comm, record=rec, checked_round=3, **kw)
        v = views[0]
        # neighbour plumbing: every rank attached exactly its neighbours' handles
        want = {}
        if rank > 0:
            want[0] = (b"mailbox-of-rank-%d" % (rank - 1)).ljust(64, b"\0")
        if rank < world - 1:
            want[1] = (b"mailbox-of-rank-%d" % (rank + 1)).ljust(64, b"\0")
        assert v.attached == want, (v.attached, want)
        if rank == 0:
            ref = pg.pigeons(engine_lib=lib, record=rec, **kw)
            a, b = pt.reduced_recorders, ref.reduced_recorders
            same = all(np.array_equal(getattr(a, k), getattr(b, k)) for k in
                       ("index_process", "swap_lr", "swap_u", "swap_accept", "swap_n", "swap_mean", "logsum_fwd",
                        "logsum_bwd", "expl_n_steps", "am_mean", "online_mean", "online_var", "target_trace"))
            same = same and np.array_equal(pg.tempering_parameters(pt.shared.tempering), pg.tempering_parameters(ref.shared.tempering))
            same = same and pg.stepping_stone(pt) == pg.stepping_stone(ref)
            same = same and a.n_round_trips == b.n_round_trips and a.n_ref_equiv_evals == b.n_ref_equiv_evals
            report[name] = bool(same)
    # all ranks hold the same adapted schedule (adaptation ran on identical merged statistics)
    sched = comm.all_gather_array(pg.tempering_parameters(pt.shared.tempering))
    report_sched = all(np.array_equal(sched[0], s) for s in sched)
    if rank == 0:
        report["schedules_identical_on_all_ranks"] = bool(report_sched)
        with open(out_path, "w") as f:
            json.dump(report, f)
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
