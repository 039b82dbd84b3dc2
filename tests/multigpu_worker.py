"""One rank of the multi-GPU parity job (tests/test_multigpu.py): the chain ladder is
sharded over WORLD_SIZE GPUs (one process per GPU, mailboxes peer-mapped with CUDA
IPC), and rank 0 checks the merged result against the single-process CPU oracle."""
import json
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import pigeons_jl_b200 as pg              # noqa: E402
from oracle_adapter import load_oracle    # noqa: E402


def main():
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local_rank)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    rank, world = dist.get_rank(), dist.get_world_size()
    comm = pg.TorchDistributed(device=torch.device("cuda", local_rank))
    lib = pg.EngineLib()
    cases = {
        "toy_slice_n11": dict(target=pg.toy_mvn_target(2), explorer=pg.SliceSampler(), n_chains=11, n_rounds=7, seed=1),
        "funnel_automala_n24": dict(target=pg.Funnel(32), explorer=pg.AutoMALA(), n_chains=24, n_rounds=6, seed=2),
        "gmm_automala_n9": dict(target=pg.eight_mode_mixture(128, 8.0), explorer=pg.AutoMALA(), n_chains=9, n_rounds=4, seed=3),
        "logreg_automala_n9": dict(target=pg.synthetic_logistic_regression(300, 24), explorer=pg.AutoMALA(), n_chains=9,
                                   n_rounds=4, seed=6),
        "toy300_automala_n9_mem": dict(target=pg.toy_mvn_target(300), explorer=pg.AutoMALA(), n_chains=9, n_rounds=4, seed=7),
        "ising_n10": dict(target=pg.IsingLogPotential(1.0, 5), n_chains=10, n_rounds=6, seed=4),
        "test_swapper_n8": dict(target=pg.TestSwapper(0.5), n_chains=8, n_rounds=6, seed=5),
        "mixed_slice_n9": dict(target=pg.MixedProduct(n_bool=3, n_int=2, n_float=2), n_chains=9, n_rounds=6, seed=8),
        # two legs of 8 chains: the balanced split separates the target chains 8 | 9 at every world size, chain 9 joins chain 8's GPU
        "two_legs_gmm_gaussian_n16": dict(target=pg.eight_mode_mixture(6, 3.0), explorer=pg.AutoMALA(), n_chains=8, n_chains_variational=8,
                                          variational=pg.GaussianReference(first_tuning_round=2), n_rounds=5, seed=11),
        "funnel_automala_n24_per_chain": dict(target=pg.Funnel(32), explorer=pg.AutoMALA(), n_chains=24, n_rounds=5, seed=9,
                                              recorder_order=1),
    }
    rec = [pg.index_process, pg.swap_trace, pg.traces]
    report = {}
    for name, kw in cases.items():
        r = [pg.index_process, pg.swap_trace] if name.startswith("test_swapper") else rec
        # checked_round: after round 3 rank 0 re-runs rounds 1..3 on ONE device and every shard's replicas (gathered from the
        # per-GPU pgn_get_state), the schedule and the explorer must equal the serial run (src/pt/checks.jl:36-78)
        pt = pg.pigeons(engine_lib=lib, comm=comm, device=local_rank, record=r, checked_round=3, **kw)
        if rank == 0:
            ref = pg.pigeons(engine_lib=load_oracle(), record=r, **kw)
            a, b = pt.reduced_recorders, ref.reduced_recorders
            keys = ["index_process", "swap_lr", "swap_u", "swap_accept", "swap_n", "swap_mean", "logsum_fwd",
                    "logsum_bwd", "expl_n_steps", "am_mean", "online_mean", "online_var"]
            if a.target_trace is not None:
                keys.append("target_trace")
            bad = [k for k in keys if not np.array_equal(getattr(a, k), getattr(b, k))]
            if not np.array_equal(pg.tempering_parameters(pt.shared.tempering), pg.tempering_parameters(ref.shared.tempering)):
                bad.append("schedule")
            if a.n_round_trips != b.n_round_trips or a.n_ref_equiv_evals != b.n_ref_equiv_evals:
                bad.append("counters")
            report[name] = bad
        pt.close()
        comm.barrier()
    if rank == 0:
        with open(sys.argv[1], "w") as f:
            json.dump(report, f)
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
