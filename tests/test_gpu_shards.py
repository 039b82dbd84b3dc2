"""Several shards of one ladder on ONE GPU (pgn_peer_attach), driven by one host thread each.

This is how the driver's single-GPU `pytest -m gpu` exercises what otherwise needs several GPUs: the
`remote` branch of the mailbox hand-shake (ghost slots of the neighbouring handle), the per-round gather of
the shards' statistics, `checked_round` on genuinely sharded engines (`src/pt/checks.jl:36-78`: every shard's
`pgn_get_state` gathered and compared with a serial re-run), and the reference's invariance guarantee
(docs/src/distributed.md:37-55, test/test_parallelism_invariance.jl:27-45): the merged result is identical to
the single-process oracle's for any number of shards."""
import numpy as np
import pytest

import pigeons_jl_b200 as pg

pytestmark = pytest.mark.gpu

KEYS = ["index_process", "swap_lr", "swap_u", "swap_accept", "swap_n", "swap_mean", "logsum_fwd", "logsum_bwd",
        "expl_acc_n", "expl_acc_mean", "expl_n_steps", "am_n", "am_mean", "rev_n", "rev_mean", "online_mean", "online_var"]

CASES = {
    "toy_slice_n11": dict(target=pg.toy_mvn_target(2), explorer=pg.SliceSampler(), n_chains=11, n_rounds=6, seed=1),
    "funnel_automala_n12": dict(target=pg.Funnel(32), explorer=pg.AutoMALA(), n_chains=12, n_rounds=5, seed=2),
    "gmm_automala_n7": dict(target=pg.eight_mode_mixture(128, 8.0), explorer=pg.AutoMALA(), n_chains=7, n_rounds=4, seed=3),
    "ising_n10": dict(target=pg.IsingLogPotential(1.0, 5), n_chains=10, n_rounds=6, seed=4),
    "test_swapper_n8": dict(target=pg.TestSwapper(0.5), n_chains=8, n_rounds=6, seed=5),
    "mixed_slice_n8": dict(target=pg.MixedProduct(n_bool=3, n_int=2, n_float=2), n_chains=8, n_rounds=6, seed=8),
    "logreg_automala_n7": dict(target=pg.synthetic_logistic_regression(300, 24), explorer=pg.AutoMALA(), n_chains=7,
                               n_rounds=4, seed=6),
    # two legs: 11 chains, targets 5 and 6 — on one shard for world 2 (1-6 | 7-11) and world 3 (1-4 | 5-8 | 9-11)
    "two_legs_gmm_gaussian_n11": dict(target=pg.eight_mode_mixture(6, 3.0), explorer=pg.AutoMALA(), n_chains=6, n_chains_variational=5,
                                      variational=pg.GaussianReference(first_tuning_round=2), n_rounds=5, seed=9),
    # equal legs on two shards: the balanced boundary 5 | 6 falls between the targets, chain 6 joins the lower shard (1-6 | 7-10)
    "two_legs_funnel_slice_n10": dict(target=pg.Funnel(8), explorer=pg.SliceSampler(), n_chains=5, n_chains_variational=5,
                                      variational=pg.GaussianReference(first_tuning_round=3), n_rounds=5, seed=10),
    "toy300_automala_n6_mem": dict(target=pg.toy_mvn_target(300), explorer=pg.AutoMALA(), n_chains=6, n_rounds=4, seed=7),
}


def run_sharded(lib, world, record, checked_round=0, engine_factory=None, **kw):
    group = pg.ThreadGroup(world)

    def one_rank(comm):
        pt = pg.pigeons(engine_lib=lib, comm=comm, record=record, checked_round=checked_round,
                        engine_factory=engine_factory, **kw)
        out = dict(rr=pt.reduced_recorders, schedule=pg.tempering_parameters(pt.shared.tempering).copy(), logz=pg.stepping_stone(pt),
                   state=pt.engine.get_state(), first=pt.engine.first_chain)
        comm.barrier()          # nobody destroys a mailbox a neighbour's kernel may still read
        pt.close()
        return out
    return group.run(one_rank)


# The LOGREG path is a sequence of ordinary kernels per shard (controller, GEMMs, post, decide): with three shards on one
# device the decide kernels of two shards spin while the third still has a dozen launches queued, and whether those run
# next to the spinning ones is up to the device's queue assignment (CUDA does not promise it) — on separate GPUs the
# question does not arise (tests/test_multigpu.py covers world 2/4/8 there).  One device: two LOGREG shards.
PAIRS = [(n, w) for n in CASES for w in (2, 3) if not (n.startswith("logreg") and w == 3)]


@pytest.mark.parametrize("name,world", PAIRS)
def test_shards_on_one_gpu_match_the_oracle(name, world, gpu_lib, oracle_lib):
    kw = CASES[name]
    rec = [pg.index_process, pg.swap_trace] if name.startswith("test_swapper") else [pg.index_process, pg.swap_trace, pg.traces]
    shards = run_sharded(gpu_lib, world, rec, checked_round=3, **kw)
    ref = pg.pigeons(engine_lib=oracle_lib, record=rec, **kw)
    b = ref.reduced_recorders
    for r, sh in enumerate(shards):                 # every rank holds the same merged recorders
        a = sh["rr"]
        for k in KEYS:
            assert np.array_equal(getattr(a, k), getattr(b, k)), f"{name} world {world} rank {r}: {k}"
        if a.target_trace is not None:
            assert np.array_equal(a.target_trace, b.target_trace)
        assert a.n_round_trips == b.n_round_trips and a.n_tempered_restarts == b.n_tempered_restarts
        assert a.n_ref_equiv_evals == b.n_ref_equiv_evals
        assert np.array_equal(sh["schedule"], pg.tempering_parameters(ref.shared.tempering))
        assert sh["logz"] == pg.stepping_stone(ref) or (np.isnan(sh["logz"]) and np.isnan(pg.stepping_stone(ref)))
    whole = {k: np.concatenate([sh["state"][k] for sh in shards], axis=0) for k in shards[0]["state"]}
    rs = ref.engine.get_state()
    for k in whole:
        assert np.array_equal(whole[k].reshape(rs[k].shape), rs[k]), f"{name} world {world}: final replica {k}"
    ref.close()


def test_checked_round_catches_a_perturbed_shard(gpu_lib):
    """Negative control of `run_checks` on sharded engines: one shard reports a state that differs in one
    coordinate -> every rank raises ChecksFailed."""
    kw = dict(target=pg.toy_mvn_target(3), explorer=pg.SliceSampler(), n_chains=8, n_rounds=4, seed=3)

    class Perturbed:
        def __init__(self, **cfg):
            self.e = pg.Engine(gpu_lib, **cfg)
            self.bad = cfg["world_size"] > 1 and cfg["rank"] == 1

        def __getattr__(self, name):
            return getattr(self.e, name)

        def get_state(self):
            st = self.e.get_state()
            if self.bad:
                st["x"][0, 0] = np.nextafter(st["x"][0, 0], np.inf)
            return st
    with pytest.raises(pg.ChecksFailed):
        run_sharded(gpu_lib, 2, [], checked_round=2, engine_factory=lambda **cfg: Perturbed(**cfg), **kw)
