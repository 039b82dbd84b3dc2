"""Shared by tests/golden/make_golden.py and tests/test_golden.py."""
import hashlib

import numpy as np

import pigeons_jl_b200 as pg

GOLDEN_CASES = {
    # BASELINE config 1: toy_mvn_target(2), 10 chains, SliceSampler, 10 rounds (2046 scans)
    "c1_toy_mvn2_slice_n10_r10": lambda: dict(target=pg.toy_mvn_target(2), explorer=pg.SliceSampler(), n_chains=10,
                                              n_rounds=10, seed=1),
    # BASELINE config 2 shape, reduced chain count
    "c2_funnel32_automala_n32_r7": lambda: dict(target=pg.Funnel(32), explorer=pg.AutoMALA(), n_chains=32, n_rounds=7, seed=1),
    # BASELINE config 3 shape, reduced chain count
    "c3_gmm128_automala_n16_r5": lambda: dict(target=pg.eight_mode_mixture(128, 8.0), explorer=pg.AutoMALA(), n_chains=16,
                                              n_rounds=5, seed=1),
    # BASELINE config 4 shape, reduced chain count
    "c4_ising32_n8_r4": lambda: dict(target=pg.IsingLogPotential(0.4406867935097715, 32), n_chains=8, n_rounds=4, seed=1),
    # BASELINE config 5 shape, reduced sizes
    "c5_logreg_d24_n300_automala_n6_r5": lambda: dict(target=pg.synthetic_logistic_regression(300, 24), explorer=pg.AutoMALA(),
                                                      n_chains=6, n_rounds=5, seed=1),
    # the explorer combinations of test/test_parallelism_invariance.jl:14-19
    "toy3_compose_slice_automala_n4_r8": lambda: dict(target=pg.toy_mvn_target(3), explorer=pg.Compose(pg.SliceSampler(), pg.AutoMALA()),
                                                      n_chains=4, n_rounds=8, seed=1),
    "funnel8_mix_automala_n6_r7": lambda: dict(
        target=pg.Funnel(8), n_chains=6, n_rounds=7, seed=1,
        explorer=pg.Mix(pg.AutoMALA(preconditioner=pg.IdentityPreconditioner(), base_n_refresh=1),
                        pg.AutoMALA(preconditioner=pg.MixDiagonalPreconditioner(0.0, 0.0), base_n_refresh=1),
                        pg.AutoMALA(preconditioner=pg.DiagonalPreconditioner(), base_n_refresh=1))),
    "gmm6_mix_slice_automala_mala_n6_r6": lambda: dict(
        target=pg.eight_mode_mixture(6, 3.0), n_chains=6, n_rounds=6, seed=6,
        explorer=pg.Mix(pg.SliceSampler(n_passes=1), pg.AutoMALA(base_n_refresh=1), pg.MALA(step_size=0.3, base_n_refresh=2))),
    # SliceSampler on Bool / Integer / Float coordinates (test/test_slice_sampler.jl:56-75)
    "mixed_bool_int_float_slice_n7_r8": lambda: dict(target=pg.MixedProduct(n_bool=3, n_int=2, n_float=2), n_chains=7,
                                                     n_rounds=8, seed=1),
    # two legs + GaussianReference (StabilizedPT.jl, GaussianReference.jl)
    "two_legs_gmm6_automala_gaussian_r7": lambda: dict(target=pg.eight_mode_mixture(6, 3.0), explorer=pg.AutoMALA(), n_chains=5,
                                                       n_chains_variational=5, variational=pg.GaussianReference(first_tuning_round=2),
                                                       n_rounds=7, seed=4),
    "two_legs_funnel8_slice_gaussian_r7": lambda: dict(target=pg.Funnel(8), explorer=pg.SliceSampler(), n_chains=5, n_chains_variational=4,
                                                       variational=pg.GaussianReference(first_tuning_round=3), n_rounds=7, seed=3),
    # test/test_DistributionLogPotential.jl:7-21 (global barrier 1.39 +- 0.1 in the reference)
    "unid_dlp_multivariate_n4_r10": lambda: dict(target=pg.UnidentifiableProduct(100, 50), n_chains=4, n_rounds=10, seed=1),
    "ising5_n10_r8": lambda: dict(target=pg.IsingLogPotential(1.0, 5), n_chains=10, n_rounds=8, seed=1),
}


def _hex(a):
    return [float(v).hex() for v in np.asarray(a, dtype=np.float64).ravel()]


def summarise(pt):
    rr = pt.reduced_recorders
    return {
        "n_scans_last_round": int(rr.n_scans),
        "index_process_sha256": hashlib.sha256(np.ascontiguousarray(rr.index_process, dtype=np.int32).tobytes()).hexdigest(),
        "index_process_head": rr.index_process[:16].tolist(),
        "swap_accept_sha256": hashlib.sha256(np.ascontiguousarray(rr.swap_accept, dtype=np.uint8).tobytes()).hexdigest(),
        "swap_u_head": _hex(rr.swap_u[:4]),
        "swap_lr_head": _hex(rr.swap_lr[:4]),
        "swap_lr_sha256": hashlib.sha256(np.ascontiguousarray(rr.swap_lr, dtype=np.float64).tobytes()).hexdigest(),
        "swap_mean": _hex(rr.swap_mean),
        "logsum_fwd": _hex(rr.logsum_fwd),
        "schedule": _hex(pg.tempering_parameters(pt.shared.tempering)),
        "stepping_stone": float(pg.stepping_stone(pt)).hex(),
        "n_round_trips": int(rr.n_round_trips),
        "n_ref_equiv_evals": int(rr.n_ref_equiv_evals),
        "expl_n_steps": [int(v) for v in rr.expl_n_steps],
    }
