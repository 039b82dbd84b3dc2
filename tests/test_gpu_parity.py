"""GPU parity tests proper: the CUDA engine, called through the C ABI, against the
CPU oracle on the same seeded inputs.  Bit-exact for the swap index permutation,
the uniforms, the accept decisions and every integer statistic; log-densities /
log-ratios are compared bit-exactly too (the two sides implement one arithmetic
spec), which is stricter than the 1e-10 relative tolerance north_star asks for —
the tolerance assertion is kept next to the exact one so a future relaxation of
the device arithmetic (e.g. FMA contraction) has a stated bar."""
import numpy as np
import pytest

import pigeons_jl_b200 as pg

pytestmark = pytest.mark.gpu

RTOL = 1e-10   # north_star: fp64 log-densities within 1e-10 relative


def run_pt(lib, **kw):
    kw.setdefault("record", [pg.index_process, pg.swap_trace, pg.traces])
    pt = pg.pigeons(engine_lib=lib, **kw)
    rr = pt.reduced_recorders
    state = pt.engine.get_state()
    out = dict(rr=rr, schedule=pg.tempering_parameters(pt.shared.tempering).copy(), logz=pg.stepping_stone(pt), state=state,
               explorer=pt.shared.explorer, log=pt.round_log)
    pt.close()
    return out


def assert_same(g, c, label=""):
    rg, rc = g["rr"], c["rr"]
    assert np.array_equal(rg.index_process, rc.index_process), f"{label}: swap index permutation differs"
    assert np.array_equal(rg.swap_accept, rc.swap_accept), f"{label}: accept decisions differ"
    assert np.array_equal(rg.swap_u, rc.swap_u), f"{label}: uniforms differ"
    np.testing.assert_allclose(rg.swap_lr, rc.swap_lr, rtol=RTOL, atol=0, err_msg=f"{label}: log ratios")
    assert np.array_equal(rg.swap_lr, rc.swap_lr), f"{label}: log ratios not bit-identical"
    for k in ("swap_n", "expl_acc_n", "expl_n_steps", "am_n", "rev_n"):
        assert np.array_equal(getattr(rg, k), getattr(rc, k)), f"{label}: {k}"
    for k in ("swap_mean", "logsum_fwd", "logsum_bwd", "expl_acc_mean", "am_mean", "rev_mean", "online_mean", "online_var"):
        np.testing.assert_allclose(getattr(rg, k), getattr(rc, k), rtol=RTOL, atol=0, err_msg=f"{label}: {k}")
        assert np.array_equal(getattr(rg, k), getattr(rc, k)), f"{label}: {k} not bit-identical"
    assert rg.n_round_trips == rc.n_round_trips and rg.n_tempered_restarts == rc.n_tempered_restarts
    assert rg.n_ref_equiv_evals == rc.n_ref_equiv_evals, f"{label}: reference-equivalent eval count"
    if rg.target_trace is not None and rg.target_trace.size:
        assert np.array_equal(rg.target_trace, rc.target_trace), f"{label}: target-chain samples differ"
    assert np.array_equal(g["schedule"], c["schedule"]), f"{label}: adapted schedule differs"
    assert g["logz"] == c["logz"] or (np.isnan(g["logz"]) and np.isnan(c["logz"])), f"{label}: stepping stone differs"
    for k in ("x", "replica_index", "rng_counter", "round_trip_state"):
        assert np.array_equal(g["state"][k], c["state"][k]), f"{label}: final replica {k}"


CASES = {
    "c1_toy_slice": dict(target=pg.toy_mvn_target(2), explorer=pg.SliceSampler(), n_chains=10, n_rounds=10, seed=1),
    "toy_default_explorer": dict(target=pg.toy_mvn_target(3), n_chains=7, n_rounds=8, seed=3),
    "toy10_automala": dict(target=pg.toy_mvn_target(10), explorer=pg.AutoMALA(), n_chains=6, n_rounds=8, seed=2),
    "toy40_slice_2cpl": dict(target=pg.toy_mvn_target(40), explorer=pg.SliceSampler(), n_chains=5, n_rounds=5, seed=5),
    "toy100_automala_4cpl": dict(target=pg.toy_mvn_target(100), explorer=pg.AutoMALA(), n_chains=5, n_rounds=6, seed=6),
    "funnel32_automala": dict(target=pg.Funnel(32), explorer=pg.AutoMALA(), n_chains=24, n_rounds=7, seed=1),
    "funnel8_slice": dict(target=pg.Funnel(8), explorer=pg.SliceSampler(), n_chains=9, n_rounds=6, seed=4),
    "funnel_identity_precond": dict(target=pg.Funnel(16), explorer=pg.AutoMALA(preconditioner=pg.IdentityPreconditioner()),
                                    n_chains=8, n_rounds=6, seed=9),
    "funnel_diag_precond": dict(target=pg.Funnel(16), explorer=pg.AutoMALA(preconditioner=pg.DiagonalPreconditioner()),
                                n_chains=8, n_rounds=6, seed=10),
    "gmm128_automala": dict(target=pg.eight_mode_mixture(128, 8.0), explorer=pg.AutoMALA(), n_chains=12, n_rounds=5, seed=1),
    "gmm6_slice": dict(target=pg.eight_mode_mixture(6, 3.0), explorer=pg.SliceSampler(), n_chains=8, n_rounds=5, seed=2),
    "gmm2_two_modes": dict(target=pg.GaussianMixture(means=[[-8.0, -8.0], [8.0, 8.0]], reference_sigma=8.0),
                           explorer=pg.AutoMALA(), n_chains=5, n_rounds=7, seed=3),
    "toy10_mala": dict(target=pg.toy_mvn_target(10), explorer=pg.MALA(step_size=0.3), n_chains=6, n_rounds=7, seed=11),
    "funnel16_mala": dict(target=pg.Funnel(16), explorer=pg.MALA(step_size=0.2), n_chains=8, n_rounds=6, seed=12),
    "gmm70_mala_4cpl": dict(target=pg.eight_mode_mixture(70, 4.0), explorer=pg.MALA(step_size=0.5), n_chains=6, n_rounds=5, seed=13),
    "logreg24_automala": dict(target=pg.synthetic_logistic_regression(300, 24), explorer=pg.AutoMALA(), n_chains=6, n_rounds=5, seed=1),
    "logreg150_mala": dict(target=pg.synthetic_logistic_regression(4500, 150, seed=3), explorer=pg.MALA(step_size=0.05), n_chains=5,
                           n_rounds=4, seed=2),
    "ising5": dict(target=pg.IsingLogPotential(1.0, 5), n_chains=10, n_rounds=7, seed=1),
    "ising32": dict(target=pg.IsingLogPotential(0.44, 32), n_chains=6, n_rounds=3, seed=2),
    "test_swapper": dict(target=pg.TestSwapper(0.6), n_chains=9, n_rounds=8, seed=7, record=[pg.index_process, pg.swap_trace]),
    "toy3_compose_slice_automala": dict(target=pg.toy_mvn_target(3), explorer=pg.Compose(pg.SliceSampler(), pg.AutoMALA()),
                                        n_chains=4, n_rounds=8, seed=1),
    "gmm70_compose_4cpl": dict(target=pg.eight_mode_mixture(70, 4.0), explorer=pg.Compose(pg.SliceSampler(), pg.AutoMALA()),
                               n_chains=5, n_rounds=4, seed=2),
    "funnel8_mix_automala": dict(target=pg.Funnel(8), n_chains=6, n_rounds=7, seed=1,
                                 explorer=pg.Mix(pg.AutoMALA(preconditioner=pg.IdentityPreconditioner(), base_n_refresh=1),
                                                 pg.AutoMALA(preconditioner=pg.MixDiagonalPreconditioner(0.0, 0.0), base_n_refresh=1),
                                                 pg.AutoMALA(preconditioner=pg.DiagonalPreconditioner(), base_n_refresh=1))),
    "toy40_mix_two_step_sizes": dict(target=pg.toy_mvn_target(40), n_chains=5, n_rounds=6, seed=3,
                                     explorer=pg.Mix(pg.AutoMALA(step_size=0.5), pg.AutoMALA(step_size=2.0, base_n_refresh=2))),
    "mixed_bool_int_float_slice": dict(target=pg.MixedProduct(n_bool=3, n_int=2, n_float=2), n_chains=7, n_rounds=8, seed=1),
    "mixed40_slice_2cpl": dict(target=pg.MixedProduct(n_bool=20, n_int=10, n_float=10, binomial_n=6, p1=0.8, q1=0.2),
                               explorer=pg.SliceSampler(w=4.0, n_passes=2), n_chains=5, n_rounds=5, seed=2),
    "mixed_int_only": dict(target=pg.MixedProduct(n_bool=0, n_int=5, n_float=0, binomial_n=30, q0=0.5, q1=0.1),
                           n_chains=6, n_rounds=7, seed=3),
    # general Compose / Mix programs (Compose.jl:16-19, Mix.jl:20-21): any order, any of ToyExplorer / SliceSampler / MALA / AutoMALA
    "funnel8_compose_automala_then_slice": dict(target=pg.Funnel(8), explorer=pg.Compose(pg.AutoMALA(base_n_refresh=1), pg.SliceSampler(n_passes=1)),
                                                n_chains=6, n_rounds=6, seed=4),
    "toy5_compose_mala_slice_automala": dict(target=pg.toy_mvn_target(5), n_chains=5, n_rounds=6, seed=5,
                                             explorer=pg.Compose(pg.MALA(step_size=0.2, base_n_refresh=1), pg.SliceSampler(n_passes=1),
                                                                 pg.AutoMALA(base_n_refresh=1))),
    "gmm6_mix_slice_automala_mala": dict(target=pg.eight_mode_mixture(6, 3.0), n_chains=6, n_rounds=6, seed=6,
                                         explorer=pg.Mix(pg.SliceSampler(n_passes=1), pg.AutoMALA(base_n_refresh=1),
                                                         pg.MALA(step_size=0.3, base_n_refresh=2))),
    "toy3_mix_toy_slice": dict(target=pg.toy_mvn_target(3), n_chains=5, n_rounds=7, seed=7,
                               explorer=pg.Mix(pg.ToyExplorer(), pg.SliceSampler())),
    "toy70_compose_slice_mala_4cpl": dict(target=pg.toy_mvn_target(70), n_chains=4, n_rounds=4, seed=8,
                                          explorer=pg.Compose(pg.SliceSampler(n_passes=1), pg.MALA(step_size=0.1, base_n_refresh=1))),
    # two legs (StabilizedPT.jl, VariationalDEO.jl) and the Gaussian variational reference (GaussianReference.jl)
    "two_legs_test_swapper": dict(target=pg.TestSwapper(0.5), n_chains=5, n_chains_variational=5, n_rounds=9, seed=1,
                                  record=[pg.index_process, pg.swap_trace]),
    "two_legs_toy_slice_no_variational": dict(target=pg.toy_mvn_target(3), explorer=pg.SliceSampler(), n_chains=4, n_chains_variational=5,
                                              n_rounds=7, seed=2),
    "two_legs_funnel8_slice_gaussian": dict(target=pg.Funnel(8), explorer=pg.SliceSampler(), n_chains=5, n_chains_variational=4,
                                            variational=pg.GaussianReference(first_tuning_round=3), n_rounds=7, seed=3),
    "two_legs_gmm6_automala_gaussian": dict(target=pg.eight_mode_mixture(6, 3.0), explorer=pg.AutoMALA(), n_chains=5,
                                            n_chains_variational=5, variational=pg.GaussianReference(first_tuning_round=2),
                                            n_rounds=7, seed=4),
    "two_legs_gmm128_automala_gaussian_4cpl": dict(target=pg.eight_mode_mixture(128, 8.0), explorer=pg.AutoMALA(), n_chains=4,
                                                   n_chains_variational=4, variational=pg.GaussianReference(first_tuning_round=2),
                                                   n_rounds=5, seed=5),
    "two_legs_funnel40_mala_gaussian_2cpl": dict(target=pg.Funnel(40), explorer=pg.MALA(step_size=0.1), n_chains=4, n_chains_variational=4,
                                                 variational=pg.GaussianReference(first_tuning_round=3), n_rounds=6, seed=6),
    "one_leg_variational_gmm1": dict(target=pg.GaussianMixture(means=[[1.0]], reference_sigma=1.0), explorer=pg.SliceSampler(),
                                     n_chains=0, n_chains_variational=6, variational=pg.GaussianReference(first_tuning_round=2),
                                     n_rounds=8, seed=7),
    # test/test_DistributionLogPotential.jl:23-31: N(3,1) against a fixed reference N(-3,1) (never re-fitted)
    "dlp_univariate_fixed_gaussian_reference": dict(
        target=pg.GaussianMixture(means=[[3.0]], reference_sigma=1.0), explorer=pg.SliceSampler(), n_chains=0, n_chains_variational=8,
        variational=pg.GaussianReference(first_tuning_round=10 ** 9, mean=np.array([-3.0]), standard_deviation=np.array([1.0])),
        n_rounds=8, seed=1),
    # the reference's own test targets: test/test_DistributionLogPotential.jl:7-21 and test/test_two_legs.jl
    "unid_dlp_multivariate": dict(target=pg.UnidentifiableProduct(100, 50), n_chains=4, n_rounds=10, seed=1),
    "unid_two_legs_100000": dict(target=pg.UnidentifiableProduct(100000), n_chains=8, n_chains_variational=7,
                                 variational=pg.GaussianReference(first_tuning_round=99), n_rounds=9, seed=1),
    "unid_two_legs_gaussian": dict(target=pg.UnidentifiableProduct(100000), n_chains=8, n_chains_variational=7,
                                   variational=pg.GaussianReference(first_tuning_round=4), n_rounds=9, seed=2),
    "two_legs_toy_mala_test_mala_jl": dict(target=pg.toy_mvn_target(2), n_chains=2, explorer=pg.MALA(), n_chains_variational=4,
                                           n_rounds=9, seed=1),
    "two_legs_ising5": dict(target=pg.IsingLogPotential(0.8, 5), n_chains=5, n_chains_variational=4, n_rounds=7, seed=8),
    "two_legs_never_activated": dict(target=pg.Funnel(8), explorer=pg.AutoMALA(), n_chains=4, n_chains_variational=4,
                                     variational=pg.GaussianReference(first_tuning_round=99), n_rounds=6, seed=9),
    "two_legs_gmm128_automala_gaussian_n96": dict(target=pg.eight_mode_mixture(128, 8.0), explorer=pg.AutoMALA(), n_chains=48,
                                                  n_chains_variational=48, variational=pg.GaussianReference(first_tuning_round=3),
                                                  n_rounds=6, seed=12),
    "single_chain": dict(target=pg.toy_mvn_target(4), explorer=pg.SliceSampler(), n_chains=1, n_rounds=5, seed=1),
    "two_chains": dict(target=pg.toy_mvn_target(2), explorer=pg.AutoMALA(), n_chains=2, n_rounds=6, seed=8),
}


@pytest.mark.parametrize("name", list(CASES))
def test_multi_round_parity(name, gpu_lib, oracle_lib):
    """Default recorder order: per-replica recorders merged by the reference's tree (recorders.jl:88-120)."""
    kw = CASES[name]
    assert_same(run_pt(gpu_lib, **kw), run_pt(oracle_lib, **kw), name)


@pytest.mark.parametrize("name", ["c1_toy_slice", "funnel32_automala", "gmm128_automala", "logreg24_automala", "ising5",
                                  "toy3_compose_slice_automala", "funnel8_mix_automala", "two_chains"])
def test_multi_round_parity_per_chain_recorders(name, gpu_lib, oracle_lib):
    """recorder_order = PGN_RECORDERS_PER_CHAIN: one accumulator per chain / pair, fitted in scan order."""
    kw = dict(CASES[name], recorder_order=1)
    assert_same(run_pt(gpu_lib, **kw), run_pt(oracle_lib, **kw), name + "/per_chain")


@pytest.mark.parametrize("name", ["toy10_automala", "toy1000_automala_mem"])
def test_recorder_orders_agree_to_rounding(name, gpu_lib):
    """The two recorder orders describe the same statistics: after ONE round (same trajectory: adaptation has not acted yet)
    all counts are equal and the means agree to rounding."""
    kw = dict(target=pg.toy_mvn_target(10 if name == "toy10_automala" else 1000), explorer=pg.AutoMALA(), n_chains=6, seed=2)
    a = run_pt(gpu_lib, n_rounds=1, **kw)["rr"]
    b = run_pt(gpu_lib, n_rounds=1, recorder_order=1, **kw)["rr"]
    for k in ("swap_n", "expl_acc_n", "am_n", "rev_n", "expl_n_steps"):
        assert np.array_equal(getattr(a, k), getattr(b, k)), k
    for k in ("swap_mean", "logsum_fwd", "logsum_bwd", "expl_acc_mean", "am_mean", "rev_mean", "online_mean", "online_var"):
        np.testing.assert_allclose(getattr(a, k), getattr(b, k), rtol=1e-12, atol=1e-300, err_msg=k)


def test_round_trips_known_answer(gpu_lib):
    """test/test_round_trips.jl:1-14 on the device: exactly 13 round trips."""
    pt = pg.pigeons(target=pg.TestSwapper(1.0), record=[pg.round_trip], n_chains=4, n_rounds=5, engine_lib=gpu_lib)
    assert pg.n_round_trips(pt) == 13
    pt.close()


def _points(rng, n, d, scale=2.0):
    return rng.normal(0.0, scale, size=(n, d))


@pytest.mark.parametrize("target", [pg.toy_mvn_target(2), pg.toy_mvn_target(77), pg.Funnel(32), pg.Funnel(5),
                                    pg.eight_mode_mixture(128, 8.0), pg.eight_mode_mixture(40, 2.0),
                                    pg.synthetic_logistic_regression(300, 24), pg.synthetic_logistic_regression(4500, 150, seed=3)])
def test_log_potential_and_gradient_entry_points(target, gpu_lib, oracle_lib):
    """pgn_log_potential / pgn_logdensity_and_gradient against the oracle, incl. beta in {0, 1}."""
    rng = np.random.default_rng(0)
    cfg = target.engine_config()
    eg = pg.Engine(gpu_lib, n_chains=4, seed=1, **cfg)
    eo = pg.Engine(oracle_lib, n_chains=4, seed=1, **cfg)
    x = _points(rng, 64, target.dim, 0.7 if isinstance(target, pg.LogisticRegression) else 2.0)
    beta = rng.uniform(0, 1, 64)
    beta[:8] = 0.0
    beta[8:16] = 1.0
    lg, lo = eg.log_potential(x, beta), eo.log_potential(x, beta)
    np.testing.assert_allclose(lg, lo, rtol=RTOL, atol=0)
    assert np.array_equal(lg, lo)
    (ldg, gg), (ldo, go) = eg.logdensity_and_gradient(x, beta), eo.logdensity_and_gradient(x, beta)
    np.testing.assert_allclose(ldg, ldo, rtol=RTOL, atol=0)
    np.testing.assert_allclose(gg, go, rtol=RTOL, atol=1e-300)
    assert np.array_equal(ldg, ldo) and np.array_equal(gg, go)
    # analytic gradient vs central finite differences of the device density itself
    h = 1e-6
    for j in (0, target.dim - 1):
        xp, xm = x.copy(), x.copy()
        xp[:, j] += h
        xm[:, j] -= h
        fd = (eg.logdensity_and_gradient(xp, beta)[0] - eg.logdensity_and_gradient(xm, beta)[0]) / (2 * h)
        np.testing.assert_allclose(gg[:, j], fd, rtol=2e-5, atol=2e-5 * max(1.0, float(np.max(np.abs(gg)))))
    eg.close(); eo.close()


def test_device_numerics_bit_identical(gpu_lib, oracle_lib):
    """The device elementary functions and the Philox draws equal the oracle's, bit for bit."""
    rng = np.random.default_rng(1)
    xs = {
        0: np.concatenate([rng.uniform(-745, 709, 20000), rng.normal(0, 1, 20000), [0.0, -0.0, np.inf, -np.inf, np.nan, 710.0, -746.0]]),
        1: np.concatenate([np.exp(rng.uniform(-700, 700, 20000)), rng.uniform(0, 2, 20000), [0.0, 1.0, np.inf, -1.0, np.nan, 5e-324]]),
        2: rng.uniform(0, 2, 40000),
        3: np.arange(0, 40000, dtype=np.float64),
        4: np.arange(0, 40000, dtype=np.float64),
        5: np.arange(0, 40000, dtype=np.float64),
        6: rng.normal(0, 30, 40000),
        7: np.concatenate([-rng.uniform(0, 1, 20000), rng.uniform(0, 5, 10000), rng.normal(0, 1e-9, 10000), [0.0, -1.0, -0.5, 1e-300, -1e-17]]),
    }
    for op, v in xs.items():
        a = gpu_lib.test_math(op, v, seed=12345678901, replica_index=17)
        b = oracle_lib.test_math(op, v, seed=12345678901, replica_index=17)
        assert np.array_equal(a, b, equal_nan=True), f"op {op}: device and oracle differ"


def test_errors_are_reported_not_swallowed(gpu_lib):
    """NaN log ratio -> rc != 0 with a message (src/log_potentials/log_potentials.jl:47-49)."""
    t = pg.Funnel(4)
    e = pg.Engine(gpu_lib, n_chains=3, seed=1, **t.engine_config())
    e.init_replicas()
    e.set_explorer(**pg.SliceSampler().engine_params(4))
    x = np.zeros((3, 4))
    x[1, 0] = np.inf     # y = +inf: ref = -inf, target = nan
    e.set_state(x=x)
    with pytest.raises(pg.EngineError):
        e.run_round(2)
    e.close()


def test_unsupported_configuration_raises(gpu_lib):
    """Device targets are a closed family: a 9-component mixture has no device implementation."""
    nine = pg.GaussianMixture(means=np.arange(18.0).reshape(9, 2), reference_sigma=3.0)
    with pytest.raises(pg.EngineError):
        pg.Engine(gpu_lib, n_chains=4, seed=1, **nine.engine_config())


_ORACLE_CACHE = {}


def oracle_result(name, oracle_lib):
    if name not in _ORACLE_CACHE:
        _ORACLE_CACHE[name] = run_pt(oracle_lib, **CASES[name])
    return _ORACLE_CACHE[name]


@pytest.mark.parametrize("team", [1, 2, 3, 4, 5, 8])
@pytest.mark.parametrize("name", ["funnel32_automala", "toy100_automala_4cpl", "gmm2_two_modes", "two_chains", "funnel_diag_precond",
                                  "toy3_compose_slice_automala", "funnel8_mix_automala", "toy5_compose_mala_slice_automala",
                                  "gmm6_mix_slice_automala_mala"])
def test_automala_team_width_parity(name, team, gpu_lib, oracle_lib, monkeypatch):
    """PGN_TEAM=W: the autoMALA step-size search evaluated W candidate steps at a time by a team of W
    warps per chain gives the reference's sequential result bit for bit, for every team width."""
    monkeypatch.setenv("PGN_TEAM", str(team))
    assert_same(run_pt(gpu_lib, **CASES[name]), oracle_result(name, oracle_lib), f"{name}/team{team}")


def test_ising_antiferromagnetic_and_large_ladder(gpu_lib, oracle_lib):
    """The tabulated Metropolis ratios (one table per chain and round) against the per-site
    evaluation of the oracle: negative coupling (the lowering moves change sign) and a 40-chain ladder."""
    for kw in (dict(target=pg.IsingLogPotential(-0.7, 8), n_chains=7, n_rounds=6, seed=3),
               dict(target=pg.IsingLogPotential(0.4406867935097715, 32), n_chains=40, n_rounds=4, seed=4),
               dict(target=pg.IsingLogPotential(1e-9, 6), n_chains=4, n_rounds=5, seed=5)):
        assert_same(run_pt(gpu_lib, **kw), run_pt(oracle_lib, **kw), f"ising {kw['target']}")


FULL_SIZE = {
    # BASELINE.json configs 2-4 at their full per-GPU width (rounds kept short so the oracle finishes in seconds)
    "c2_funnel32_automala_256": dict(target=pg.Funnel(32), explorer=pg.AutoMALA(), n_chains=256, n_rounds=5, seed=1,
                                     record=[pg.index_process, pg.swap_trace]),
    "c3_gmm128_automala_1024": dict(target=pg.eight_mode_mixture(128, 8.0), explorer=pg.AutoMALA(), n_chains=1024, n_rounds=3, seed=1,
                                    record=[pg.index_process, pg.swap_trace]),
    "c4_ising32_512": dict(target=pg.IsingLogPotential(0.4406867935097715, 32), n_chains=512, n_rounds=4, seed=1,
                           record=[pg.index_process, pg.swap_trace]),
}


def check_scan_invariants(rr, n_chains):
    """Size-independent properties of a PT scan trace (src/swap/swap.jl:6-39, OddEven.jl:23-31): every scan's
    index process is a permutation; only DEO neighbours of the scan's parity exchange replicas; a pair's two
    sides agree; and the decision is u_lower < min(1, exp(lr_lower + lr_upper))."""
    ip, acc, lr, u = rr.index_process, rr.swap_accept, rr.swap_lr, rr.swap_u
    n_scans = ip.shape[0]
    assert ip.shape == (n_scans, n_chains)
    ident = np.arange(n_chains)
    for s in range(n_scans):
        assert np.array_equal(np.sort(ip[s]), np.sort(ip[0])), "index process is not a permutation"
        even = ((s + 1) % 2 == 0)                      # scans are 1-based inside a round; rounds restart the parity
        partner = np.where(((ident + 1) % 2 == 0) == even, ident + 1, ident - 1)
        partner = np.clip(partner, 0, n_chains - 1)
        lo = ident[(partner == ident + 1)]
        assert np.array_equal(acc[s][lo], acc[s][lo + 1]), "the two sides of a pair disagree"
        alone = ident[partner == ident]
        assert not acc[s][alone].any()
        a = np.minimum(1.0, np.exp(lr[s][lo] + lr[s][lo + 1]))
        decided = u[s][lo] < a
        clear = np.abs(u[s][lo] - a) > 1e-12           # numpy's exp is not the spec's exp_: skip razor-edge cases
        assert np.array_equal(decided[clear], acc[s][lo][clear].astype(bool)), "accept decision is not u < min(1, exp(sum lr))"
    return True


@pytest.mark.parametrize("name", list(FULL_SIZE))
def test_full_size_configs(name, gpu_lib, oracle_lib):
    """BASELINE configs at full width: bit-exact against the oracle over the first rounds, plus the
    size-independent scan invariants on the device trace of the last round."""
    kw = FULL_SIZE[name]
    g, c = run_pt(gpu_lib, **kw), run_pt(oracle_lib, **kw)
    assert_same(g, c, name)
    rr = g["rr"]
    # rounds restart the DEO parity: check the last round's block of scans
    n_last = 2 ** kw["n_rounds"]
    last = type("R", (), dict(index_process=rr.index_process[-n_last:], swap_accept=rr.swap_accept[-n_last:],
                              swap_lr=rr.swap_lr[-n_last:], swap_u=rr.swap_u[-n_last:]))
    check_scan_invariants(last, kw["n_chains"])


def test_known_answers_on_the_device(gpu_lib):
    """The reference's tolerance-based known answers, run on the CUDA path itself: stepping stone of
    toy_mvn_target(10) (test/test_stepping_stone.jl:15-28), cumulative barrier of toy_mvn_target(2)
    (test/test_cumulative_barrier.jl:1-11), Ising 5x5 log Z (examples/custom-sampler.jl:4-5)."""
    pt = pg.pigeons(target=pg.toy_mvn_target(10), explorer=pg.AutoMALA(), n_chains=6, n_rounds=12, seed=1, engine_lib=gpu_lib)
    truth = 0.5 * 10 * (np.log(1.0) - np.log(10.0))
    assert abs(pg.stepping_stone(pt) - truth) < 0.2
    pt.close()
    pt = pg.pigeons(target=pg.IsingLogPotential(1.0, 5), n_chains=10, n_rounds=12, seed=1, engine_lib=gpu_lib)
    assert abs(pg.stepping_stone(pt) - 33.37317482430507) < 0.15
    pt.close()


MEM_FORCED = ["c1_toy_slice", "toy_default_explorer", "toy10_automala", "toy40_slice_2cpl", "toy100_automala_4cpl",
              "funnel32_automala", "funnel8_slice", "funnel_diag_precond", "gmm128_automala", "gmm6_slice",
              "gmm2_two_modes", "toy10_mala", "gmm70_mala_4cpl", "single_chain", "two_chains"]


@pytest.mark.parametrize("name", MEM_FORCED)
def test_memory_resident_kernel_parity(name, gpu_lib, oracle_lib, monkeypatch):
    """PGN_FORCE_MEM=1: the memory-resident scan kernel (any d, any number of chains) reproduces the
    oracle bit for bit on the same cases as the register-resident kernel."""
    monkeypatch.setenv("PGN_FORCE_MEM", "1")
    kw = CASES[name]
    assert_same(run_pt(gpu_lib, **kw), run_pt(oracle_lib, **kw), name + "/mem")


BIG_CASES = {
    # dimensions beyond the register-resident limit (d > 128)
    "toy1000_automala": dict(target=pg.toy_mvn_target(1000), explorer=pg.AutoMALA(), n_chains=4, n_rounds=5, seed=1),
    "funnel300_slice": dict(target=pg.Funnel(300), explorer=pg.SliceSampler(), n_chains=3, n_rounds=3, seed=2),
    "gmm200_automala": dict(target=pg.eight_mode_mixture(200, 4.0), explorer=pg.AutoMALA(), n_chains=5, n_rounds=4, seed=3),
    "funnel500_mala": dict(target=pg.Funnel(500), explorer=pg.MALA(step_size=0.05), n_chains=4, n_rounds=4, seed=4),
    # more chains than fit co-resident: one warp serves several chains
    "toy2_slice_6000_chains": dict(target=pg.toy_mvn_target(2), explorer=pg.SliceSampler(), n_chains=6000, n_rounds=3, seed=5,
                                   record=[pg.index_process, pg.swap_trace]),
    "funnel8_automala_5000_chains": dict(target=pg.Funnel(8), explorer=pg.AutoMALA(), n_chains=5000, n_rounds=2, seed=6,
                                         record=[pg.index_process, pg.swap_trace]),
}


@pytest.mark.parametrize("name", list(BIG_CASES))
def test_large_dimension_and_many_chains(name, gpu_lib, oracle_lib):
    kw = BIG_CASES[name]
    assert_same(run_pt(gpu_lib, **kw), run_pt(oracle_lib, **kw), name)


def test_entry_points_large_dimension(gpu_lib, oracle_lib):
    rng = np.random.default_rng(7)
    for target in (pg.toy_mvn_target(1000), pg.Funnel(300), pg.eight_mode_mixture(200, 4.0)):
        cfg = target.engine_config()
        eg = pg.Engine(gpu_lib, n_chains=4, seed=1, **cfg)
        eo = pg.Engine(oracle_lib, n_chains=4, seed=1, **cfg)
        x = rng.normal(0, 1, (9, target.dim))
        beta = np.array([0.0, 1.0, 0.5, 0.1, 0.9, 0.3, 0.7, 0.2, 0.8])
        assert np.array_equal(eg.log_potential(x, beta), eo.log_potential(x, beta))
        (ldg, gg), (ldo, go) = eg.logdensity_and_gradient(x, beta), eo.logdensity_and_gradient(x, beta)
        assert np.array_equal(ldg, ldo) and np.array_equal(gg, go)
        eg.close(); eo.close()


def test_dmma_is_a_sequential_fma_chain(gpu_lib):
    """mma.sync.m8n8k4.f64 accumulates its four products as an fma chain in ascending k starting
    from C (exactly-rounded reference via rational arithmetic).  This is what makes the
    tensor-core GEMM a bit-identical replacement for the SIMT one."""
    from fractions import Fraction
    rng = np.random.default_rng(0)
    n = 24
    scale = lambda shape: rng.normal(0, 1, shape) * np.exp(rng.uniform(-8, 8, shape))   # noqa: E731
    a, b, c = scale((n, 8, 4)), scale((n, 4, 8)), scale((n, 8, 8))
    d = gpu_lib.test_dmma(a, b, c)
    for t in range(n):
        for i in range(8):
            for j in range(8):
                acc = float(c[t, i, j])
                for k in range(4):
                    acc = float(Fraction(float(a[t, i, k])) * Fraction(float(b[t, k, j])) + Fraction(acc))
                assert acc == d[t, i, j]


@pytest.mark.parametrize("name", ["logreg24_automala", "logreg150_mala"])
def test_logreg_parity_with_simt_gemm(name, gpu_lib, oracle_lib, monkeypatch):
    """PGN_GEMM=simt: the DFMA GEMM (the default is the FP64 tensor-core kernel) gives the same bits."""
    monkeypatch.setenv("PGN_GEMM", "simt")
    kw = CASES[name]
    assert_same(run_pt(gpu_lib, **kw), run_pt(oracle_lib, **kw), name + "/simt")


@pytest.mark.parametrize("target", [pg.toy_mvn_target(10), pg.Funnel(8), pg.eight_mode_mixture(6, 3.0)])
def test_leapfrog_involution_on_the_device(target, gpu_lib):
    """test/test_auto_mala.jl:51-85 with the gradients of the CUDA entry point."""
    from test_oracle_known_answers import leapfrog_involution_error
    e = pg.Engine(gpu_lib, n_chains=2, seed=1, **target.engine_config())
    for beta in (0.0, 0.37, 1.0):
        dx, dp, moved = leapfrog_involution_error(e, target, beta)
        assert moved > 1e-3 and dx < 1e-9 and dp < 1e-9
    e.close()


@pytest.mark.parametrize("target", [pg.toy_mvn_target(10), pg.Funnel(8), pg.eight_mode_mixture(6, 3.0),
                                    pg.toy_mvn_target(100), pg.eight_mode_mixture(128, 8.0)])
def test_device_integrator_is_an_involution_and_matches_the_oracle(target, gpu_lib, oracle_lib):
    """test/test_auto_mala.jl:51-85 on the DEVICE's own leapfrog (VecChain::run_trial, the code autoMALA and MALA run):
    forward n steps, flip the momentum, forward n steps, flip -> back at the start; and every intermediate result is the
    oracle's `hamiltonian_dynamics!` bit for bit."""
    rng = np.random.default_rng(11)
    d = target.dim
    n_pts = 9
    x0 = rng.standard_normal((n_pts, d)) * 0.7
    p0 = rng.standard_normal((n_pts, d))
    betas = np.linspace(0.0, 1.0, n_pts)
    dev = pg.Engine(gpu_lib, n_chains=2, seed=1, **target.engine_config())
    orc = pg.Engine(oracle_lib, n_chains=2, seed=1, **target.engine_config())
    eps, n = 0.02, 40
    x1, p1 = dev.hamiltonian_dynamics(x0, p0, betas, eps, n)
    ox1, op1 = orc.hamiltonian_dynamics(x0, p0, betas, eps, n)
    assert np.array_equal(x1, ox1) and np.array_equal(p1, op1)
    assert np.abs(x1 - x0).max() > 1e-3                       # it moved
    x2, p2 = dev.hamiltonian_dynamics(x1, -p1, betas, eps, n)
    ox2, op2 = orc.hamiltonian_dynamics(ox1, -op1, betas, eps, n)
    assert np.array_equal(x2, ox2) and np.array_equal(p2, op2)
    assert np.abs(x2 - x0).max() < 1e-9 and np.abs(-p2 - p0).max() < 1e-9
    # the reference's own variant (test/test_auto_mala.jl:51-85): a non-trivial diagonal preconditioner, and the "flip step"
    # form (forward with +eps, back with -eps) next to the "flip momentum" form above
    cond = rng.uniform(0.5, 2.5, d)
    xc, pc = dev.hamiltonian_dynamics(x0, p0, betas, 0.1 * eps * 5, n, diag_precond=cond)
    oxc, opc = orc.hamiltonian_dynamics(x0, p0, betas, 0.1 * eps * 5, n, diag_precond=cond)
    assert np.array_equal(xc, oxc) and np.array_equal(pc, opc)
    assert np.abs(xc - x0).max() > 1e-3 and not np.array_equal(xc, x1)
    xb, pb = dev.hamiltonian_dynamics(xc, pc, betas, -0.1 * eps * 5, n, diag_precond=cond)
    assert np.abs(xb - x0).max() < 1e-9 and np.abs(pb - p0).max() < 1e-9
    # zero steps is the identity; one step equals the leapfrog written out in numpy on the gradient entry point
    xz, pz = dev.hamiltonian_dynamics(x0, p0, betas, eps, 0)
    assert np.array_equal(xz, x0) and np.array_equal(pz, p0)
    _, g0 = dev.logdensity_and_gradient(x0, betas)
    ph = p0 + (eps / 2) * g0
    xs = x0 + eps * ph
    _, g1 = dev.logdensity_and_gradient(xs, betas)
    xd, pd = dev.hamiltonian_dynamics(x0, p0, betas, eps, 1)
    assert np.array_equal(xd, xs) and np.array_equal(pd, ph + (eps / 2) * g1)
    dev.close(); orc.close()


def test_variational_entry_points_match_the_oracle(gpu_lib, oracle_lib):
    """With a GaussianReference installed the parity entry points evaluate the variational leg's path: log potential,
    logdensity + gradient and the integrator, device vs oracle bit for bit; and the reference's own 'Manual diff check'
    (test/test_variational.jl:71-84): the hand-written gradient equals the derivative of gaussian_logdensity."""
    rng = np.random.default_rng(5)
    for target in (pg.Funnel(8), pg.eight_mode_mixture(70, 4.0)):
        d = target.dim
        mean, sd = rng.standard_normal(d), rng.uniform(0.3, 2.0, d)
        x = rng.standard_normal((7, d))
        betas = np.array([0.0, 0.1, 0.37, 0.5, 0.9, 1.0, 0.0])
        res = []
        for lib in (gpu_lib, oracle_lib):
            e = pg.Engine(lib, n_chains=4, n_chains_variational=2, seed=1, **target.engine_config())
            e.set_variational(mean, sd)
            lp = e.log_potential(x, betas)
            ld, g = e.logdensity_and_gradient(x, betas)
            hd = e.hamiltonian_dynamics(x, rng.standard_normal((7, d)) * 0 + 0.5, betas, 0.01, 10)
            e.set_variational(None, None)
            lp_fixed = e.log_potential(x, betas)
            e.close()
            res.append((lp, ld, g, hd[0], hd[1], lp_fixed))
        for a, b in zip(*res):
            assert np.array_equal(a, b)
        lp, ld, g = res[0][0], res[0][1], res[0][2]
        ref0 = np.sum(-0.5 * np.log(2.0 * np.pi * sd ** 2) - (x[0] - mean) ** 2 / (2.0 * sd ** 2))      # beta = 0: the reference alone
        np.testing.assert_allclose(lp[0], ref0, rtol=1e-12)
        np.testing.assert_allclose(g[0], -(x[0] - mean) / sd ** 2, rtol=1e-12)
        assert not np.array_equal(res[0][0][:5], res[0][5][:5]) and res[0][0][5] == res[0][5][5]         # beta = 1: the target alone


@pytest.mark.parametrize("name,cap", [("toy100_automala_4cpl", 2), ("funnel32_automala", 7), ("gmm128_automala", 5),
                                      ("toy40_mix_two_step_sizes", 1), ("gmm2_two_modes", 9), ("funnel_diag_precond", 0)])
def test_mixed_teams_parity(name, cap, gpu_lib, oracle_lib, monkeypatch):
    """"Mixed teams": blocks of two warps serve either one chain as a team of two or two chains with one warp each, the
    teams going to the chains that did most work in the previous round.  Forced here on small ladders (PGN_TEAM=1 selects the
    single-warp launch the mixed one replaces; PGN_MIXED_MAX_TEAMS caps the teams so that team blocks, pair blocks and a
    half-empty block all occur): bit-identical to the oracle, like every other team width."""
    monkeypatch.setenv("PGN_TEAM", "1")
    monkeypatch.setenv("PGN_MIXED_TEAMS", "1")
    monkeypatch.setenv("PGN_MIXED_MAX_TEAMS", str(cap))
    kw = CASES[name]
    assert_same(run_pt(gpu_lib, **kw), run_pt(oracle_lib, **kw), name + "/mixed")


def test_mixed_teams_at_c3_width_match_the_oracle(gpu_lib, oracle_lib, monkeypatch):
    """BASELINE config 3 at its full width (1024 chains, GMM d = 128, autoMALA): the single-warp launch leaves 160 warp slots
    free; with mixed teams they go to the 160 chains that worked most.  4 rounds against the oracle, bit for bit."""
    monkeypatch.setenv("PGN_MIXED_TEAMS", "1")
    kw = dict(target=pg.eight_mode_mixture(128, 8.0), explorer=pg.AutoMALA(), n_chains=1024, n_rounds=4, seed=3,
              record=[pg.index_process, pg.swap_trace])
    assert_same(run_pt(gpu_lib, **kw), run_pt(oracle_lib, **kw), "c3_width/mixed")


def test_two_legs_at_c3_width(gpu_lib):
    """BASELINE config 3's ladder as TWO legs of 512 chains (GMM d = 128, autoMALA, GaussianReference from round 3): too large
    for the oracle, so size-independent properties — every scan's index process is a permutation, both partners log the same
    decision, the pair between the two target chains always swaps with log ratio 0, references and targets are visited
    (tempered restarts > 0), the fitted reference is finite, and a second run is bit-identical."""
    kw = dict(target=pg.eight_mode_mixture(128, 8.0), explorer=pg.AutoMALA(), n_chains=512, n_chains_variational=512,
              variational=pg.GaussianReference(first_tuning_round=3), n_rounds=6, seed=1,
              record=[pg.index_process, pg.swap_trace, pg.round_trip], engine_lib=gpu_lib)
    pt = pg.pigeons(**kw)
    rr, n, nv = pt.reduced_recorders, 1024, 512
    ip = rr.index_process
    assert ip.shape == (64, n) and np.array_equal(np.sort(ip, axis=1), np.broadcast_to(np.arange(1, n + 1), ip.shape))
    for s in range(ip.shape[0]):
        even = (s + 1) % 2 == 0
        c = np.arange(1, n + 1)
        partner = np.clip(c + np.where((c % 2 == 0) == even, 1, -1), 1, n)
        assert np.array_equal(rr.swap_accept[s], rr.swap_accept[s][partner - 1])
        if partner[nv - 1] == nv + 1:
            assert rr.swap_accept[s, nv - 1] == 1 and rr.swap_lr[s, nv - 1] == 0.0 and rr.swap_lr[s, nv] == 0.0
    v = pt.inputs.variational
    assert v.mean is not None and np.all(np.isfinite(v.mean)) and np.all(v.standard_deviation > 0)
    assert isinstance(pt.shared.tempering, pg.StabilizedPT) and pg.global_barrier_variational(pt) > 0
    again = pg.pigeons(**kw)
    assert np.array_equal(again.reduced_recorders.index_process, ip)
    assert np.array_equal(again.reduced_recorders.online_mean, rr.online_mean)
    assert np.array_equal(pg.tempering_parameters(again.shared.tempering), pg.tempering_parameters(pt.shared.tempering))
    pt.close(); again.close()


def test_two_legs_refuse_what_they_cannot_run(gpu_lib):
    with pytest.raises(pg.EngineError):      # per-chain recorder order
        pg.Engine(gpu_lib, n_chains=6, n_chains_variational=3, seed=1, recorder_order=1, **pg.Funnel(4).engine_config())
    # several shards: the balanced split 1-3 | 4-6 would separate the targets 3 and 4: chain 4 joins the lower shard
    e = pg.Engine(gpu_lib, n_chains=6, n_chains_variational=3, seed=1, rank=0, world_size=2, **pg.Funnel(4).engine_config())
    assert (e.first_chain, e.n_local) == (1, 4) == pg.shard_layout(6, 2, 3)[0]
    e.close()
    e = pg.Engine(gpu_lib, n_chains=6, n_chains_variational=3, seed=1, rank=1, world_size=2, **pg.Funnel(4).engine_config())
    assert (e.first_chain, e.n_local) == (5, 2) == pg.shard_layout(6, 2, 3)[1]
    e.close()
    with pytest.raises(pg.EngineError):      # ... which must not empty the upper shard
        pg.Engine(gpu_lib, n_chains=4, n_chains_variational=1, seed=1, rank=0, world_size=4, **pg.Funnel(4).engine_config())
    e = pg.Engine(gpu_lib, n_chains=6, n_chains_variational=3, seed=1, **pg.toy_mvn_target(4).engine_config())
    with pytest.raises(pg.EngineError):      # a Gaussian reference needs an InterpolatingPath (FUNNEL, GMM)
        e.set_variational(np.zeros(4), np.ones(4))
    e.close()
    e = pg.Engine(gpu_lib, n_chains=6, n_chains_variational=3, seed=1, **pg.Funnel(4).engine_config())
    with pytest.raises(pg.EngineError):
        e.set_variational(np.zeros(4), np.array([1.0, 0.0, 1.0, 1.0]))
    e.close()


RESUME_CASES = {
    "toy_slice": dict(target=pg.toy_mvn_target(3), explorer=pg.SliceSampler(), n_chains=5, seed=2),
    "funnel_automala_team": dict(target=pg.Funnel(32), explorer=pg.AutoMALA(), n_chains=12, seed=3),
    "gmm_automala": dict(target=pg.eight_mode_mixture(128, 8.0), explorer=pg.AutoMALA(), n_chains=6, seed=4),
    "ising": dict(target=pg.IsingLogPotential(0.6, 5), n_chains=5, seed=4),
    "logreg": dict(target=pg.synthetic_logistic_regression(300, 24), explorer=pg.AutoMALA(), n_chains=5, seed=5),
    "toy300_mem": dict(target=pg.toy_mvn_target(300), explorer=pg.MALA(step_size=0.1), n_chains=4, seed=6),
    "test_swapper": dict(target=pg.TestSwapper(0.7), n_chains=6, seed=7),
    "two_legs_gaussian": dict(target=pg.eight_mode_mixture(6, 3.0), explorer=pg.AutoMALA(), n_chains=4, n_chains_variational=4,
                              variational=pg.GaussianReference(first_tuning_round=2), seed=8),
}


@pytest.mark.parametrize("name", list(RESUME_CASES))
def test_device_resume_is_bit_identical(name, gpu_lib, oracle_lib):
    """test/test_resume.jl on the DEVICE: pgn_get_state after round 4 -> a NEW handle -> pgn_set_state -> rounds 5..6
    gives the run that went straight to round 6 bit for bit, and both equal the oracle's straight run."""
    kw = RESUME_CASES[name]
    rec = [pg.index_process, pg.swap_trace]
    straight = pg.pigeons(engine_lib=gpu_lib, n_rounds=6, record=rec, **kw)
    first = pg.pigeons(engine_lib=gpu_lib, n_rounds=4, record=rec, **kw)
    ckpt = pg.write_checkpoint(first)
    first.close()                                    # the handle is gone: only the checkpoint survives
    resumed = pg.resume(ckpt, pg.Inputs(engine_lib=gpu_lib, n_rounds=6, record=rec, **kw))
    ref = pg.pigeons(engine_lib=oracle_lib, n_rounds=6, record=rec, **kw)
    for other, label in ((resumed, "resumed"), (ref, "oracle")):
        a, b = straight.reduced_recorders, other.reduced_recorders
        for k in ("index_process", "swap_lr", "swap_u", "swap_accept", "swap_mean", "logsum_fwd", "logsum_bwd", "expl_n_steps",
                  "expl_acc_mean", "am_mean", "rev_mean"):
            assert np.array_equal(getattr(a, k), getattr(b, k)), f"{name}/{label}: {k}"
        assert np.array_equal(pg.tempering_parameters(straight.shared.tempering), pg.tempering_parameters(other.shared.tempering))
        assert straight.shared.explorer == other.shared.explorer
        sa, sb = straight.engine.get_state(), other.engine.get_state()
        for k in sa:
            assert np.array_equal(sa[k], sb[k]), f"{name}/{label}: replica {k}"
        assert a.n_round_trips == b.n_round_trips
    straight.close(); resumed.close(); ref.close()


def test_set_state_is_validated_on_the_device_binding(gpu_lib):
    t = pg.toy_mvn_target(3)
    e = pg.Engine(gpu_lib, n_chains=5, seed=1, **t.engine_config())
    e.init_replicas()
    with pytest.raises(ValueError):
        e.set_state(x=np.zeros((4, 3)))
    with pytest.raises(ValueError):
        e.set_state(rng_counter=np.zeros(6, dtype=np.uint64))
    e.close()


# ---- BASELINE config 5 at its full shape (d = 4096, n_data = 65536) ---------------------------------------------------
@pytest.fixture(scope="module")
def c5(oracle_lib):
    """The C5 data set (2 GiB design matrix) and the oracle's answers on three points, computed once."""
    target = pg.synthetic_logistic_regression(65536, 4096)
    rng = np.random.default_rng(5)
    x = rng.normal(0.0, 0.05, size=(3, 4096))
    x[1] *= 10.0                                      # larger |z|: both branches of the Bernoulli terms
    beta = np.array([0.0, 0.37, 1.0])
    eo = pg.Engine(oracle_lib, n_chains=4, seed=1, **target.engine_config())
    out = dict(target=target, x=x, beta=beta, lp=eo.log_potential(x, beta), ldg=eo.logdensity_and_gradient(x, beta))
    eo.close()
    return out


@pytest.mark.parametrize("gemm", ["dmma", "simt"])
def test_c5_full_shape_entry_points(gemm, c5, gpu_lib, monkeypatch):
    """pgn_log_potential / pgn_logdensity_and_gradient at d = 4096, n_data = 65536 (16 split-K chunks, 512 row tiles):
    the FP64 tensor-core GEMM and the SIMT GEMM against the oracle's sequential-fma statement, bit for bit."""
    monkeypatch.setenv("PGN_GEMM", gemm)
    eg = pg.Engine(gpu_lib, n_chains=4, seed=1, **c5["target"].engine_config())
    lp = eg.log_potential(c5["x"], c5["beta"])
    ld, g = eg.logdensity_and_gradient(c5["x"], c5["beta"])
    eg.close()
    np.testing.assert_allclose(lp, c5["lp"], rtol=RTOL, atol=0)
    np.testing.assert_allclose(ld, c5["ldg"][0], rtol=RTOL, atol=0)
    np.testing.assert_allclose(g, c5["ldg"][1], rtol=RTOL, atol=1e-300)
    assert np.array_equal(lp, c5["lp"]) and np.array_equal(ld, c5["ldg"][0]) and np.array_equal(g, c5["ldg"][1])


def test_c5_full_shape_round(c5, gpu_lib, oracle_lib):
    """A 4-chain LOGREG ladder at the full C5 shape, driven through the C ABI directly: 3 scans (MH and the reversed
    search active from scan 2) with one refreshment per scan so that the oracle finishes in seconds, on a schedule whose
    first pairs are close enough to exchange their 32 KB states; every output and the final replicas bit for bit."""
    betas = np.array([0.0, 2e-6, 5e-6, 1.0])
    ex = pg.AutoMALA(base_n_refresh=1, exponent_n_refresh=0.0, step_size=0.02)
    outs = []
    for lib in (gpu_lib, oracle_lib):
        e = pg.Engine(lib, n_chains=4, seed=1, **c5["target"].engine_config())
        e.init_replicas()
        e.set_schedule(betas)
        e.set_explorer(**ex.engine_params(4096))
        r = e.run_round(3, log_index_process=True, log_swaps=True, log_target_trace=True)
        outs.append((r, e.get_state()))
        e.close()
    (rg, sg), (ro, so) = outs
    assert ro.swap_accept.sum() > 0, "the schedule was meant to produce accepted swaps"
    for k in ("index_process", "swap_accept", "swap_u", "swap_lr", "swap_n", "swap_mean", "logsum_fwd", "logsum_bwd",
              "expl_acc_n", "expl_acc_mean", "expl_n_steps", "am_n", "am_mean", "rev_n", "rev_mean", "online_mean",
              "online_var", "target_trace"):
        assert np.array_equal(getattr(rg, k), getattr(ro, k)), f"c5 full shape: {k}"
    assert rg.n_ref_equiv_evals == ro.n_ref_equiv_evals and rg.n_round_trips == ro.n_round_trips
    for k in sg:
        assert np.array_equal(sg[k], so[k]), f"c5 full shape: final replica {k}"


def test_mixed_target_entry_point_and_support(gpu_lib, oracle_lib):
    """pgn_log_potential of the Bool / Integer / Float product target, including points outside the support of the
    discrete coordinates (a Bool that is neither 0 nor 1, a non-integer or out-of-range count): -Inf like
    Distributions.logpdf, bit-equal to the oracle everywhere else."""
    t = pg.MixedProduct(n_bool=2, n_int=2, n_float=1)
    eg = pg.Engine(gpu_lib, n_chains=4, seed=1, **t.engine_config())
    eo = pg.Engine(oracle_lib, n_chains=4, seed=1, **t.engine_config())
    rng = np.random.default_rng(3)
    x = np.column_stack([rng.integers(0, 2, 40), rng.integers(0, 2, 40), rng.integers(0, 11, 40), rng.integers(0, 11, 40),
                         rng.normal(0, 2, 40)]).astype(np.float64)
    x[0, 0] = 0.5; x[1, 2] = 2.5; x[2, 3] = -1.0; x[3, 2] = 11.0
    beta = rng.uniform(0, 1, 40); beta[4] = 0.0; beta[5] = 1.0; beta[0] = 0.0; beta[1] = 1.0
    lg, lo = eg.log_potential(x, beta), eo.log_potential(x, beta)
    assert np.array_equal(lg, lo)
    assert np.all(np.isneginf(lg[:4])) and np.all(np.isfinite(lg[4:]))
    eg.close(); eo.close()


def test_slice_sampler_bool_and_integer_on_the_device(gpu_lib):
    """test/test_slice_sampler.jl:56-75 on the CUDA path: moments of [Bernoulli(0.5), Binomial(10, 0.5), Normal(0, 1)] within
    the reference's 0.2, the discrete coordinates stay in their supports, and a non-integer width on an integer coordinate
    is an error (test/test_slice_sampler.jl:113-121)."""
    t = pg.MixedProduct(n_bool=1, n_int=1, n_float=1)
    pt = pg.pigeons(target=t, n_chains=6, n_rounds=11, seed=1, record=[pg.online, pg.traces], engine_lib=gpu_lib)
    mean, std = t.target_moments()
    rr = pt.reduced_recorders
    assert np.all(np.abs(rr.online_mean - mean) <= 0.2) and np.all(np.abs(np.sqrt(rr.online_var) - std) <= 0.2)
    tr = rr.target_trace
    assert set(np.unique(tr[:, 0])) <= {0.0, 1.0}
    assert np.all(tr[:, 1] == np.floor(tr[:, 1])) and tr[:, 1].min() >= 0 and tr[:, 1].max() <= 10
    assert abs(pg.stepping_stone(pt)) < 0.1
    pt.close()
    with pytest.raises(pg.EngineError):
        pg.pigeons(target=pg.MixedProduct(n_bool=0, n_int=2, n_float=0), explorer=pg.SliceSampler(w=0.1, n_passes=1),
                   n_chains=3, n_rounds=2, engine_lib=gpu_lib)


@pytest.mark.parametrize("name", ["ising5", "ising32"])
def test_ising_table_free_kernel_parity(name, gpu_lib, oracle_lib, monkeypatch):
    """PGN_ISING_LITE=1: the Ising kernel that evaluates the Metropolis ratios where they are needed (no shared-memory
    table, 64 registers; selected automatically for shards beyond ~1.9 K chains) gives the same bits."""
    monkeypatch.setenv("PGN_ISING_LITE", "1")
    kw = CASES[name]
    assert_same(run_pt(gpu_lib, **kw), run_pt(oracle_lib, **kw), name + "/lite")


def test_c4_full_ladder_on_one_gpu(gpu_lib, oracle_lib):
    """BASELINE config 4 at its full shape on ONE GPU: 4096 chains of the 32 x 32 lattice (more than fit with a ratio table
    per chain, so the table-free kernel runs), rounds 1..2, bit-exact against the oracle + the scan invariants."""
    kw = dict(target=pg.IsingLogPotential(0.4406867935097715, 32), n_chains=4096, n_rounds=2, seed=1,
              record=[pg.index_process, pg.swap_trace])
    g, c = run_pt(gpu_lib, **kw), run_pt(oracle_lib, **kw)
    assert_same(g, c, "c4_ising32_4096")
    rr = g["rr"]
    last = type("R", (), dict(index_process=rr.index_process[-4:], swap_accept=rr.swap_accept[-4:], swap_lr=rr.swap_lr[-4:],
                              swap_u=rr.swap_u[-4:]))
    check_scan_invariants(last, 4096)


def test_plain_c_driver_matches_the_python_binding(gpu_lib):
    """tests/c_abi_smoke.c drives the library from plain C through include/pigeons_b200.h (dlopen, no Python, no C++):
    toy MVN d=2, 10 chains, SliceSampler, rounds 1..6 on the default schedule.  Same calls through the ctypes binding give
    the same numbers to the last printed digit (17 significant digits = every bit of a double)."""
    import json
    import os
    import subprocess
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    exe, src = os.path.join(root, "tests", "c_abi_smoke.bin"), os.path.join(root, "tests", "c_abi_smoke.c")
    if not os.path.exists(exe) or os.path.getmtime(exe) < os.path.getmtime(src):
        subprocess.run(["gcc", "-std=c11", "-Wall", "-I", os.path.join(root, "include"), "-o", exe, src, "-ldl", "-lm"], check=True)
    c = json.loads(subprocess.run([exe, pg.default_library_path()], check=True, capture_output=True, text=True).stdout)
    assert c["rc"] == 0
    t = pg.toy_mvn_target(2)
    e = pg.Engine(gpu_lib, n_chains=10, seed=1, **t.engine_config())
    e.init_replicas()
    e.set_schedule(np.arange(10) / 9.0)
    e.set_explorer(**pg.SliceSampler().engine_params(2))
    for rnd in range(1, 7):
        r = e.run_round(2 ** rnd)
    e.close()
    assert c["swap_mean"] == [float(v) for v in r.swap_mean[:9]]
    assert c["online_mean"] == [float(v) for v in r.online_mean]
    assert c["n_round_trips"] == r.n_round_trips
    e1 = sum(float(r.logsum_fwd[i]) - np.log(float(r.swap_n[i])) for i in range(9))
    e2 = sum(float(r.logsum_bwd[i]) - np.log(float(r.swap_n[i])) for i in range(9))
    assert abs(c["stepping_stone"] - 0.5 * (e1 - e2)) < 1e-12
