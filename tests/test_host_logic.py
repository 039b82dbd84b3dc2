"""Host-side logic that needs neither GPU nor oracle: LoadBalance, the
Fritsch-Carlson schedule update, explorer parameter derivation, C-ABI symbols."""
import ctypes as C
import os
import re

import numpy as np
import pytest

import pigeons_jl_b200 as pg

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_load_balance_partition():
    """test/test_mpi_utils.jl:6-23: the slices partition 1..n for all (p, n) <= (20, 30)."""
    for p in range(1, 21):
        for n in range(p, 31):
            seen = []
            for i in range(1, p + 1):
                lb = pg.LoadBalance(i, p, n)
                idx = list(lb.my_global_indices())
                assert len(idx) == lb.my_load()
                for g in idx:
                    assert lb.find_process(g) == i
                seen += idx
            assert seen == list(range(1, n + 1))


def test_monotone_cubic_interpolates_and_is_monotone():
    rng = np.random.default_rng(0)
    x = np.concatenate([[0.0], np.cumsum(rng.uniform(0.01, 1, 12))])
    y = np.concatenate([[0.0], np.cumsum(rng.uniform(0.0, 1, 12))])
    f = pg.MonotoneCubic(x, y)
    np.testing.assert_allclose(f(x), y, rtol=0, atol=1e-12)
    t = np.linspace(x[0], x[-1], 4001)
    assert np.all(np.diff(f(t)) >= -1e-12)
    h = 1e-6
    np.testing.assert_allclose(f.gradient(t[1:-1]), (f(t[1:-1] + h) - f(t[1:-1] - h)) / (2 * h), rtol=1e-4, atol=1e-5)


def test_optimal_schedule_equalises_rejection():
    """adaptation.jl:74-93: with a uniform intensity the schedule is unchanged; with a
    skewed one the new grid points move toward the high-rejection region."""
    old = pg.equally_spaced_schedule(6)
    same = pg.optimal_schedule(np.full(5, 0.3), old)
    np.testing.assert_allclose(same.grids, old.grids, atol=1e-12)
    skew = pg.optimal_schedule(np.array([0.9, 0.1, 0.1, 0.1, 0.1]), old)
    assert skew.grids[0] == 0.0 and skew.grids[-1] == 1.0 and np.all(np.diff(skew.grids) > 0)
    assert skew.grids[1] < old.grids[1]
    # zero intensities are nudged (adaptation.jl:81-84)
    z = pg.optimal_schedule(np.array([0.0, 0.0, 0.5, 0.0, 0.0]), old)
    assert np.all(np.diff(z.grids) > 0)


def test_rejections_default_one_half():
    r = pg.rejections(np.array([4, 0, 2, 0]), np.array([0.25, 0.0, 1.0, 0.0]), 4)
    np.testing.assert_allclose(r, [0.75, 0.5, 0.0])


def test_automala_n_refresh():
    """AutoMALA.jl:122: base_n_refresh * ceil(Int, d^0.35)."""
    a = pg.AutoMALA()
    assert [a.n_refresh(d) for d in (1, 2, 10, 32, 128, 1000, 4096)] == [3, 6, 9, 12, 18, 36, 57]


def test_default_explorers():
    assert isinstance(pg.toy_mvn_target(2).default_explorer(), pg.ToyExplorer)       # toy_mvn_target.jl:13
    assert isinstance(pg.IsingLogPotential().default_explorer(), pg.IsingMetropolis)  # examples/ising.jl:95
    s = pg.SliceSampler()
    assert (s.w, s.p, s.n_passes, s.max_iter) == (10.0, 20, 3, 1024)                  # SliceSampler.jl:8-20


def declared_symbols():
    hdr = open(os.path.join(ROOT, "include", "pigeons_b200.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    return sorted(set(re.findall(r"\b(pgn_[a-z_0-9]+)\s*\(", hdr)))


def test_header_and_binding_agree():
    from pigeons_jl_b200 import _capi
    assert declared_symbols() == sorted(_capi.DECLARED_SYMBOLS)


def test_capi_library_loads_and_exports_every_symbol():
    """The C-ABI shared library loads on a machine without a GPU and exports every
    symbol include/pigeons_b200.h declares (no compute calls here)."""
    path = pg.default_library_path()
    if not os.path.exists(path):
        import __graft_entry__ as g
        g.build()
    lib = C.CDLL(path)
    for name in declared_symbols():
        assert hasattr(lib, name), name
    lib.pgn_abi_version.restype = C.c_int
    assert lib.pgn_abi_version() == pg._capi.ABI_VERSION == 4


def test_product_fails_loudly_without_gpu():
    """No CPU fallback: without a CUDA device pgn_create returns PGN_ERR_NO_DEVICE."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    lib = pg.EngineLib()
    with pytest.raises(pg.EngineError) as ei:
        pg.Engine(lib, n_chains=4, seed=1, **pg.toy_mvn_target(2).engine_config())
    assert ei.value.code == 2


def test_package_does_not_reference_the_oracle():
    pkg = os.path.join(ROOT, "pigeons.jl_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                text = open(os.path.join(dirpath, f)).read()
                assert "liborc" not in text and "oracle_adapter" not in text and "orc_engine" not in text, f


@pytest.mark.parametrize("kw", [
    dict(target=pg.toy_mvn_target(3), explorer=pg.SliceSampler(), n_chains=5, seed=2),
    dict(target=pg.Funnel(6), explorer=pg.AutoMALA(), n_chains=6, seed=3),
    dict(target=pg.IsingLogPotential(0.6, 5), n_chains=5, seed=4),
    dict(target=pg.toy_mvn_target(4), explorer=pg.Compose(pg.SliceSampler(), pg.AutoMALA()), n_chains=4, seed=5),
], ids=["toy_slice", "funnel_automala", "ising", "toy_compose"])
def test_checkpoint_resume_is_bit_identical(kw):
    """test/test_resume.jl / test_checkpoint.jl: stopping after round 5 and resuming to round 8 gives the run
    that went straight to round 8 — schedule, explorer, statistics and every replica, bit for bit
    (Shared + Replica state through pgn_get_state / pgn_set_state, on the CPU engine here)."""
    from oracle_adapter import load_oracle
    lib = load_oracle()
    rec = [pg.index_process, pg.swap_trace]
    straight = pg.pigeons(engine_lib=lib, n_rounds=8, record=rec, **kw)
    first = pg.pigeons(engine_lib=lib, n_rounds=5, record=rec, **kw)
    ckpt = pg.write_checkpoint(first)
    first.close()
    resumed = pg.resume(ckpt, pg.Inputs(engine_lib=lib, n_rounds=8, record=rec, **kw))
    a, b = straight.reduced_recorders, resumed.reduced_recorders
    for k in ("index_process", "swap_lr", "swap_u", "swap_accept", "swap_mean", "logsum_fwd", "logsum_bwd", "expl_n_steps"):
        assert np.array_equal(getattr(a, k), getattr(b, k)), k
    assert np.array_equal(straight.shared.tempering.schedule.grids, resumed.shared.tempering.schedule.grids)
    assert straight.shared.explorer == resumed.shared.explorer
    sa, sb = straight.engine.get_state(), resumed.engine.get_state()
    for k in sa:
        assert np.array_equal(sa[k], sb[k]), k
    assert pg.stepping_stone(straight) == pg.stepping_stone(resumed)
    straight.close(); resumed.close()


def test_report_shape_of_a_default_run():
    """test/test_apis.jl:12-20: a default run (10 chains, 10 rounds) reports 10 rounds x 9 swap pairs."""
    from oracle_adapter import load_oracle
    pt = pg.pigeons(target=pg.toy_mvn_target(2), engine_lib=load_oracle())
    assert pt.inputs.n_chains == 10 and pt.inputs.n_rounds == 10 and len(pt.round_log) == 10
    rr = pt.reduced_recorders
    assert rr.n_scans == 2 ** 10
    pairs = int(np.sum(rr.swap_n > 0))
    assert pairs == 9 and rr.swap_n[-1] == 0        # pair (i, i+1) is kept by chain i; chain N has none
    assert np.all(rr.swap_n[:9] == 2 ** 9)          # every pair is proposed on every other scan
    pt.close()


def test_checked_round_serial_rerun_and_failure_detection():
    """src/pt/checks.jl:36-78: `checked_round = k` re-runs rounds 1..k serially and compares every replica,
    the schedule and the explorer; a run that differs raises."""
    from oracle_adapter import load_oracle
    lib = load_oracle()
    kw = dict(target=pg.Funnel(6), explorer=pg.AutoMALA(), n_chains=5, n_rounds=5, seed=7, engine_lib=lib)
    pt = pg.pigeons(checked_round=3, **kw)            # passes silently
    pt.close()

    class Drifting:                                    # an engine whose replicas' RNG positions are off by one
        built = 0

        def __init__(self, **cfg):
            self.e = pg.Engine(lib, **cfg)
            Drifting.built += 1
            self.drift = Drifting.built == 1           # the checked run drifts; the serial re-run (same factory) does not

        def __getattr__(self, name):
            return getattr(self.e, name)

        def get_state(self):
            st = self.e.get_state()
            if self.drift:
                st["rng_counter"] = st["rng_counter"] + np.uint64(1)
            return st
    with pytest.raises(pg.ChecksFailed):
        pg.pigeons(checked_round=2, engine_factory=lambda **cfg: Drifting(**cfg), **kw)


def test_serial_rerun_failure_reaches_every_rank():
    """A serial re-run that raises on rank 0 must come back as ChecksFailed (through the verdict gather), not as a hang
    of the other ranks."""
    from oracle_adapter import load_oracle
    lib = load_oracle()
    kw = dict(target=pg.toy_mvn_target(2), explorer=pg.SliceSampler(), n_chains=4, n_rounds=3, seed=1, engine_lib=lib)
    built = []

    def factory(**cfg):
        built.append(1)
        if len(built) > 1:
            raise RuntimeError("no room for the serial ladder")
        return pg.Engine(lib, **cfg)
    with pytest.raises(pg.ChecksFailed, match="serial re-run on rank 0 failed"):
        pg.pigeons(checked_round=2, engine_factory=factory, **kw)


def test_set_state_rejects_wrong_shapes_and_resume_rejects_another_layout():
    """pgn_set_state copies n_local rows unconditionally, so the binding refuses buffers of any other size;
    `resume` refuses a checkpoint written by another shard layout or another target."""
    from oracle_adapter import load_oracle
    lib = load_oracle()
    kw = dict(target=pg.toy_mvn_target(3), explorer=pg.SliceSampler(), n_chains=5, seed=2)
    pt = pg.pigeons(engine_lib=lib, n_rounds=2, **kw)
    st = pt.engine.get_state()
    with pytest.raises(ValueError):
        pt.engine.set_state(x=st["x"][:-1])
    with pytest.raises(ValueError):
        pt.engine.set_state(replica_index=np.arange(1, 7, dtype=np.int32))
    with pytest.raises(ValueError):
        pt.engine.set_state(replica_index=np.array([0, 1, 2, 3, 4], dtype=np.int32))
    with pytest.raises(ValueError):
        pt.engine.set_state(x=np.zeros((5, 4)))
    ck = pg.write_checkpoint(pt)
    pt.close()
    assert ck["world_size"] == 1 and ck["n_local"] == 5 and ck["dim"] == 3
    with pytest.raises(ValueError):
        pg.resume(ck, pg.Inputs(engine_lib=lib, n_rounds=3, **{**kw, "target": pg.toy_mvn_target(4)}))
    bad = dict(ck, world_size=2)
    with pytest.raises(ValueError):
        pg.resume(bad, pg.Inputs(engine_lib=lib, n_rounds=3, **kw))


def test_adaptation_without_swap_statistics():
    """adaptation.jl:103-112: with no swap recorded a pair counts as acceptance 0.5, so the barriers exist
    (global barrier 0.5 (N-1)); a single chain leaves the tempering untouched."""
    from oracle_adapter import load_oracle
    lib = load_oracle()
    pt = pg.pigeons(target=pg.toy_mvn_target(2), explorer=pg.SliceSampler(), n_chains=1, n_rounds=2, engine_lib=lib)
    assert np.isnan(pg.global_barrier(pt))
    pt.close()
    from pigeons_jl_b200.pt import adapt
    pt = pg.create_pt(pg.Inputs(target=pg.toy_mvn_target(2), explorer=pg.SliceSampler(), n_chains=5, n_rounds=0, engine_lib=lib))
    pt.shared.iterators.round = 1
    pt.engine.set_schedule(pt.shared.tempering.schedule.grids)
    pt.engine.set_explorer(**pt.shared.explorer.engine_params(2))
    from pigeons_jl_b200.recorders import merge_round_results
    res = pt.engine.run_round(0)       # a round with no scans records nothing
    rr = merge_round_results(pt.inputs.comm, res, 5, 2)
    assert not rr.has_swap_stats
    pt = adapt(pt, rr)
    assert pg.global_barrier(pt) == pytest.approx(0.5 * 4)
    pt.close()


def test_shard_layout_keeps_both_target_chains_together():
    """Engine rule for two legs (pgn_local_range): balanced LoadBalance blocks, except that chain n_var + 1 joins chain
    n_var's shard when a block boundary would separate the two target chains."""
    for n in range(2, 40):
        for world in range(1, min(n, 9) + 1):
            plain = pg.shard_layout(n, world)
            assert plain == [(pg.LoadBalance(r, world, n).my_first_global_idx(), pg.LoadBalance(r, world, n).my_load())
                             for r in range(1, world + 1)]
            for nv in range(1, n):
                try:
                    lay = pg.shard_layout(n, world, nv)
                except ValueError:
                    assert any(f + c - 1 == nv and plain[r + 1][1] < 2 for r, (f, c) in enumerate(plain[:-1]))
                    continue
                assert lay[0][0] == 1 and sum(c for _, c in lay) == n and all(c >= 1 for _, c in lay)
                assert all(lay[r][0] + lay[r][1] == lay[r + 1][0] for r in range(world - 1))          # contiguous, ordered
                owner = [r for r, (f, c) in enumerate(lay) if f <= nv < f + c]
                assert len(owner) == 1 and lay[owner[0]][0] <= nv + 1 < lay[owner[0]][0] + lay[owner[0]][1]
                assert sum(a != b for a, b in zip(lay, plain)) in (0, 2)
