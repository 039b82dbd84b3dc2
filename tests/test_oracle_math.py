"""Pins the oracle's arithmetic (oracle/orc_math.hpp): elementary functions vs
mpmath / numpy, Philox4x32-10 vs the Random123 known-answer vectors, variate
distributions, and the canonical summation tree."""
import ctypes as C

import numpy as np
import pytest

mp = pytest.importorskip("mpmath")
mp.mp.prec = 200


def ulp_err(f_mp, xs, ys):
    worst = 0.0
    for x, y in zip(xs, ys):
        t = f_mp(mp.mpf(float(x)))
        if t == 0 or not np.isfinite(float(t)):
            continue
        e = abs((mp.mpf(float(y)) - t) / mp.mpf(float(np.spacing(abs(float(t))))))
        worst = max(worst, float(e))
    return worst


def test_exp_within_one_ulp(oracle_lib):
    rng = np.random.default_rng(0)
    x = np.concatenate([rng.uniform(-745, 709, 4000), rng.uniform(-1, 1, 4000), rng.normal(0, 1e-3, 500)])
    assert ulp_err(mp.exp, x, oracle_lib.test_math(0, x)) < 1.0


def test_log_within_one_ulp(oracle_lib):
    rng = np.random.default_rng(1)
    x = np.concatenate([np.exp(rng.uniform(-700, 700, 4000)), rng.uniform(0.5, 2, 4000),
                        1 + rng.normal(0, 1e-6, 500), rng.uniform(0, 1, 500) * 5e-310])
    assert ulp_err(mp.log, x, oracle_lib.test_math(1, x)) < 1.0


def test_special_values(oracle_lib):
    e = oracle_lib.test_math(0, np.array([np.inf, -np.inf, 710.0, -746.0, 0.0]))
    assert e[0] == np.inf and e[1] == 0.0 and e[2] == np.inf and e[3] == 0.0 and e[4] == 1.0
    assert np.isnan(oracle_lib.test_math(0, np.array([np.nan]))[0])
    l = oracle_lib.test_math(1, np.array([np.inf, 0.0, 1.0]))
    assert l[0] == np.inf and l[1] == -np.inf and l[2] == 0.0
    assert np.all(np.isnan(oracle_lib.test_math(1, np.array([-1.0, np.nan]))))
    # exp is >= 1 for non-negative arguments (the Ising accept test relies on it)
    x = np.abs(np.random.default_rng(2).normal(0, 1e-8, 10000))
    assert np.all(oracle_lib.test_math(0, x) >= 1.0)


def test_cospi_accuracy(oracle_lib):
    x = np.random.default_rng(3).uniform(0, 2, 5000)
    truth = np.array([float(mp.cos(mp.pi * mp.mpf(float(t)))) for t in x])
    assert np.max(np.abs(oracle_lib.test_math(2, x) - truth)) < 2.3e-16


def test_logaddexp(oracle_lib):
    rng = np.random.default_rng(4)
    ab = rng.normal(0, 30, 4000)
    got = oracle_lib.test_math(6, ab)
    np.testing.assert_allclose(got, np.logaddexp(ab[0::2], ab[1::2]), rtol=4e-16, atol=0)
    assert oracle_lib.test_math(6, np.array([-np.inf, 3.0]))[0] == 3.0
    assert oracle_lib.test_math(6, np.array([2.0, -np.inf]))[0] == 2.0


def test_log1p(oracle_lib):
    """log1p_ (Kahan: log(w) t / (w - 1), w = 1 + t) on (-1, inf): the unidentifiable target evaluates log1p(-p1 p2)."""
    rng = np.random.default_rng(6)
    t = np.concatenate([-rng.uniform(0, 1, 4000), rng.uniform(0, 5, 2000), rng.normal(0, 1e-9, 2000)])
    got = oracle_lib.test_math(7, t)
    np.testing.assert_allclose(got, np.log1p(t), rtol=5e-16, atol=0)
    edge = oracle_lib.test_math(7, np.array([0.0, -1.0, 1e-300, -1e-17]))
    assert edge[0] == 0.0 and edge[1] == -np.inf and edge[2] == 1e-300 and edge[3] == -1e-17


def test_philox_known_answers(oracle_lib):
    """Random123 kat_vectors for philox4x32-10."""
    f = oracle_lib.lib.orc_philox

    def ph(c, k):
        c = (C.c_uint32 * 4)(*c); k = (C.c_uint32 * 2)(*k); o = (C.c_uint32 * 4)()
        f(c, k, o)
        return list(o)
    assert ph([0, 0, 0, 0], [0, 0]) == [0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8]
    assert ph([0xffffffff] * 4, [0xffffffff] * 2) == [0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd]
    assert ph([0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344], [0xa4093822, 0x299f31d0]) == \
        [0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1]


def test_variates_are_correctly_distributed(oracle_lib):
    from scipy import stats
    ctr = np.arange(200000, dtype=np.float64)
    u = oracle_lib.test_math(4, ctr, seed=7, replica_index=3)
    assert u.min() >= 0.0 and u.max() < 1.0
    assert stats.kstest(u, "uniform").pvalue > 1e-3
    assert np.all(u * 2.0 ** 52 == np.floor(u * 2.0 ** 52))        # 52-bit construction, like Julia's rand(Float64)
    z = oracle_lib.test_math(3, ctr, seed=7, replica_index=3)
    assert stats.kstest(z, "norm").pvalue > 1e-3
    e = oracle_lib.test_math(5, ctr, seed=7, replica_index=3)
    assert e.min() >= 0.0 and stats.kstest(e, "expon").pvalue > 1e-3
    # streams of different replicas / seeds are distinct (test/test_split.jl idea)
    u2 = oracle_lib.test_math(4, ctr[:1000], seed=7, replica_index=4)
    u3 = oracle_lib.test_math(4, ctr[:1000], seed=8, replica_index=3)
    assert not np.any(u[:1000] == u2) and not np.any(u[:1000] == u3)
    assert abs(np.corrcoef(u[:1000], u2)[0, 1]) < 0.15


def test_toy_density_uses_canonical_tree(oracle_lib):
    """-0.5 * prec(beta) * tree_sum(x^2) with lanes c%32 and the xor butterfly."""
    import pigeons_jl_b200 as pg
    rng = np.random.default_rng(5)
    for d in (1, 2, 31, 32, 33, 77, 128):
        e = pg.Engine(oracle_lib, n_chains=2, seed=1, **pg.toy_mvn_target(d).engine_config())
        x = rng.normal(0, 1, (5, d))
        beta = np.array([0.0, 0.25, 0.5, 0.75, 1.0])
        got = e.log_potential(x, beta)
        for i in range(5):
            v = np.zeros(32)
            for lane in range(32):
                acc = 0.0
                for c in range(lane, d, 32):
                    acc = acc + x[i, c] * x[i, c]
                v[lane] = acc
            off = 16
            while off >= 1:
                v = np.array([v[l] + v[l ^ off] for l in range(32)])
                off //= 2
            prec = (1.0 - beta[i]) * 1.0 + beta[i] * 10.0
            assert got[i] == -0.5 * prec * v[0]
        e.close()
