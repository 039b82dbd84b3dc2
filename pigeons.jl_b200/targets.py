"""Device target families and the `target` informal interface.

Mirrors src/targets/target.jl:4-76 (`initialization`, `default_explorer`,
`default_reference`, `sample_iid!`, `create_path`) for the closed family of
targets the engine implements on the device.  Arbitrary host callables cannot
run inside the scan kernel; an unsupported target raises (no CPU fallback).
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field
from typing import Optional, Sequence

import numpy as np

from . import _capi
from .explorers import IsingMetropolis, SliceSampler, ToyExplorer


class Target:
    """Base class: anything `pigeons(target=...)` accepts."""
    dim: int

    def default_explorer(self):          # target.jl:24
        return SliceSampler()

    def engine_config(self) -> dict:     # -> kwargs of _capi.Engine
        raise NotImplementedError


@dataclass
class ScaledPrecisionNormalPath(Target):
    """src/paths/ScaledPrecisionNormalPath.jl:5-78: zero-mean normals whose
    *precision* is interpolated, prec(b) = (1-b)*precision0 + b*precision1."""
    precision0: float
    precision1: float
    dim: int

    def default_explorer(self):          # toy_mvn_target.jl:13
        return ToyExplorer()

    def engine_config(self):
        return dict(target_kind=_capi.TARGET_TOY_MVN, dim=self.dim, p=(self.precision0, self.precision1))

    # known answers used by the reference's own tests
    def analytic_cumulativebarrier(self):     # ScaledPrecisionNormalPath.jl:56-64
        from scipy.special import beta as beta_fn
        b = beta_fn(self.dim / 2.0, self.dim / 2.0)

        def cumulativebarrier(beta):
            sigma0 = 1.0 / math.sqrt(self.precision0)
            sigmab = 1.0 / math.sqrt((1.0 - beta) * self.precision0 + beta * self.precision1)
            return 2.0 ** (2.0 - self.dim) / b * math.log(sigma0 / sigmab)
        return cumulativebarrier

    def analytic_lognormalization(self):      # ScaledPrecisionNormalPath.jl:66-71
        return 0.5 * self.dim * (math.log(self.precision0) - math.log(self.precision1))


def toy_mvn_target(dim: int) -> ScaledPrecisionNormalPath:
    """src/targets/toy_mvn_target.jl:8 and ScaledPrecisionNormalPath.jl:43-44."""
    return ScaledPrecisionNormalPath(1.0, 10.0, int(dim))


def _normal_ref_params(sigma_ref: float):
    return (float(sigma_ref), math.log(sigma_ref), 1.0 / (sigma_ref * sigma_ref))


@dataclass
class Funnel(Target):
    """Neal's funnel as in test/supporting/dimensional-analysis.jl:33-47:
    y ~ N(0, scale), z_i ~ N(0, exp(y/2)), i = 2..dim.  The reference repo has no
    PT reference for it; we use N(0, reference_sigma^2 I)."""
    dim: int
    scale: float = 3.0
    reference_sigma: float = 3.0

    def default_explorer(self):
        from .explorers import AutoMALA
        return AutoMALA()

    def engine_config(self):
        s = float(self.scale)
        p = (s, math.log(s), 1.0 / (s * s)) + _normal_ref_params(self.reference_sigma)
        return dict(target_kind=_capi.TARGET_FUNNEL, dim=self.dim, p=p)


@dataclass
class GaussianMixture(Target):
    """DistributionLogPotential(MixtureModel([MvNormal(m_k, sigma^2 I)...], w))
    (src/targets/DistributionLogPotential.jl:5-41; pattern of
    test/test_auto_mala.jl:126-132) with reference MvNormal(0, reference_sigma^2 I)."""
    means: np.ndarray                  # [K, d]
    weights: Optional[Sequence[float]] = None
    sigma: float = 1.0
    reference_sigma: float = 1.0
    dim: int = field(init=False)

    def __post_init__(self):
        self.means = np.ascontiguousarray(self.means, dtype=np.float64)
        assert self.means.ndim == 2
        self.dim = int(self.means.shape[1])
        k = self.means.shape[0]
        w = np.full(k, 1.0 / k) if self.weights is None else np.asarray(self.weights, dtype=np.float64)
        assert w.size == k and abs(w.sum() - 1.0) < 1e-12
        self.weights = w

    def default_explorer(self):
        from .explorers import AutoMALA
        return AutoMALA()

    def engine_config(self):
        d, s = self.dim, float(self.sigma)
        cst = d * math.log(s) + 0.5 * d * math.log(2.0 * math.pi)
        p = (s, cst, 1.0 / (s * s)) + _normal_ref_params(self.reference_sigma)
        return dict(target_kind=_capi.TARGET_GMM, dim=d, p=p, means=self.means,
                    log_weights=np.log(self.weights), n_modes=int(self.means.shape[0]))


def eight_mode_mixture(dim: int = 128, mu: float = 8.0) -> GaussianMixture:
    """BASELINE config 3: 8 modes at (+-mu, +-mu, +-mu, 0, ..., 0), identity
    covariances, equal weights, reference N(0, mu^2 I)."""
    means = np.zeros((8, dim))
    for k in range(8):
        for j in range(3):
            means[k, j] = mu if (k >> j) & 1 else -mu
    return GaussianMixture(means=means, sigma=1.0, reference_sigma=mu)


@dataclass
class LogisticRegression(Target):
    """BASELINE config 5: Bayesian logistic regression with an analytic gradient (the
    `BufferedAD` custom-gradient pattern of test/test_custom_gradient.jl:14-24,
    docs/src/input-julia.md:147-192).  Prior = reference = N(0, prior_sigma^2 I);
    target = prior x prod_n Bernoulli(y_n | sigmoid(x_n . theta))."""
    x: np.ndarray                      # [n_data, d]
    y: np.ndarray                      # [n_data] in {0, 1}
    prior_sigma: float = 1.0
    dim: int = field(init=False)

    def __post_init__(self):
        self.x = np.ascontiguousarray(self.x, dtype=np.float64)
        self.y = np.ascontiguousarray(self.y, dtype=np.float64)
        assert self.x.ndim == 2 and self.y.shape == (self.x.shape[0],)
        self.dim = int(self.x.shape[1])

    def default_explorer(self):
        from .explorers import AutoMALA
        return AutoMALA()

    def engine_config(self):
        p = (float(self.x.shape[0]), 0.0, 0.0) + _normal_ref_params(self.prior_sigma)
        return dict(target_kind=_capi.TARGET_LOGREG, dim=self.dim, p=p, data_x=self.x, data_y=self.y)


def synthetic_logistic_regression(n_data: int, dim: int, seed: int = 2) -> LogisticRegression:
    """SURVEY.md §8(d) C5 inputs: X ~ N(0,1)/sqrt(d), theta* ~ N(0, I), y ~ Bernoulli(sigmoid(X theta*))."""
    rng = np.random.Generator(np.random.Philox(seed))
    x = rng.standard_normal((n_data, dim)) / math.sqrt(dim)
    theta = rng.standard_normal(dim)
    pr = 1.0 / (1.0 + np.exp(-(x @ theta)))
    y = (rng.uniform(size=n_data) < pr).astype(np.float64)
    return LogisticRegression(x=x, y=y, prior_sigma=1.0)


@dataclass
class MixedProduct(Target):
    """The mixed Bool / Integer / Float state of test/test_slice_sampler.jl:56-75 as a PT target:
    target = Bernoulli(p1)^n_bool x Binomial(n, q1)^n_int x Normal(0, 1)^n_float,
    reference = Bernoulli(p0)^n_bool x Binomial(n, q0)^n_int x Normal(0, sigma_ref)^n_float.
    It exists to exercise SliceSampler's Bool (full conditional, SliceSampler.jl:65-86) and Integer
    (integer end points and draws, :136-142,:189) coordinate updates; both ends are normalised, so log(Z1/Z0) = 0."""
    n_bool: int = 1
    n_int: int = 1
    n_float: int = 1
    binomial_n: int = 10
    p0: float = 0.3
    p1: float = 0.5
    q0: float = 0.35
    q1: float = 0.5
    sigma_ref: float = 2.0

    @property
    def dim(self):
        return self.n_bool + self.n_int + self.n_float

    def default_explorer(self):
        from .explorers import SliceSampler
        return SliceSampler()

    def engine_config(self):
        n = self.binomial_n
        table = [math.log(self.p0), math.log1p(-self.p0), math.log(self.p1), math.log1p(-self.p1),
                 math.log(self.q0), math.log1p(-self.q0), math.log(self.q1), math.log1p(-self.q1), self.p0, self.q0]
        table += [math.lgamma(n + 1) - math.lgamma(k + 1) - math.lgamma(n - k + 1) for k in range(n + 1)]
        p = (float(self.n_bool), float(self.n_int), float(n)) + _normal_ref_params(self.sigma_ref)
        return dict(target_kind=_capi.TARGET_MIXED, dim=self.dim, p=p, means=np.array(table), n_modes=len(table))

    def target_moments(self):
        """mean and standard deviation of every coordinate under the target"""
        n, q = self.binomial_n, self.q1
        mean = [self.p1] * self.n_bool + [n * q] * self.n_int + [0.0] * self.n_float
        std = [math.sqrt(self.p1 * (1 - self.p1))] * self.n_bool + [math.sqrt(n * q * (1 - q))] * self.n_int + [1.0] * self.n_float
        return np.array(mean), np.array(std)


@dataclass
class UnidentifiableProduct(Target):
    """The unidentifiable two-parameter model the reference's own tests use (test/test_DistributionLogPotential.jl:7-21
    in the constrained parametrisation; toy_turing_unid_target, ext/PigeonsDynamicPPLExt/toy_examples.jl:17-19):
    n_successes ~ Binomial(n_trials, p1 p2), p1, p2 ~ Uniform(0, 1); l(p1, p2) = s log(p1 p2) + (n - s) log1p(-p1 p2) on the
    unit square, -Inf outside; the reference is the Uniform(0,1)^2 prior.  SliceSampler only."""
    n_trials: int = 100
    n_successes: Optional[int] = None

    def __post_init__(self):
        if self.n_successes is None:
            self.n_successes = math.ceil(self.n_trials / 2)

    @property
    def dim(self):
        return 2

    def default_explorer(self):
        from .explorers import SliceSampler
        return SliceSampler()

    def engine_config(self):
        return dict(target_kind=_capi.TARGET_UNID, dim=2, p=(float(self.n_trials), float(self.n_successes)))

    def analytic_lognormalization(self) -> float:
        """log of int_0^1 int_0^1 (p1 p2)^s (1 - p1 p2)^(n - s) dp1 dp2 = log( sum_{k >= s+1}^{n+1} 1/k * B(s+1, n-s+1) ... ) — by
        substitution u = p1 p2 (density -log u on (0,1)): Z = int_0^1 u^s (1-u)^(n-s) (-log u) du, evaluated by quadrature."""
        from scipy import integrate, special
        s, n = self.n_successes, self.n_trials
        logb = special.betaln(s + 1, n - s + 1)
        # E_{u ~ Beta(s+1, n-s+1)}[-log u] = digamma(n + 2) - digamma(s + 1)
        return float(logb + math.log(special.digamma(n + 2) - special.digamma(s + 1)))


@dataclass
class IsingLogPotential(Target):
    """examples/ising.jl:6-117: l(state) = beta * sum_<ij> s_i s_j on an L x L
    torus; reference = same with beta = 0 (i.i.d. Bernoulli(1/2) spins)."""
    beta: float = 1.0
    base_length: int = 5

    @property
    def dim(self):
        return self.base_length * self.base_length

    def default_explorer(self):          # examples/ising.jl:95
        return IsingMetropolis()

    def engine_config(self):
        return dict(target_kind=_capi.TARGET_ISING, dim=self.dim, p=(self.beta, self.base_length))


@dataclass
class TestSwapper(Target):
    """src/swap/pair_swapper.jl:100-149: every swap has the same acceptance probability."""
    constant_swap_accept_pr: float
    dim: int = 0
    __test__ = False   # not a pytest class

    def default_explorer(self):
        return None

    def engine_config(self):
        return dict(target_kind=_capi.TARGET_TEST_SWAPPER, dim=0, p=(self.constant_swap_accept_pr,))
