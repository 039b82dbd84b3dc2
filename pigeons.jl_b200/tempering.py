"""Between-round adaptation of the annealing schedule (host side, once per round).

Restates, for the standalone Python harness, the step on the far side of the
hot path.  In production this stays Julia.

Reference: src/tempering/adaptation.jl:56-112, src/schedules/Schedule.jl:5-44,
src/tempering/NonReversiblePT.jl:7-74.  The monotone cubic interpolation is the
Fritsch-Carlson scheme of Interpolations.jl (`FritschCarlsonMonotonicInterpolation`,
third party, not vendored; restated from Fritsch & Carlson 1980).
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Callable, Optional

import numpy as np


class MonotoneCubic:
    """Fritsch-Carlson monotone cubic Hermite interpolant through (x_k, y_k)."""

    def __init__(self, x, y):
        x = np.asarray(x, dtype=np.float64)
        y = np.asarray(y, dtype=np.float64)
        assert x.ndim == 1 and x.shape == y.shape and x.size >= 2
        assert np.all(np.diff(x) > 0), "knots must be strictly increasing"
        n = x.size
        h = np.diff(x)
        delta = np.diff(y) / h
        m = np.empty(n)
        m[0] = delta[0]
        m[-1] = delta[-1]
        for k in range(1, n - 1):
            m[k] = 0.0 if delta[k] * delta[k - 1] < 0 else (delta[k - 1] + delta[k]) / 2
        for k in range(n - 1):
            if delta[k] == 0.0:
                m[k] = 0.0
                m[k + 1] = 0.0
                continue
            a = m[k] / delta[k]
            b = m[k + 1] / delta[k]
            tau = 3.0 / np.sqrt(a * a + b * b) if (a != 0.0 or b != 0.0) else np.inf
            if tau < 1.0:
                m[k] = tau * a * delta[k]
                m[k + 1] = tau * b * delta[k]
        self.x, self.y, self.m, self.h, self.delta = x, y, m, h, delta
        self.c = (3.0 * delta - 2.0 * m[:-1] - m[1:]) / h
        self.d = (m[:-1] + m[1:] - 2.0 * delta) / (h * h)

    def _locate(self, t):
        k = np.searchsorted(self.x, t, side="right") - 1
        return np.clip(k, 0, self.x.size - 2)

    def __call__(self, t):
        t = np.asarray(t, dtype=np.float64)
        k = self._locate(t)
        s = t - self.x[k]
        return self.y[k] + s * (self.m[k] + s * (self.c[k] + s * self.d[k]))

    def gradient(self, t):
        t = np.asarray(t, dtype=np.float64)
        k = self._locate(t)
        s = t - self.x[k]
        return self.m[k] + s * (2.0 * self.c[k] + 3.0 * s * self.d[k])


@dataclass
class Schedule:
    """src/schedules/Schedule.jl:5-28 — grid points 0 = b_1 < ... < b_N = 1."""
    grids: np.ndarray

    def __post_init__(self):
        g = np.asarray(self.grids, dtype=np.float64)
        if g.size == 1:
            assert g[0] == 1.0
        else:
            assert np.all(np.diff(g) > 0) and g[0] == 0.0 and g[-1] == 1.0, f"Invalid schedule: {g}"
        self.grids = g

    @property
    def n_chains(self) -> int:
        return int(self.grids.size)


def equally_spaced_schedule(n_chains: int) -> Schedule:
    """Schedule.jl:36-44 (Julia range 0.0:(1/(n-1)):1.0: element i = i*step, last exactly 1.0)."""
    assert n_chains >= 1
    if n_chains == 1:
        return Schedule(np.array([1.0]))
    step = 1.0 / (n_chains - 1)
    g = np.arange(n_chains, dtype=np.float64) * step
    g[-1] = 1.0
    return Schedule(g)


def rejections(swap_n: np.ndarray, swap_mean: np.ndarray, n_chains: int) -> np.ndarray:
    """adaptation.jl:103-112: 1 - mean acceptance, default 0.5 when a pair never recorded."""
    acc = np.where(np.asarray(swap_n[: n_chains - 1]) > 0, np.asarray(swap_mean[: n_chains - 1]), 0.5)
    return 1.0 - acc


@dataclass
class CommunicationBarriers:
    """adaptation.jl:56-65."""
    localbarrier: Callable
    cumulativebarrier: Callable
    globalbarrier: float


def communication_barriers(intensity, schedule_grids) -> CommunicationBarriers:
    intensity = np.asarray(intensity, dtype=np.float64)
    x = np.asarray(schedule_grids, dtype=np.float64)
    assert x.size == intensity.size + 1 and np.all(intensity >= 0)
    y = np.concatenate([[0.0], np.cumsum(intensity)])
    cum = MonotoneCubic(x, y)
    return CommunicationBarriers(localbarrier=cum.gradient, cumulativebarrier=cum, globalbarrier=float(np.sum(intensity)))


def optimal_schedule_generator(intensity, old_schedule, nudged: bool = False) -> MonotoneCubic:
    """adaptation.jl:74-86."""
    intensity = np.asarray(intensity, dtype=np.float64)
    old = np.asarray(old_schedule, dtype=np.float64)
    assert old.size == intensity.size + 1 and np.all(intensity >= 0), f"Bad intensities: {intensity}"
    x = np.concatenate([[0.0], np.cumsum(intensity)])
    with np.errstate(invalid="ignore", divide="ignore"):      # all-zero intensities: 0/0 = NaN, caught by the uniqueness test as in the reference
        x = x / x[-1]
    if np.unique(x).size != x.size:
        assert not nudged
        return optimal_schedule_generator(intensity + 1e-6, old, True)
    return MonotoneCubic(x, old)


def optimal_schedule(intensity, old_schedule: Schedule, new_n_chains: Optional[int] = None) -> Schedule:
    """adaptation.jl:88-93."""
    n = new_n_chains or old_schedule.n_chains
    gen = optimal_schedule_generator(intensity, old_schedule.grids)
    step = 1.0 / (n - 1)
    # Julia: step:step:(1-step) — the n-2 interior points k*step
    grid = step * np.arange(1, n - 1)
    return Schedule(np.concatenate([[0.0], gen(grid), [1.0]]))
