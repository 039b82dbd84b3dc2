"""Explorer configuration structs (the `explorer` informal interface,
src/explorers/explorer.jl:7-39).  These are host-side parameter records with the
reference's field names and defaults; the kernels that execute them live in
csrc/pgn_kernels.cuh (VecChain / IsingChain) and csrc/pgn_logreg.cuh.  `adapt_explorer` restates src/explorers/AutoMALA.jl:70-79.
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field, replace
from typing import Optional

import numpy as np

from . import _capi


@dataclass(frozen=True)
class ToyExplorer:
    """src/explorers/ToyExplorer.jl:5-14: i.i.d. sampling at every chain (toy paths only)."""

    def engine_params(self, dim: int) -> dict:
        return dict(kind=_capi.EXPLORER_TOY)


@dataclass(frozen=True)
class SliceSampler:
    """src/explorers/SliceSampler.jl:8-20."""
    w: float = 10.0
    p: int = 20
    n_passes: int = 3
    max_iter: int = 1024

    def engine_params(self, dim: int) -> dict:
        return dict(kind=_capi.EXPLORER_SLICE, slice_w=self.w, slice_p=self.p, slice_n_passes=self.n_passes,
                    slice_max_iter=self.max_iter)


@dataclass(frozen=True)
class IdentityPreconditioner:        # Preconditioner.jl:14
    kind: int = _capi.PRECOND_IDENTITY


@dataclass(frozen=True)
class DiagonalPreconditioner:        # Preconditioner.jl:22
    kind: int = _capi.PRECOND_DIAGONAL


@dataclass(frozen=True)
class MixDiagonalPreconditioner:     # Preconditioner.jl:40-50 (defaults 1//3, 1//3)
    p0: float = 1.0 / 3.0
    p1: float = 1.0 / 3.0
    kind: int = _capi.PRECOND_MIX_DIAGONAL

    def __post_init__(self):
        if not (0.0 <= self.p0 + self.p1 <= 1.0):
            raise ValueError("p0+p1 < 0 or p0+p1 > 1")


@dataclass(frozen=True)
class AutoMALA:
    """src/explorers/AutoMALA.jl:29-68.  `default_autodiff_backend` has no device
    counterpart: every device target ships an analytic gradient."""
    base_n_refresh: int = 3
    exponent_n_refresh: float = 0.35
    step_size: float = 1.0
    preconditioner: object = field(default_factory=MixDiagonalPreconditioner)
    estimated_target_std_deviations: Optional[tuple] = None

    def n_refresh(self, dim: int) -> int:          # AutoMALA.jl:122
        return self.base_n_refresh * math.ceil(dim ** self.exponent_n_refresh)

    def engine_params(self, dim: int) -> dict:
        pc = self.preconditioner
        sd = None if self.estimated_target_std_deviations is None else np.asarray(self.estimated_target_std_deviations)
        p0 = getattr(pc, "p0", 1.0 / 3.0)
        p1 = getattr(pc, "p1", 1.0 / 3.0)
        return dict(kind=_capi.EXPLORER_AUTOMALA, n_refresh=self.n_refresh(dim), step_size=self.step_size,
                    precond_kind=pc.kind, mix_p0=p0, mix_p01=p0 + p1, std_devs=sd)

    def adapt(self, round_result) -> "AutoMALA":
        """adapt_explorer (AutoMALA.jl:70-79): new step size = old * mean over chains
        of the mean am_factor; std devs from the target chain's online variance."""
        am_n = np.asarray(round_result.am_n)
        am_mean = np.asarray(round_result.am_mean)
        recorded = am_n > 0
        factor = float(np.mean(am_mean[recorded])) if recorded.any() else 1.0
        sd = None
        if self.preconditioner.kind != _capi.PRECOND_IDENTITY:     # Preconditioner.jl:53-55
            sd = tuple(np.sqrt(np.asarray(round_result.online_var)).tolist())
        return replace(self, step_size=self.step_size * factor, estimated_target_std_deviations=sd)


@dataclass(frozen=True)
class MALA:
    """src/explorers/MALA.jl:19-55: fixed step size, preconditioner adapted every round."""
    base_n_refresh: int = 3
    exponent_n_refresh: float = 0.35
    step_size: float = 1.0
    preconditioner: object = field(default_factory=MixDiagonalPreconditioner)
    estimated_target_std_deviations: Optional[tuple] = None

    def n_refresh(self, dim: int) -> int:          # MALA.jl:80
        return self.base_n_refresh * math.ceil(dim ** self.exponent_n_refresh)

    def engine_params(self, dim: int) -> dict:
        pc = self.preconditioner
        sd = None if self.estimated_target_std_deviations is None else np.asarray(self.estimated_target_std_deviations)
        p0 = getattr(pc, "p0", 1.0 / 3.0)
        p1 = getattr(pc, "p1", 1.0 / 3.0)
        return dict(kind=_capi.EXPLORER_MALA, n_refresh=self.n_refresh(dim), step_size=self.step_size,
                    precond_kind=pc.kind, mix_p0=p0, mix_p01=p0 + p1, std_devs=sd)

    def adapt(self, round_result) -> "MALA":
        """adapt_explorer (MALA.jl:57-63): only the preconditioner's std devs change."""
        sd = None
        if self.preconditioner.kind != _capi.PRECOND_IDENTITY:
            sd = tuple(np.sqrt(np.asarray(round_result.online_var)).tolist())
        return replace(self, estimated_target_std_deviations=sd)


@dataclass(frozen=True)
class IsingMetropolis:
    """examples/ising.jl:91-93."""
    n_steps: int = 3

    def engine_params(self, dim: int) -> dict:
        return dict(kind=_capi.EXPLORER_ISING_METROPOLIS, ising_n_steps=self.n_steps)


def _program_steps(explorers, dim: int):
    """(kind, n_refresh, step_size, precond_kind, p0, p01) per explorer of a Compose / Mix, the shared SliceSampler
    parameters and the shared std-dev estimate (every explorer adapts from the same recorders, Compose.jl:10-14, Mix.jl:14-17)."""
    steps, slice_params, sds = [], None, None
    for e in explorers:
        q = e.engine_params(dim)
        kind = q["kind"]
        if kind == _capi.EXPLORER_SLICE:
            sp = {k: q[k] for k in ("slice_w", "slice_p", "slice_n_passes", "slice_max_iter")}
            if slice_params is not None and sp != slice_params:
                raise NotImplementedError("the SliceSamplers of one Compose / Mix share their parameters on the device")
            slice_params = sp
            steps.append((kind, 0, 1.0, _capi.PRECOND_IDENTITY, 1.0 / 3.0, 2.0 / 3.0))
        elif kind in (_capi.EXPLORER_AUTOMALA, _capi.EXPLORER_MALA):
            steps.append((kind, q["n_refresh"], q["step_size"], q["precond_kind"], q["mix_p0"], q["mix_p01"]))
            if q.get("std_devs") is not None and sds is None:
                sds = np.asarray(q["std_devs"])
        elif kind == _capi.EXPLORER_TOY:
            steps.append((kind, 0, 1.0, _capi.PRECOND_IDENTITY, 1.0 / 3.0, 2.0 / 3.0))
        else:
            raise NotImplementedError(f"{type(e).__name__} cannot be part of a device Compose / Mix (no CPU fallback)")
    return steps, (slice_params or {}), sds


@dataclass(frozen=True, init=False)
class Compose:
    """src/explorers/Compose.jl:5-27: every explorer in turn at each step, all feeding the same recorders.
    Device support: 1..4 explorers out of ToyExplorer (toy MVN), SliceSampler, MALA, AutoMALA, in any order
    (`Compose(SliceSampler(), AutoMALA())` is the combination the reference documents and tests: Compose.jl:3,
    test/test_parallelism_invariance.jl:19, test/test_DistributionLogPotential.jl:57)."""
    explorers: tuple

    def __init__(self, *explorers):
        if len(explorers) == 1 and isinstance(explorers[0], (tuple, list)):
            explorers = tuple(explorers[0])
        if not (1 <= len(explorers) <= _capi.MAX_MIX):
            raise NotImplementedError(f"the device composes 1..{_capi.MAX_MIX} explorers")
        object.__setattr__(self, "explorers", tuple(explorers))

    @property
    def first(self):
        return self.explorers[0]

    @property
    def second(self):
        return self.explorers[1]

    def engine_params(self, dim: int) -> dict:
        steps, slice_params, sds = _program_steps(self.explorers, dim)
        return dict(kind=_capi.EXPLORER_COMPOSE, steps=steps, std_devs=sds, **slice_params)

    def adapt(self, round_result) -> "Compose":      # Compose.jl:10-14
        return Compose(*(e.adapt(round_result) if hasattr(e, "adapt") else e for e in self.explorers))


@dataclass(frozen=True, init=False)
class Mix:
    """src/explorers/Mix.jl:7-30: one of the explorers, drawn uniformly from the replica's stream,
    performs the step.  Device support: 2..4 explorers out of ToyExplorer (toy MVN), SliceSampler, MALA, AutoMALA; a
    mixture of AutoMALA kernels only (they may differ in preconditioner, step size and number of refreshments —
    test/test_parallelism_invariance.jl:14-18) runs on the plain autoMALA kernel, which draws the variant itself."""
    explorers: tuple

    def __init__(self, *explorers):
        if len(explorers) == 1 and isinstance(explorers[0], (tuple, list)):
            explorers = tuple(explorers[0])
        if not (2 <= len(explorers) <= _capi.MAX_MIX):
            raise NotImplementedError(f"the device mixes 2..{_capi.MAX_MIX} explorers")
        object.__setattr__(self, "explorers", tuple(explorers))

    def engine_params(self, dim: int) -> dict:
        if not all(isinstance(e, AutoMALA) for e in self.explorers):
            steps, slice_params, sds = _program_steps(self.explorers, dim)
            return dict(kind=_capi.EXPLORER_MIX, steps=steps, std_devs=sds, **slice_params)
        p = dict(self.explorers[0].engine_params(dim))
        variants = []
        for e in self.explorers:
            q = e.engine_params(dim)
            variants.append((q["n_refresh"], q["step_size"], q["precond_kind"], q["mix_p0"], q["mix_p01"]))
        # the variants share one std-dev estimate (every explorer adapts from the same recorders, Mix.jl:14-17)
        sds = [e.estimated_target_std_deviations for e in self.explorers if e.estimated_target_std_deviations is not None]
        p["std_devs"] = None if not sds else np.asarray(sds[0])
        p["mix_variants"] = variants
        return p

    def adapt(self, round_result) -> "Mix":          # Mix.jl:14-17
        return Mix(*(e.adapt(round_result) if hasattr(e, "adapt") else e for e in self.explorers))
