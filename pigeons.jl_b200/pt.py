"""Host-side PT driver: `pigeons(...)`, `Inputs`, `PT`, `Shared`, `Iterators`.

Mirror of the reference's host orchestration for the standalone harness
(src/api.jl:8-19, src/pt/Inputs.jl:9-102, src/pt/PT.jl:6-50, src/pt/Shared.jl,
src/pt/Iterators.jl:27-49, src/pt/pigeons.jl:12-28,152-162).  The only thing
that differs from the reference's round loop is `run_one_round!`: instead of
looping over scans on the host it makes ONE C-ABI call per round
(`pgn_run_round`) that runs all 2^round scans on the GPU.
"""
from __future__ import annotations

import copy
import math
from dataclasses import dataclass, field, replace
from typing import Callable, List, Optional, Sequence

import numpy as np

from . import _capi
from .distributed import Communicator, SingleProcess, shard_layout
from .explorers import MALA, AutoMALA, Compose, Mix
from .recorders import ReducedRecorders, merge_round_results
from .tempering import (CommunicationBarriers, Schedule, communication_barriers, equally_spaced_schedule,
                        optimal_schedule, rejections)

# recorder names accepted in `record=[...]` (src/recorders/recorder.jl)
traces = "traces"
index_process = "index_process"
round_trip = "round_trip"
online = "online"
swap_trace = "swap_trace"      # engine extension: per-scan SwapStat log


@dataclass
class Inputs:
    """src/pt/Inputs.jl:9-102 (fields the scan path honours)."""
    target: object
    seed: int = 1
    n_rounds: int = 10
    n_chains: int = 10
    n_chains_variational: int = 0      # Inputs.jl:21-32: chains of the variational leg (two legs when n_chains > 0 as well)
    variational: object = None         # Inputs.jl:42-44: GaussianReference(...) or None
    explorer: object = None
    record: Sequence[str] = ()
    multithreaded: bool = False        # accepted for API compatibility; the device runs all replicas concurrently
    show_report: bool = False
    checked_round: int = 0
    # engine plumbing (not in the reference)
    recorder_order: int = _capi.RECORDERS_PER_REPLICA   # per-replica recorders + tree merge (recorders.jl:88-120); 1 = per chain
    engine_lib: Optional[_capi.EngineLib] = None
    engine_factory: Optional[Callable] = None     # (n_chains, seed, rank, world_size, device, **target_cfg) -> Engine-like
    device: int = 0
    comm: Optional[Communicator] = None

    def __post_init__(self):
        if self.explorer is None:
            self.explorer = self.target.default_explorer()     # target.jl:24, explorer.jl:49-54
        if self.comm is None:
            self.comm = SingleProcess()
        if self.variational is not None:     # the run owns (and mutates) its reference: never the caller's object
            self.variational = copy.deepcopy(self.variational)
        if self.n_chains_variational > 0 and self.recorder_order != _capi.RECORDERS_PER_REPLICA and self.n_chains > 0:
            raise ValueError("two legs need recorder_order = RECORDERS_PER_REPLICA")

    @property
    def n_chains_total(self) -> int:       # Inputs.jl:128
        return self.n_chains + self.n_chains_variational


@dataclass
class Iterators:
    """src/pt/Iterators.jl:8-24."""
    round: int = 0
    scan: int = 0


def n_scans_in_round(it: Iterators) -> int:      # Iterators.jl:49
    return 2 ** it.round


@dataclass
class NonReversiblePT:
    """src/tempering/NonReversiblePT.jl:7-28."""
    schedule: Schedule
    communication_barriers: Optional[CommunicationBarriers] = None


@dataclass
class GaussianReference:
    """src/variational/GaussianReference.jl:4-28: mean-field Gaussian reference of the variational leg, re-fitted
    after every round >= first_tuning_round from the target chains' online mean / variance."""
    first_tuning_round: int = 6
    mean: Optional[np.ndarray] = None
    standard_deviation: Optional[np.ndarray] = None

    def activate(self, it: "Iterators") -> bool:          # :16-18
        return it.round >= self.first_tuning_round

    def update_reference(self, rr) -> None:               # :22-28
        self.mean = np.asarray(rr.online_mean, dtype=np.float64).copy()
        self.standard_deviation = np.sqrt(np.asarray(rr.online_var, dtype=np.float64))


@dataclass
class StabilizedPT:
    """src/tempering/StabilizedPT.jl:8-66: a fixed and a variational leg sharing the target.  Global chains
    1..n_var are the variational leg (reference -> target), n_var+1..N the fixed leg reversed (target -> reference)
    (:96-116); the swap graph is the ordinary even/odd graph over all N chains (VariationalDEO.jl, OddEven.jl:16-48)."""
    fixed_leg: NonReversiblePT
    variational_leg: NonReversiblePT

    @property
    def n_var(self) -> int:
        return self.variational_leg.schedule.n_chains

    @property
    def n_fixed(self) -> int:
        return self.fixed_leg.schedule.n_chains

    @property
    def communication_barriers(self):                     # global_barrier(::StabilizedPT): the fixed leg's (:131)
        return self.fixed_leg.communication_barriers


def tempering_parameters(tempering) -> np.ndarray:
    """The annealing parameter of every global chain: `schedule.grids`, or for two legs
    vcat(variational leg, reverse(fixed leg)) (concatenate_log_potentials, StabilizedPT.jl:63-65)."""
    if isinstance(tempering, StabilizedPT):
        return np.concatenate([tempering.variational_leg.schedule.grids, tempering.fixed_leg.schedule.grids[::-1]])
    return tempering.schedule.grids


def create_tempering(inputs: "Inputs"):
    """tempering.jl:65-71."""
    if inputs.n_chains == 0 or inputs.n_chains_variational == 0:
        return NonReversiblePT(equally_spaced_schedule(inputs.n_chains_total))
    return StabilizedPT(NonReversiblePT(equally_spaced_schedule(inputs.n_chains)),
                        NonReversiblePT(equally_spaced_schedule(inputs.n_chains_variational)))


@dataclass
class Shared:
    """src/pt/Shared.jl:12-41."""
    iterators: Iterators
    tempering: object       # NonReversiblePT or StabilizedPT
    explorer: object


@dataclass
class PT:
    """src/pt/PT.jl:6-50."""
    inputs: Inputs
    shared: Shared
    engine: _capi.Engine
    reduced_recorders: Optional[ReducedRecorders] = None
    round_log: List[dict] = field(default_factory=list)

    def close(self):
        self.engine.close()


def create_pt(inputs: Inputs) -> PT:
    """PT(inputs) (PT.jl:46-51): Shared + create_replicas (replicas.jl:65-99)."""
    comm = inputs.comm
    cfg = inputs.target.engine_config()
    n_total = inputs.n_chains_total
    kw = dict(n_chains=n_total, seed=inputs.seed, rank=comm.rank, world_size=comm.world_size,
              device=inputs.device, recorder_order=inputs.recorder_order, **cfg)
    if inputs.n_chains_variational > 0:     # only then: engine factories of single-leg callers need not know the keyword
        kw["n_chains_variational"] = inputs.n_chains_variational
    if inputs.engine_factory is not None:
        engine = inputs.engine_factory(**kw)
    else:
        engine = _capi.Engine(inputs.engine_lib or _capi.EngineLib(), **kw)
    layout = shard_layout(n_total, comm.world_size, inputs.n_chains_variational if inputs.n_chains > 0 else 0)
    assert (engine.first_chain, engine.n_local) == layout[comm.rank], "engine shard geometry disagrees with LoadBalance"
    comm.connect_neighbours(engine)
    shared = Shared(Iterators(), create_tempering(inputs), inputs.explorer)
    engine.init_replicas()
    return PT(inputs, shared, engine)


def run_one_round(pt: PT) -> ReducedRecorders:
    """run_one_round! (pigeons.jl:46-55) — one engine call for the whole round,
    then reduce_recorders! across shards (recorders.jl:88-120)."""
    it = pt.shared.iterators
    eng = pt.engine
    dim = pt.inputs.target.dim
    eng.set_schedule(tempering_parameters(pt.shared.tempering))
    var = pt.inputs.variational
    if pt.inputs.n_chains_variational > 0:      # the path of the variational leg (update_path_variational, variational.jl:36-40)
        if var is not None and var.mean is not None:
            eng.set_variational(var.mean, var.standard_deviation)
        else:
            eng.set_variational(None, None)
    if pt.shared.explorer is None:
        eng.set_explorer(kind=_capi.EXPLORER_NONE)
    else:
        eng.set_explorer(**pt.shared.explorer.engine_params(dim))
    rec = set(pt.inputs.record)
    res = eng.run_round(n_scans_in_round(it),
                        log_index_process=index_process in rec,
                        log_swaps=swap_trace in rec,
                        log_target_trace=traces in rec)
    merged = merge_round_results(pt.inputs.comm, res, pt.inputs.n_chains_total, dim,
                                 pt.inputs.n_chains_variational if pt.inputs.n_chains > 0 else 0)
    return merged


def adapt(pt: PT, rr: ReducedRecorders) -> PT:
    """adapt (pigeons.jl:152-162): adapt_tempering (NonReversiblePT.jl:52-66) then adapt_explorer."""
    temp = pt.shared.tempering
    var = pt.inputs.variational

    def adapt_leg(leg: NonReversiblePT, pair_idx, variational) -> NonReversiblePT:
        """adapt_tempering(::NonReversiblePT, ..., chain_indices) (NonReversiblePT.jl:52-66); pair_idx: 0-based index of
        the lower chain of every pair of the leg, from its reference to the chain before its target."""
        if leg.schedule.n_chains == 1:
            return leg
        if variational is not None and variational.activate(pt.shared.iterators):     # update_path_if_needed, variational.jl:28-34
            variational.update_reference(rr)
        acc = np.where(np.asarray(rr.swap_n)[pair_idx] > 0, np.asarray(rr.swap_mean)[pair_idx], 0.5)   # adaptation.jl:103-112
        rej = 1.0 - acc
        return NonReversiblePT(optimal_schedule(rej, leg.schedule), communication_barriers(rej, leg.schedule.grids))

    if isinstance(temp, StabilizedPT):      # StabilizedPT.jl:51-61
        n_var, n = temp.n_var, temp.n_var + temp.n_fixed
        new_var = adapt_leg(temp.variational_leg, np.arange(0, n_var - 1), var)
        new_fixed = adapt_leg(temp.fixed_leg, np.arange(n - 2, n_var - 1, -1), None)    # pairs (N-1,N), ..., (n_var+1,n_var+2)
        new_temp = StabilizedPT(new_fixed, new_var)
    else:
        n = temp.schedule.n_chains
        leg_var = var if pt.inputs.n_chains_variational > 0 else None
        new_temp = adapt_leg(temp, np.arange(0, n - 1), leg_var)
    explorer = pt.shared.explorer
    if isinstance(explorer, (AutoMALA, MALA, Compose, Mix)):
        explorer = explorer.adapt(rr)
    pt.shared = Shared(pt.shared.iterators, new_temp, explorer)
    pt.reduced_recorders = rr
    return pt


class ChecksFailed(AssertionError):
    """check_against_serial found a difference (src/pt/checks.jl:52-78)."""


def run_checks(pt: PT) -> None:
    """run_checks / check_against_serial (src/pt/checks.jl:5-78, the oracle of
    test/test_parallelism_invariance.jl): after round `checked_round`, rank 0 re-runs rounds 1..checked_round
    in ONE process on one device — a fresh engine of the same library holding the whole ladder — and the
    sharded run must agree with it field by field: every replica (state, chain <-> replica index, RNG
    position, round-trip state), the adapted schedule and the adapted explorer.  With one process this is
    the determinism check of the reference's serial re-run."""
    inputs, it = pt.inputs, pt.shared.iterators
    if inputs.checked_round <= 0 or it.round != inputs.checked_round:
        return
    comm = inputs.comm
    st = pt.engine.get_state()
    whole = {k: np.concatenate(comm.all_gather_array(np.ascontiguousarray(v)), axis=0) for k, v in st.items()}
    bad, failure = [], ""
    if comm.rank == 0:
        try:
            # the serial re-run uses the same engine library / factory as the run it checks
            fresh_var = None if inputs.variational is None else replace(inputs.variational, mean=None, standard_deviation=None)
            serial = replace(inputs, comm=SingleProcess(), n_rounds=inputs.checked_round, checked_round=0, show_report=False,
                             variational=fresh_var)
            ref = pigeons_pt(create_pt(serial))
            rs = ref.engine.get_state()
            bad = [k for k in whole if not np.array_equal(whole[k].reshape(rs[k].shape), rs[k])]
            if not np.array_equal(tempering_parameters(pt.shared.tempering), tempering_parameters(ref.shared.tempering)):
                bad.append("schedule")
            if pt.shared.explorer != ref.shared.explorer:
                bad.append("explorer")
            ref.close()
        except Exception as e:      # noqa: BLE001 - the other ranks wait in the gather below: tell them instead of hanging them
            failure = f"{type(e).__name__}: {e}"
    code = -1 if failure else len(bad)
    verdicts = comm.all_gather_array(np.array([code], dtype=np.int64))     # every rank learns rank 0's verdict
    v = int(verdicts[0][0])
    if v < 0:
        raise ChecksFailed(f"round {it.round}: the serial re-run on rank 0 failed" + (f" ({failure})" if failure else ""))
    if v != 0:
        raise ChecksFailed(f"round {it.round}: the run on {comm.world_size} process(es) differs from the serial run"
                           + (f" in {bad}" if bad else ""))


def pigeons_pt(pt: PT) -> PT:
    """pigeons(pt::PT) (pigeons.jl:12-28)."""
    it = pt.shared.iterators
    while it.round + 1 <= pt.inputs.n_rounds:       # next_round! (Iterators.jl:27-35)
        it.round += 1
        rr = run_one_round(pt)
        pt = adapt(pt, rr)
        pt.round_log.append(dict(round=it.round, n_scans=n_scans_in_round(it), kernel_ms=rr.kernel_ms,
                                 wall_s=rr.wall_s, global_barrier=global_barrier(pt),
                                 stepping_stone=stepping_stone(pt) if rr.has_swap_stats else float("nan"),
                                 n_round_trips=rr.n_round_trips))
        run_checks(pt)
        if pt.inputs.show_report:
            r = pt.round_log[-1]
            print(f"round {r['round']:3d}  scans {r['n_scans']:8d}  Λ {r['global_barrier']:.4g}  "
                  f"log(Z1/Z0) {r['stepping_stone']:.6g}  kernel {r['kernel_ms']:.3f} ms")
    return pt


def _target_fingerprint(target) -> tuple:
    """(kind, dim, scalar parameters) of the device target: enough to refuse a checkpoint of another model."""
    cfg = target.engine_config()
    return (int(cfg["target_kind"]), int(cfg["dim"]), tuple(float(v) for v in cfg.get("p", ())))


def write_checkpoint(pt: PT) -> dict:
    """write_checkpoint (src/pt/checkpoint.jl:110-145) for the harness: everything a later `resume` needs to
    continue the run bit for bit — `Shared` (round counter, schedule, adapted explorer) and the `Replica`s
    (state, chain <-> replica index, RNG position, round-trip state; `pgn_get_state`).  A plain dict of numpy
    arrays / dataclasses, not the reference's `.jls` wire format."""
    return dict(round=pt.shared.iterators.round, grids=tempering_parameters(pt.shared.tempering).copy(),
                tempering=copy.deepcopy(pt.shared.tempering), variational=copy.deepcopy(pt.inputs.variational),
                communication_barriers=pt.shared.tempering.communication_barriers, explorer=pt.shared.explorer,
                replicas=pt.engine.get_state(), n_chains=pt.inputs.n_chains_total, seed=pt.inputs.seed,
                world_size=pt.inputs.comm.world_size, rank=pt.inputs.comm.rank, first_chain=pt.engine.first_chain,
                n_local=pt.engine.n_local, dim=pt.inputs.target.dim, target_config=_target_fingerprint(pt.inputs.target))


def resume(checkpoint: dict, inputs: Inputs) -> PT:
    """PT(exec_folder) + pigeons(pt) (src/pt/PT.jl:56-92, checkpoint.jl:18-60): rebuild the PT from a
    checkpoint and run rounds `checkpoint round + 1 .. inputs.n_rounds`.  `inputs` supplies what a
    checkpoint does not serialise in the reference either (the target and the engine handle are rebuilt,
    `ext/PigeonsBridgeStanExt/interface.jl:27-48`)."""
    if inputs.n_chains_total != checkpoint["n_chains"] or inputs.seed != checkpoint["seed"]:
        raise ValueError("the checkpoint was written by a run with a different n_chains / seed")
    if checkpoint.get("target_config") is not None and checkpoint["target_config"] != _target_fingerprint(inputs.target):
        raise ValueError("the checkpoint was written for a different target")
    pt = create_pt(inputs)
    # a checkpoint holds ONE shard: it can only be loaded by the same shard of the same layout
    for key, have in (("world_size", inputs.comm.world_size), ("rank", inputs.comm.rank),
                      ("first_chain", pt.engine.first_chain), ("n_local", pt.engine.n_local), ("dim", inputs.target.dim)):
        if key in checkpoint and checkpoint[key] != have:
            pt.close()
            raise ValueError(f"the checkpoint was written by a shard with {key} = {checkpoint[key]}, this one has {have}")
    st = checkpoint["replicas"]
    pt.engine.set_state(x=st["x"] if st["x"].size else None, replica_index=st["replica_index"],
                        rng_counter=st["rng_counter"], round_trip_state=st["round_trip_state"])
    if checkpoint.get("tempering") is not None:
        tempering = copy.deepcopy(checkpoint["tempering"])
    else:
        sched = Schedule(np.asarray(checkpoint["grids"], dtype=np.float64).copy())
        tempering = NonReversiblePT(sched, checkpoint["communication_barriers"])
    if checkpoint.get("variational") is not None:
        pt.inputs.variational = copy.deepcopy(checkpoint["variational"])
    pt.shared = Shared(Iterators(round=checkpoint["round"]), tempering, checkpoint["explorer"])
    return pigeons_pt(pt)


def pigeons(**kwargs) -> PT:
    """pigeons(; target, n_chains, explorer, ...) (src/api.jl:16-19)."""
    return pigeons_pt(create_pt(Inputs(**kwargs)))


# ---- post-processing ---------------------------------------------------------
def stepping_stone_pair(pt: PT):
    """src/evidence/stepping_stone.jl:28-43."""
    rr = pt.reduced_recorders
    e1 = 0.0
    e2 = 0.0
    temp = pt.shared.tempering
    # two legs: only the pairs inside the variational leg (stepping_stone_keys, stepping_stone.jl:50-66)
    n_pairs = temp.n_var - 1 if isinstance(temp, StabilizedPT) else pt.inputs.n_chains_total - 1
    for i in range(n_pairs):
        if rr.swap_n[i] > 0:
            e1 += rr.logsum_fwd[i] - math.log(rr.swap_n[i])
            e2 += rr.logsum_bwd[i] - math.log(rr.swap_n[i])
    return (e1, -e2)


def stepping_stone(pt: PT) -> float:
    """src/evidence/stepping_stone.jl:9-18."""
    p = stepping_stone_pair(pt)
    if not math.isfinite(p[0]):
        return p[1]
    if not math.isfinite(p[1]):
        return p[0]
    return (p[0] + p[1]) / 2.0


def global_barrier(pt: PT) -> float:        # NonReversiblePT.jl:74
    cb = pt.shared.tempering.communication_barriers
    if cb is None:      # a single chain (or no round run yet): the reference leaves the tempering untouched
        return float("nan")
    return cb.globalbarrier


def global_barrier_variational(pt: PT) -> float:      # StabilizedPT.jl:133
    temp = pt.shared.tempering
    if not isinstance(temp, StabilizedPT):
        raise TypeError("global_barrier_variational needs two legs")
    cb = temp.variational_leg.communication_barriers
    return float("nan") if cb is None else cb.globalbarrier


def n_round_trips(pt: PT) -> int:           # RoundTripRecorder.jl:23
    return pt.reduced_recorders.n_round_trips


def n_tempered_restarts(pt: PT) -> int:     # RoundTripRecorder.jl:21
    return pt.reduced_recorders.n_tempered_restarts


def sample_array(pt: PT) -> np.ndarray:
    """process_sample.jl:19-32 restricted to the target chain: [n_scans, d] of the last round."""
    tr = pt.reduced_recorders.target_trace
    if tr is None:
        raise ValueError("record=[traces] was not requested")
    return tr
