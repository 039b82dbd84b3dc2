"""pigeons.jl_b200 — B200-native engine for the inner scan of non-reversible
parallel tempering, behind the Pigeons.jl `pigeons(...)` / `Inputs` /
`log_potential` / `explorer` surface.

The directory name contains a dot, so import it through the alias module at
the repository root:  ``import pigeons_jl_b200 as pg``.

Only the scan path (explore! + DEO swap!) runs on the device; everything in
this package is host-side glue around the C ABI in include/pigeons_b200.h.
"""
from ._capi import Engine, EngineError, EngineLib, default_library_path          # noqa: F401
from .distributed import LoadBalance, SingleProcess, ThreadComm, ThreadGroup, TorchDistributed, shard_layout   # noqa: F401
from .explorers import (MALA, AutoMALA, Compose, Mix, DiagonalPreconditioner, IdentityPreconditioner,  # noqa: F401
                        IsingMetropolis, MixDiagonalPreconditioner, SliceSampler, ToyExplorer)
from .pt import (PT, ChecksFailed, GaussianReference, Inputs, Iterators, NonReversiblePT, Shared, StabilizedPT, adapt,   # noqa: F401
                 create_pt, create_tempering, global_barrier, global_barrier_variational, tempering_parameters,
                 index_process, n_round_trips, n_scans_in_round, n_tempered_restarts, online, pigeons, resume, write_checkpoint,
                 pigeons_pt, round_trip, run_checks, run_one_round, sample_array, stepping_stone, stepping_stone_pair,
                 swap_trace, traces)
from .recorders import ReducedRecorders                                            # noqa: F401
from .targets import (Funnel, GaussianMixture, IsingLogPotential, LogisticRegression, MixedProduct,  # noqa: F401
                      ScaledPrecisionNormalPath, TestSwapper, UnidentifiableProduct, eight_mode_mixture,
                      synthetic_logistic_regression, toy_mvn_target)
from .tempering import (MonotoneCubic, Schedule, communication_barriers, equally_spaced_schedule,  # noqa: F401
                        optimal_schedule, rejections)

__version__ = "0.1.0"
