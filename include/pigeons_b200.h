/*
 * pigeons_b200.h — C ABI of the B200-native parallel-tempering scan engine.
 *
 * This is the drop-in boundary for the ONE hot path of Pigeons.jl that this
 * repository accelerates: the inner PT scan loop
 *     run_one_round!            (reference src/pt/pigeons.jl:46-55)
 *       explore!(all replicas)  (src/pt/pigeons.jl:82-132)
 *       communicate! / swap!    (src/pt/pigeons.jl:64-69, src/swap/swap.jl:6-26)
 * Host orchestration (rounds, adaptation, reports, checkpoints) stays in the
 * host language and calls `pgn_run_round` once per round.
 *
 * Conventions (modelled on the only FFI the reference has, the BridgeStan
 * ccall in ext/PigeonsBridgeStanExt/interface.jl:118-183):
 *   - every entry point returns an int rc, 0 = OK;
 *   - the last argument is `char** err`; on rc != 0 it receives a malloc'd
 *     message owned by the library, released with pgn_free_string();
 *   - handles are opaque pointers; all output buffers are caller-allocated;
 *   - one host thread drives a handle ("funneled", cf.
 *     src/mpi_utils/misc_mpi_utils.jl:13-19); calls return after stream sync;
 *   - no exceptions cross the boundary; there is NO CPU fallback: every
 *     compute entry point fails with PGN_ERR_NO_DEVICE when no sm_100 GPU is
 *     usable.
 *
 * Plain C: pointers and sizes only, no torch / C++ types.
 */
#ifndef PIGEONS_B200_H
#define PIGEONS_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define PGN_ABI_VERSION 4

/* ---- return codes ------------------------------------------------------- */
#define PGN_OK 0
#define PGN_ERR_INVALID 1        /* bad argument / unsupported configuration    */
#define PGN_ERR_NO_DEVICE 2      /* no usable CUDA device (no CPU fallback)      */
#define PGN_ERR_CUDA 3           /* CUDA runtime error                           */
#define PGN_ERR_NAN_RATIO 4      /* NaN log-unnormalized ratio in swap           */
                                 /*   (src/log_potentials/log_potentials.jl:47-49) */
#define PGN_ERR_BAD_DENSITY 5    /* non-finite density after a slice move        */
                                 /*   (src/explorers/SliceSampler.jl:35-37,52-59) */
#define PGN_ERR_SLICE_MAX_ITER 6 /* shrink loop exceeded max_iter (:179-184)     */
#define PGN_ERR_STEP_UNDERFLOW 7 /* autoMALA step size shrank to 0               */
                                 /*   (src/explorers/AutoMALA.jl:240-242)        */
#define PGN_ERR_NOT_POSITIVE 8   /* autoMALA started at zero density (:127)      */
#define PGN_ERR_TIMEOUT 9        /* neighbour hand-shake spin limit exceeded     */

/* ---- closed family of device targets ------------------------------------ */
/* (arbitrary host log_potential callables cannot run on the device)          */
#define PGN_TARGET_TOY_MVN 1      /* src/paths/ScaledPrecisionNormalPath.jl:5-78, src/targets/toy_mvn_target.jl */
#define PGN_TARGET_FUNNEL 2       /* test/supporting/dimensional-analysis.jl:33-47, ref N(0, s^2 I)              */
#define PGN_TARGET_GMM 3          /* DistributionLogPotential(MixtureModel(MvNormal...)), test/test_auto_mala.jl:126-132 */
#define PGN_TARGET_ISING 4        /* examples/ising.jl:6-117                                                      */
#define PGN_TARGET_LOGREG 5       /* analytic-gradient target pattern, test/test_custom_gradient.jl:1-33          */
#define PGN_TARGET_TEST_SWAPPER 6 /* src/swap/pair_swapper.jl:100-149                                             */
#define PGN_TARGET_MIXED 7        /* product of Bernoulli, Binomial and Normal coordinates: the mixed Bool / Integer /
                                     Float state of test/test_slice_sampler.jl:56-75 (SliceSampler.jl:65-86,136-142,189) */
#define PGN_TARGET_UNID 8         /* the unidentifiable product of test/test_DistributionLogPotential.jl:7-21 (and of
                                     toy_turing_unid_target, ext/PigeonsDynamicPPLExt/toy_examples.jl:17-19): dim 2,
                                     l(p1, p2) = s log(p1 p2) + (n - s) log1p(-p1 p2) on [0,1]^2, -Inf outside;
                                     reference Uniform(0,1)^2; p[0] = n_trials n, p[1] = n_successes s; initial state
                                     (0.5, 0.5); SliceSampler only (no analytic gradient is built) */

/* ---- explorers ----------------------------------------------------------- */
#define PGN_EXPLORER_NONE 0             /* TestSwapper: step! is a no-op (pair_swapper.jl:140-141) */
#define PGN_EXPLORER_TOY 1              /* src/explorers/ToyExplorer.jl:5-14                       */
#define PGN_EXPLORER_SLICE 2            /* src/explorers/SliceSampler.jl:8-237                     */
#define PGN_EXPLORER_AUTOMALA 3         /* src/explorers/AutoMALA.jl:29-294                        */
#define PGN_EXPLORER_ISING_METROPOLIS 4 /* examples/ising.jl:91-117                                */
#define PGN_EXPLORER_MALA 5             /* src/explorers/MALA.jl:19-104 (fixed step size)          */
#define PGN_EXPLORER_COMPOSE 6          /* Compose(e1, e2, ...): every explorer in turn (src/explorers/Compose.jl:16-19)   */
#define PGN_EXPLORER_MIX 7              /* Mix(e1, e2, ...): one explorer, drawn uniformly from the replica's stream (Mix.jl:20-21) */
#define PGN_MAX_MIX 4                   /* explorers in a Mix (src/explorers/Mix.jl:7-21)          */

/* ---- preconditioners (src/explorers/Preconditioner.jl:7-77) --------------- */
#define PGN_PRECOND_IDENTITY 0
#define PGN_PRECOND_DIAGONAL 1
#define PGN_PRECOND_MIX_DIAGONAL 2

/* ---- recorder accumulation order (pgn_config.recorder_order) --------------- */
#define PGN_RECORDERS_PER_REPLICA 0
#define PGN_RECORDERS_PER_CHAIN 1

typedef struct pgn_handle pgn_handle;

/* Static configuration of one engine instance (one GPU / one rank).
 * Mirrors the fields of Inputs that the scan path reads
 * (src/pt/Inputs.jl:9-102: target, seed, n_chains) plus the sharding rank. */
typedef struct pgn_config {
  int32_t abi_version;   /* PGN_ABI_VERSION */
  int32_t target_kind;   /* PGN_TARGET_* */
  int32_t dim;           /* state dimension d (Ising: L*L spins; TestSwapper: 0) */
  int32_t n_chains;      /* global number of chains N */
  int64_t seed;          /* Inputs.seed */
  int32_t rank;          /* this shard, 0-based */
  int32_t world_size;    /* number of shards P; chains are split contiguously by
                            the LoadBalance rule (src/mpi_utils/LoadBalance.jl:70-73,119-128) */
  int32_t device;        /* CUDA device ordinal */
  int32_t n_modes;       /* GMM: number of mixture components K */
  /* family-specific scalars:
   *  TOY_MVN : p[0]=precision0, p[1]=precision1
   *  FUNNEL  : p[0]=sigma_y (3.0), p[1]=log(sigma_y), p[2]=1/sigma_y^2,
   *            p[3]=sigma_ref, p[4]=log(sigma_ref), p[5]=1/sigma_ref^2
   *  GMM     : p[0]=sigma_mode, p[1]=d*log(sigma_mode)+d/2*log(2pi), p[2]=1/sigma_mode^2,
   *            p[3]=sigma_ref, p[4]=log(sigma_ref), p[5]=1/sigma_ref^2
   *  ISING   : p[0]=beta_model, p[1]=L
   *  LOGREG  : p[0]=n_data, p[3]=sigma_ref, p[4]=log(sigma_ref), p[5]=1/sigma_ref^2
   *  TEST_SWAPPER : p[0]=constant_swap_accept_pr
   *  MIXED   : p[0]=n_bool, p[1]=n_int, p[2]=binomial n, p[3]=sigma_ref, p[4]=log(sigma_ref), p[5]=1/sigma_ref^2;
   *            state = n_bool Bool coordinates (0.0 / 1.0), then n_int Integer coordinates, then dim - n_bool - n_int
   *            Float coordinates; target = Bernoulli(p1) x Binomial(n, q1) x Normal(0, 1), reference = Bernoulli(p0) x
   *            Binomial(n, q0) x Normal(0, sigma_ref); `means` holds [log p0, log(1-p0), log p1, log(1-p1), log q0,
   *            log(1-q0), log q1, log(1-q1), p0, q0, log C(n,0..n)] (n_modes = 10 + n + 1 entries)                      */
  double p[8];
  const double* means;        /* GMM: [K][d] row-major, host pointer (copied)   */
  const double* log_weights;  /* GMM: [K] log mixture weights (copied)          */
  const double* data_x;       /* LOGREG: [n_data][d] row-major (copied)         */
  const double* data_y;       /* LOGREG: [n_data] labels in {0,1} (copied)      */
  /* How the float statistics of a round are accumulated (their counts and every integer are the same either way):
   *  PGN_RECORDERS_PER_REPLICA (0, default): as the reference does — every replica accumulates its own recorders,
   *    keyed by chain / pair, and the round ends with reduce_recorders!: a binary-tree merge over replica indices
   *    (src/recorders/recorders.jl:88-120, src/mpi_utils/Entangler.jl:214-277).  Needs n_chains x n_local entries
   *    of device memory.
   *  PGN_RECORDERS_PER_CHAIN (1): one accumulator per chain / pair, fitted in scan order; O(n_local) memory.  The
   *    means differ from the reference's at rounding level. */
  int32_t recorder_order;
  int32_t n_chains_variational; /* 0: one leg (NonReversiblePT).  k > 0: chains 1..k form the variational leg; with
                                 * k < n_chains the ladder has TWO legs (StabilizedPT.jl:37-66, VariationalDEO.jl):
                                 *   chain 1 (variational reference) .. chain k (target) | chain k+1 (target) .. chain
                                 *   n_chains (fixed reference);  is_reference = {1, n_chains}, is_target = {k, k+1}
                                 * (VariationalDEO.jl:19-20); pgn_set_schedule takes the concatenated parameters
                                 * vcat(variational leg, reverse(fixed leg)) (StabilizedPT.jl:63-65).  Two legs need
                                 * recorder_order = PGN_RECORDERS_PER_REPLICA; with several shards chain k+1 joins
                                 * chain k's shard when the balanced split would separate them (pgn_local_range). */
} pgn_config;

/* Explorer parameters; mirrors the @kwdef explorer structs
 * (SliceSampler.jl:8-20, AutoMALA.jl:29-68) after host-side adaptation. */
typedef struct pgn_explorer_params {
  int32_t kind;               /* PGN_EXPLORER_* */
  /* SliceSampler */
  double slice_w;             /* 10.0 */
  int32_t slice_p;            /* 20   */
  int32_t slice_n_passes;     /* 3    */
  int32_t slice_max_iter;     /* 1024 */
  /* AutoMALA */
  int32_t n_refresh;          /* base_n_refresh * ceil(Int, d^exponent_n_refresh), host-computed (AutoMALA.jl:122) */
  double step_size;           /* explorer.step_size */
  int32_t precond_kind;       /* PGN_PRECOND_* */
  double mix_p0;              /* MixDiagonalPreconditioner.p0 (1//3) */
  double mix_p01;             /* p0 + p1 (2//3) */
  const double* std_devs;     /* estimated_target_std_deviations [d] or NULL (round 1) */
  /* IsingMetropolis */
  int32_t ising_n_steps;      /* 3 */
  /* Mix(AutoMALA(...), AutoMALA(...), ...) (src/explorers/Mix.jl:20-21: one of the explorers, drawn uniformly
     from the replica's stream, performs the step; test/test_parallelism_invariance.jl:14-18 mixes autoMALA
     kernels that differ in their preconditioner).  n_mix <= 1: the single autoMALA described above.
     n_mix in 2..PGN_MAX_MIX: variant v uses (mix_n_refresh[v], mix_step_size[v], mix_precond_kind[v],
     mix_variant_p0[v], mix_variant_p01[v]); std_devs is shared (every variant adapts from the same
     recorders, Mix.jl:14-17). */
  /* kind = PGN_EXPLORER_COMPOSE / PGN_EXPLORER_MIX: n_steps explorers (1..PGN_MAX_MIX) of kinds step_kind[s] in
     {PGN_EXPLORER_TOY, PGN_EXPLORER_SLICE, PGN_EXPLORER_AUTOMALA, PGN_EXPLORER_MALA}; an autoMALA / MALA step s takes
     its parameters from mix_n_refresh[s], mix_step_size[s], mix_precond_kind[s], mix_variant_p0[s], mix_variant_p01[s];
     SliceSampler steps share the slice_* fields; std_devs is shared (every explorer adapts from the same recorders).
     Vector targets on the register-resident kernels (d <= 128). */
  int32_t n_steps;
  int32_t step_kind[PGN_MAX_MIX];
  int32_t n_mix;
  int32_t mix_n_refresh[PGN_MAX_MIX];
  int32_t mix_precond_kind[PGN_MAX_MIX];
  double mix_step_size[PGN_MAX_MIX];
  double mix_variant_p0[PGN_MAX_MIX];
  double mix_variant_p01[PGN_MAX_MIX];
} pgn_explorer_params;

/* Outputs of one round.  Every pointer is caller-allocated; optional logs may
 * be NULL.  "local chain" i (0-based) is global chain first_chain + i (1-based
 * numbering as in the reference).  Pair slot i describes the pair
 * (chain, chain+1) whose LOWER chain is local chain i (the replica holding the
 * lower chain records, src/swap/swap.jl:119-121). */
typedef struct pgn_round_out {
  /* swap_acceptance_pr = GroupBy((i,i+1) -> Mean)   (src/recorders/recorder.jl:60) */
  int64_t* swap_n;         /* [n_local] */
  double* swap_mean;       /* [n_local] */
  /* log_sum_ratio = GroupBy((i,j) -> LogSum)        (recorder.jl:87, LogSum.jl:10-18) */
  double* logsum_fwd;      /* [n_local] key (i,i+1); -inf when swap_n == 0 */
  double* logsum_bwd;      /* [n_local] key (i+1,i) */
  /* explorer recorders, keyed by chain              (recorder.jl:67,74; AutoMALA.jl:277,294) */
  int64_t* expl_acc_n;     /* [n_local] */
  double* expl_acc_mean;   /* [n_local] explorer_acceptance_pr */
  int64_t* expl_n_steps;   /* [n_local] explorer_n_steps (Sum) */
  int64_t* am_n;           /* [n_local] */
  double* am_mean;         /* [n_local] am_factors */
  int64_t* rev_n;          /* [n_local] */
  double* rev_mean;        /* [n_local] reversibility_rate */
  /* round_trip                                       (RoundTripRecorder.jl:43-54) */
  int64_t n_tempered_restarts;
  int64_t n_round_trips;
  /* _transformed_online / online at the target chain (OnlineStateRecorder.jl:87-110);
   * filled only by the shard owning chain N */
  int64_t online_n;
  double* online_mean;     /* [d] */
  double* online_var;      /* [d] OnlineStats Variance value (bessel-corrected) */
  /* optional event logs */
  int32_t* index_process;  /* [n_scans][n_local] replica_index (1-based) sitting at each chain, recorded before the swap (swap.jl:110) */
  double* swap_lr;         /* [n_scans][n_local] SwapStat.log_ratio of the replica at each chain (NaN for self-partnered) */
  double* swap_u;          /* [n_scans][n_local] SwapStat.uniform */
  uint8_t* swap_accept;    /* [n_scans][n_local] 1 if that chain's pair swapped */
  double* target_trace;    /* [n_scans][d] state at chain N after explore (traces recorder, recorder.jl:27-43);
                            * two legs: [n_scans][2][d], target chains k then k+1 */
  /* work counters */
  int64_t n_density_points;   /* distinct (state) points at which the device evaluated ref+target densities */
  int64_t n_ref_equiv_evals;  /* log_potential / logdensity[_and_gradient] calls the reference code path would have made */
  double kernel_ms;           /* CUDA-event duration of the scan kernel(s) of this round */
  double gemm_ms;             /* LOGREG: device time inside the two FP64 GEMMs of this round (0 otherwise) */
  int64_t batch_steps;        /* LOGREG: number of batched density/gradient evaluations (0 otherwise) */
  int64_t n_launches;         /* kernels of this library launched for the round (1 for the persistent scan kernels) */
  int64_t active_columns;     /* LOGREG: sum over batch steps of the chains whose pending point was evaluated */
  int64_t gemm_columns;       /* LOGREG: sum over batch steps of the columns the GEMMs multiplied (tiles of 128) */
} pgn_round_out;

/* Replica state for checkpoint / inspection, in chain order:
 * row i = the replica currently at local chain i. */
typedef struct pgn_replica_state {
  double* x;               /* [n_local][d] (Ising: 0.0/1.0 per spin) */
  int32_t* replica_index;  /* [n_local] 1-based */
  uint64_t* rng_counter;   /* [n_local] Philox draws consumed by that replica */
  int32_t* round_trip_state; /* [n_local] RoundTripRecorder.state (0,1,2) */
} pgn_replica_state;

typedef struct pgn_device_info_t {
  int32_t sm_major, sm_minor, n_sms;
  int64_t global_mem_bytes;
  int32_t max_resident_chains;  /* co-resident warp capacity of the scan kernel */
  char name[128];
} pgn_device_info_t;

/* lifecycle */
int pgn_abi_version(void);
int pgn_create(const pgn_config* cfg, pgn_handle** out, char** err);
int pgn_destroy(pgn_handle* h);
void pgn_free_string(char* s);
int pgn_device_info(int device, pgn_device_info_t* out, char** err);

/* shard geometry: first global chain (1-based) and number of local chains */
int pgn_local_range(const pgn_handle* h, int32_t* first_chain, int32_t* n_local);

/* tempering.schedule.grids (src/schedules/Schedule.jl:5-44): all N betas */
int pgn_set_schedule(pgn_handle* h, const double* beta, int32_t n, char** err);
/* shared.explorer after adapt_explorer (src/explorers/AutoMALA.jl:70-79) */
int pgn_set_explorer(pgn_handle* h, const pgn_explorer_params* ep, char** err);

/* create_replicas / initialization (src/replicas/replicas.jl:87-99):
 * replica i starts at chain i with its own RNG stream and the target's
 * default initial state. */
int pgn_init_replicas(pgn_handle* h, char** err);
int pgn_get_state(pgn_handle* h, pgn_replica_state* out, char** err);
int pgn_set_state(pgn_handle* h, const pgn_replica_state* in, char** err);

/* run_one_round! (src/pt/pigeons.jl:46-55): n_scans x {explore!, swap!}. */
int pgn_run_round(pgn_handle* h, int64_t n_scans, pgn_round_out* out, char** err);

/* Parity entry points mirroring the log_potential callable contract
 * (src/log_potentials/log_potential.jl:1-14, InterpolatedLogPotential.jl:10-17)
 * and LogDensityProblems.logdensity_and_gradient as used by autoMALA
 * (src/explorers/BufferedAD.jl:89-111).  x: [n_points][d], beta: [n_points]. */
int pgn_log_potential(pgn_handle* h, const double* x, int32_t n_points,
                      const double* beta, double* out, char** err);
int pgn_logdensity_and_gradient(pgn_handle* h, const double* x, int32_t n_points,
                                const double* beta, double* logdens, double* grad,
                                char** err);

/* GaussianReference (src/variational/GaussianReference.jl:4-54): a mean-field Gaussian as the reference of the
 * variational leg (chains 1..n_chains_variational): log density sum_i -0.5 log(2 pi sd_i^2) - (x_i - mean_i)^2 / (2 sd_i^2)
 * (:45-53), gradient -(x - mean) / sd^2 (:72-80), sample_iid! randn * sd_i + mean_i (:30-37).  mean, sd: [dim] — what
 * update_reference! (:22-28) computed from the target-chain online statistics; mean == NULL switches back to the fixed
 * reference (activate_variational false, :16-18).  Takes effect at the next round.  Vector targets with an
 * InterpolatingPath on the register-resident kernels (FUNNEL, GMM, UNID; d <= 128). */
int pgn_set_variational(pgn_handle* h, const double* mean, const double* sd, char** err);

/* hamiltonian_dynamics!(target_log_potential, diag_precond, state, momentum, step_size, n_steps)
 * (src/explorers/hamiltonian_dynamics.jl:39-84) as n_steps applications of leap_frog! (:86-93) — the device's own integrator,
 * the code the autoMALA / MALA kernels run (the reference's n-step form merges the inner half-steps, :73-76: the same map up to
 * rounding) — from (x, p) at inverse temperature beta.  diag_precond: [d], or NULL for the identity.  step_size may be
 * negative.  Parity / property entry point (test/test_auto_mala.jl:51-85: forward, then flip the momentum or the step, returns
 * to the start).  Vector targets with a gradient on the register-resident kernels (d <= 128).
 * x, p, x_out, p_out: [n_points][d]. */
int pgn_hamiltonian_dynamics(pgn_handle* h, const double* x, const double* p, int32_t n_points, const double* beta,
                             const double* diag_precond, double step_size, int32_t n_steps, double* x_out, double* p_out,
                             char** err);

/* Multi-GPU (one process per GPU): neighbour mailboxes are peer-mapped with
 * CUDA IPC.  Replaces the role of Entangler.transmit! (src/mpi_utils/Entangler.jl:133-184)
 * for the single boundary pair per shard.  handle64: 64-byte cudaIpcMemHandle_t. */
int pgn_ipc_export(pgn_handle* h, void* handle64, char** err);
int pgn_ipc_attach(pgn_handle* h, int32_t side /*0=left,1=right*/, const void* handle64, char** err);
/* same-process multi-GPU (tests): attach another handle's mailbox directly */
int pgn_peer_attach(pgn_handle* h, int32_t side, pgn_handle* neighbour, char** err);

/* Measured FP64 FMA throughput of the device (register-resident DFMA loop, all SMs):
 * the roofline denominator for the GEMM-shaped LOGREG path. */
int pgn_measure_fp64_peak(int32_t device, double* tflops, char** err);

/* Numerics self-test hook: evaluates the device elementary functions / RNG
 * (op: 0 exp, 1 log, 2 cospi, 3 normal_at(ctr), 4 uniform_at(ctr),
 * 5 exponential_at(ctr), 6 logaddexp(in[2i],in[2i+1])) on the GPU. */
int pgn_test_math(int32_t device, int32_t op, const double* in, double* out, int64_t n,
                  int64_t seed, int32_t replica_index, char** err);

/* Probe of the FP64 tensor-core instruction mma.sync.m8n8k4.f64 (DMMA): n_trials
 * independent products D = A(8x4) B(4x8) + C(8x8), row-major inputs, one warp each.
 * Used to establish the accumulation order of the hardware before it may replace
 * the SIMT GEMM of the LOGREG path (whose summation order is part of the spec). */
int pgn_test_dmma(int32_t device, const double* a, const double* b, const double* c, double* d_out,
                  int32_t n_trials, char** err);

#ifdef __cplusplus
}
#endif
#endif /* PIGEONS_B200_H */
