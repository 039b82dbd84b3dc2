// pgn_kernels.cuh — the persistent PT scan kernel (explore! + DEO swap!) for sm_100a.
//
// One warp owns one CHAIN (one annealing parameter beta_c) for a whole round.
// The replica sitting at that chain — its state, RNG stream, replica index and
// round-trip state — lives in the warp's registers; HBM is touched once when
// the round starts and once when it ends.  A scan is
//
//   explore : sample_iid! at chain 1, else SliceSampler / autoMALA / IsingMetropolis
//             (src/pt/pigeons.jl:101-132), every lane owning the coordinates
//             c = lane, lane+32, ... and sums running through the canonical
//             warp butterfly;
//   swap    : the chain posts its SwapStat (log_ratio, uniform) + replica
//             payload into its mailbox slot, waits for its DEO partner's post
//             (src/swap/OddEven.jl:23-31), both sides take the identical
//             accept decision (src/swap/pair_swapper.jl:81-88) and on accept
//             adopt each other's replica.
//
// There is NO grid-wide barrier: DEO couples a chain only to its two
// neighbours, so warps synchronise pairwise through flag-in-data mailbox words
// (8-byte {payload32, tag32} words, relaxed stores and loads, no fences) in
// L2 (or, across GPUs, in the neighbour's peer-mapped mailbox over NVLink).
// All warps are co-resident (cooperative launch), so the spin-waits cannot
// deadlock; a spin limit turns any lost hand-shake into PGN_ERR_TIMEOUT.
#pragma once
#include <cstdint>

#include "../../include/pigeons_b200.h"
#include "pgn_numerics.cuh"

namespace pgn {

constexpr int KMAX_MODES = 8;      // GMM components held in registers
constexpr int MAIL_RINGS = 8;      // 2 round banks x 4 scans (see DESIGN.md "mailbox ring")
constexpr int MAIL_HDR_BYTES = 64; // [flag u64 | pad | MailHdr 32B]

struct MeanAcc {   // OnlineStatsBase.Mean (EqualWeight)
  long long n;
  double mu;
  template <bool COMPACT = false>
  __device__ __forceinline__ void fit(double x) { n += 1; mu = mu + div_<COMPACT>(1.0, (double)n) * (x - mu); }
};
struct LogSumAcc {  // src/recorders/LogSum.jl:1-24
  long long n;
  double value;
  template <bool COMPACT = false>
  __device__ __forceinline__ void fit(double y) { value = logaddexp_<COMPACT>(value, y); n += 1; }
};

// OnlineStatsBase._merge! for the statistics of this path (a statistic that was never fitted does not exist in the
// reference's GroupBy: merging with it is a copy)
__device__ __forceinline__ void merge_mean(MeanAcc& a, const MeanAcc& b) {
  if (b.n == 0) return;
  if (a.n == 0) { a = b; return; }
  a.n += b.n;
  a.mu = a.mu + ((double)b.n / (double)a.n) * (b.mu - a.mu);
}
__device__ __forceinline__ void merge_logsum(LogSumAcc& a, const LogSumAcc& b) {   // src/recorders/LogSum.jl:15-18
  if (b.n == 0) return;
  if (a.n == 0) { a = b; return; }
  a.value = logaddexp_(a.value, b.value);
  a.n += b.n;
}

// Per-replica recorders (PGN_RECORDERS_PER_REPLICA): what replica r has accumulated while it sat at local chain c lives
// in rec_table[(r - 1) * n_local + c].  The warps of a chain keep the entry of the replica they currently hold in
// registers and exchange it for the incoming replica's entry when a swap is accepted; the round ends with the
// reference's tree merge over replica indices (merge_recorders_kernel).
struct RecEntry {   // 96 bytes
  MeanAcc expl_acc, am, rev, swap_acc;
  LogSumAcc ls_fwd, ls_bwd;
};
struct OnEntry {    // OnlineStatsBase.Variance of one coordinate, as one replica accumulated it at the target chain
  long long n;
  double mu, s2;
};
// The tables are filled with "absent" entries when a round starts (init_recorder_tables_kernel): every count 0, every
// mean 0, every LogSum at -inf — an entry can be loaded into the registers as it is, no test on the load's result.

struct ChainStatsDev {   // one per local chain, written when the round ends
  long long swap_n; double swap_mean; double ls_fwd; double ls_bwd;
  long long expl_acc_n; double expl_acc_mean; long long n_steps;
  long long am_n; double am_mean; long long rev_n; double rev_mean;
  long long n_restarts, n_round_trips;
  long long n_points, n_ref_evals;
  long long trial_cycles, barrier_cycles, decide_cycles;   // diagnostics: autoMALA search phases of team warp 0
  long long explore_cycles, wait_cycles;   // diagnostics (PGN_TIMING_DUMP): SM clocks spent exploring / waiting for the swap partner
};

struct MailHdr {   // 32 bytes
  double lr, u;
  unsigned long long ctr;
  int replica_index, rt_state;
};

struct Params {
  // geometry
  int target_kind, d, d_pad, n_chains, first_chain, n_local;
  long long n_scans;
  unsigned int seed_lo, seed_hi;
  unsigned int epoch;
  int pool_refresh;        // refreshments the team's shared-memory momentum pool has room for (0: no pool)
  unsigned int tag_base;   // scans run by this handle before this round (mod 2^32): mailbox tag = tag_base + scan
  double p[8];
  int n_modes;
  const double* means;      // [K][d_pad]
  const double* log_w;      // [K]
  const double* beta;       // [N]
  // explorer
  double slice_w; int slice_p, slice_n_passes, slice_max_iter;
  int n_refresh; double step_size; int precond_kind; double mix_p0, mix_p01;
  int n_steps; int step_kind[PGN_MAX_MIX]; int program_is_mix;   // Compose / Mix program (step s: kind + mix_*[s])
  int n_mix;   // Mix of autoMALA kernels (Mix.jl:20-21): variant v = mix_*[v]
  int mix_n_refresh[PGN_MAX_MIX]; int mix_precond_kind[PGN_MAX_MIX];
  double mix_step_size[PGN_MAX_MIX], mix_variant_p0[PGN_MAX_MIX], mix_variant_p01[PGN_MAX_MIX];
  const double* std_devs;   // [d] or null
  int ising_n_steps;
  // replica state in HBM, chain order
  double* x;                // [n_local][d_pad]
  int* replica_index;
  unsigned long long* rng_ctr;
  int* rt_state;
  // mailboxes: slot-major, slot 0/1 = left/right ghost, 2.. = local chains
  char* mail; char* mail_left; char* mail_right;
  unsigned long long slot_bytes;
  // outputs
  ChainStatsDev* stats;
  double* online_mean; double* online_s2; long long* online_n;
  int* index_process; double* swap_lr; double* swap_u; unsigned char* swap_accept; double* target_trace;
  int* error_flag;
  unsigned long long timeout_ns;
  // per-replica recorders (null: one accumulator per chain, fitted in scan order)
  unsigned long long* progress;   // scans completed by the chains of this shard (all rounds): the hand-shake spin limit counts
                                  // time WITHOUT progress anywhere on the shard, not time since the wait began
  RecEntry* rec_table;     // [n_chains replicas][n_local]
  OnEntry* on_table;       // [n_chains replicas][d_pad], used by the shard owning the target chain(s)
  // legs (pgn_config.n_chains_variational): chains 1..n_var are the variational leg; two legs when 0 < n_var < n_chains
  int n_var;
  // mixed teams (VecChain<..., MIXED>): per block of two warps, [2 b] = local chain of warp 0, [2 b + 1] = local chain of warp 1,
  // -1 (no chain) or -2 (warp 1 is the second warp of warp 0's team)
  const int* block_map;
  const double* var_tab;   // GaussianReference, or null: [5][d_pad] mean | sd | -0.5 log(2 pi sd^2) | 1/(2 sd^2) | 1/sd^2
};
__device__ __forceinline__ bool two_legs(const Params& P) { return P.n_var > 0 && P.n_var < P.n_chains; }
// is_reference / is_target of the swap graph (DEO.jl:13-14, VariationalDEO.jl:19-20)
__device__ __forceinline__ bool chain_is_reference(const Params& P, int chain) {
  return (chain == 1 && P.n_chains > 1) || (two_legs(P) && chain == P.n_chains);
}
__device__ __forceinline__ bool chain_is_target(const Params& P, int chain) {
  return two_legs(P) ? (chain == P.n_var || chain == P.n_var + 1) : chain == P.n_chains;
}

__device__ __forceinline__ unsigned long long globaltimer_ns() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}
__device__ __forceinline__ void st_release_gpu(unsigned long long* p, unsigned long long v) {
  asm volatile("st.release.gpu.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ void st_release_sys(unsigned long long* p, unsigned long long v) {
  asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned long long ld_relaxed_sys(const unsigned long long* p) {
  unsigned long long v;
  asm volatile("ld.relaxed.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_relaxed_sys(unsigned long long* p, unsigned long long v) {
  asm volatile("st.relaxed.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
// Flag-in-data mailbox words (the scheme of NCCL's LL protocol): every 8-byte word carries 32 bits of
// payload and the 32-bit tag of the (round, scan) it belongs to.  An aligned 8-byte store is atomic, so
// a reader that sees the tag has the payload: posting and polling need no fence and no separate flag,
// locally (through L2) and across GPUs (peer-mapped stores over NVLink) alike.
__device__ __forceinline__ void ll_store(unsigned long long* p, unsigned int data, unsigned int tag) {
  st_relaxed_sys(p, ((unsigned long long)tag << 32) | (unsigned long long)data);
}
constexpr unsigned int LL_PAYLOAD_SPIN_LIMIT = 1u << 24;
// poll one word until it carries `tag`; returns false if it never does (the header of the same post was
// already seen, so this only guards against a broken peer)
__device__ __forceinline__ bool ll_load(const unsigned long long* p, unsigned int tag, unsigned int& data) {
  unsigned long long w = ld_relaxed_sys(p);
  unsigned int it = 0;
  while ((unsigned int)(w >> 32) != tag) {
    if (++it > LL_PAYLOAD_SPIN_LIMIT) { data = 0u; return false; }
    w = ld_relaxed_sys(p);
  }
  data = (unsigned int)w;
  return true;
}
constexpr int LL_HDR_WORDS = 8;   // lr (2), u (2), rng counter (2), replica_index, round-trip state
__device__ __forceinline__ void fence_acq_rel_sys() { asm volatile("fence.acq_rel.sys;" ::: "memory"); }
__device__ __forceinline__ void fence_acq_rel_gpu() { asm volatile("fence.acq_rel.gpu;" ::: "memory"); }

// ===========================================================================
// Vector-state chains: TOY_MVN / FUNNEL / GMM with ToyExplorer, SliceSampler, autoMALA
// ===========================================================================
// VAR: the chain may sit on the variational leg and then uses the GaussianReference tables instead of the fixed
// reference (separate instantiations, csrc/Makefile vec_*_var_*: the plain kernels carry none of this)
// MIXED: blocks of two warps that are EITHER one chain's team of two OR two independent single-warp chains (Params::block_map):
// the warp slots the register file leaves free go to the slowest chains (pgn_engine.cu, "mixed teams")
template <int TK, int CPL, int EX, bool VAR = false, bool MIXED = false>
struct VecChain {
  static constexpr bool kTestSwapper = false;
  static constexpr bool kMixed = MIXED;
  // barrier over the warps of THIS chain's team: the whole block, or (mixed blocks) a named barrier over the team's 32 W threads
  __device__ __forceinline__ void team_barrier() const {
    if constexpr (MIXED) {
      if (W > 1) asm volatile("bar.sync 1, %0;" ::"r"(W * 32) : "memory");
      else __syncwarp();
    } else {
      __syncthreads();
    }
  }
  bool var_ref;   // VAR only: this chain's reference is the Gaussian variational one
  // autoMALA runs with a TEAM of W warps per chain (block = team): the step-size search evaluates W
  // candidate steps per round, one per warp, and replays the reference's sequential decisions on the
  // results (see automala()).  Everything else is executed redundantly by all warps of the team on
  // identical data; only team warp 0 talks to the mailboxes and writes results.
  static constexpr bool kTeam = (EX == PGN_EXPLORER_AUTOMALA || EX == PGN_EXPLORER_COMPOSE);
  static constexpr bool kIsing = false;
  // one coordinate per lane: teams of up to 6 warps, two teams per SM (168 registers per thread)
  static constexpr int kMaxThreads = (kTeam && CPL == 1) ? 192 : 256;
  static constexpr int kMinBlocksPerSM = (kTeam && CPL == 1) ? 2 : 1;
  static constexpr bool kCompact = CPL > 1;                  // large kernels call divisions / generators out of line (instruction cache)
  static constexpr int SLOT_DOUBLES = 3 * CPL * 32 + 8;    // x1 | p1 | g1c | {a0, a1, lp1, h_after, diff, eps, -, -}
  static constexpr int TEAM_CTL_DOUBLES = 24 + CPL * 32;   // control words + shared copy of an adopted state
  static constexpr int POOL_DOUBLES = CPL * 32 + 8;        // per refreshment: momentum | a, b, u, |p|^2, log a, log b, -, -
  int tw, W, gen;
  long long t_trial, t_barrier, t_decide;   // diagnostics (clock64 deltas)
  int own1, own2, own3, own4;   // team warp keeping statistic j (j mod W)
  int e_prev;           // exponent chosen by this chain's previous forward search (candidate placement only)
  double* slots;        // [3][W][SLOT_DOUBLES] trial results, rotating buffers
  double* rng_pool;     // [n_refresh][d_pad + 8] a scan's momenta and bound uniforms, drawn by the team in one go (or null)
  const Params* P;
  const double* sm_means;
  int lane, d;
  double beta;
  double x[CPL];
  double e0, e1;        // densities at x: (S, -) for toy MVN, (l_ref, l_tgt) otherwise
  Rng rng;
  MeanAcc expl_acc, am, rev;
  long long n_steps, n_points, n_ref;
  int err;
  // target-chain online statistics (OnlineStatsBase.Variance per coordinate)
  double on_mu[CPL], on_s2[CPL];
  long long on_n;
  double pool;
  unsigned long long pool_base;
  bool pool_valid;

  static __host__ __device__ int target_smem_doubles(int d_pad) {
    return TK == PGN_TARGET_GMM ? KMAX_MODES * d_pad + KMAX_MODES : 0;
  }
  __device__ void set_team(int tw_, int W_, double* slots_, double* rng_pool_) { tw = tw_; W = W_; slots = slots_; gen = 0; rng_pool = rng_pool_;
    own1 = 1 % W_; own2 = 2 % W_; own3 = 3 % W_; own4 = 4 % W_;
  }
  // The running statistics do not feed back into the chain, so each is kept by ONE warp of the team
  // (the others would only repeat its divisions): statistic j lives in team warp j mod W.
  __device__ __forceinline__ int own(int j) const { return j == 1 ? own1 : (j == 2 ? own2 : (j == 3 ? own3 : (j == 4 ? own4 : 0))); }
  __device__ __forceinline__ double* slot(int buf, int w) const { return slots + ((size_t)buf * W + w) * SLOT_DOUBLES; }
  static __device__ void stage_shared(const Params& P, double* smem) {
    if (TK == PGN_TARGET_GMM) {
      // P.means is the staged layout [KMAX_MODES][d_pad] followed by KMAX_MODES log weights; components
      // beyond K are padded with weight exp(-inf) = 0: their terms add exact zeros, so the mode loops
      // run over all KMAX_MODES slots without predicates and give the same bits as K terms
      const int n = KMAX_MODES * P.d_pad + KMAX_MODES;
      for (int i = threadIdx.x; i < n; i += blockDim.x) smem[i] = P.means[i];
    }
  }

  __device__ __forceinline__ bool valid(int k) const { return k * 32 + lane < d; }
  // one standard normal at tick `ctr` of this replica's stream; with several coordinates per lane the
  // generator is called out of line (CPL unrolled copies of Philox + Box-Muller are ~15 KB of code)
  __device__ __forceinline__ double normal_tick(unsigned long long ctr) const {
    if (CPL > 1) return normal_at_ni(rng.key0, rng.key1, rng.c2, rng.c3, ctr);
    return normal_at(rng, ctr);
  }

  __device__ void init(const Params& Pr, const double* smem, int wl, int lane_, int replica_index) {
    P = &Pr; sm_means = smem; lane = lane_; d = Pr.d;
    beta = Pr.beta[Pr.first_chain + wl - 1];
    var_ref = VAR && Pr.var_tab != nullptr && Pr.first_chain + wl <= Pr.n_var;
#pragma unroll
    for (int k = 0; k < CPL; ++k) {
      x[k] = valid(k) ? Pr.x[(size_t)wl * Pr.d_pad + k * 32 + lane] : 0.0;
      on_mu[k] = 0.0; on_s2[k] = 0.0;
    }
    on_n = 0;
    rng.key0 = Pr.seed_lo; rng.key1 = (unsigned int)replica_index; rng.c2 = Pr.seed_hi; rng.c3 = 0u;
    rng.ctr = Pr.rng_ctr[wl];
    expl_acc = MeanAcc{0, 0.0}; am = MeanAcc{0, 0.0}; rev = MeanAcc{0, 0.0};
    n_steps = n_points = n_ref = 0;
    err = 0; e0 = e1 = 0.0;
    pool = 0.0; pool_base = 0; pool_valid = false;
    tw = 0; W = 1; gen = 0; slots = nullptr; e_prev = 0; rng_pool = nullptr;
    own1 = own2 = own3 = own4 = 0;
    t_trial = t_barrier = t_decide = 0;
  }
  __device__ void store(int wl) {
#pragma unroll
    for (int k = 0; k < CPL; ++k)
      if (valid(k)) P->x[(size_t)wl * P->d_pad + k * 32 + lane] = x[k];
  }
  __device__ void post(unsigned long long* pay, unsigned int tag) const {
#pragma unroll
    for (int k = 0; k < CPL; ++k) {
      const unsigned long long b = double_to_bits(x[k]);
      ll_store(pay + 2 * (k * 32 + lane), (unsigned int)b, tag);
      ll_store(pay + 2 * (k * 32 + lane) + 1, (unsigned int)(b >> 32), tag);
    }
  }
  __device__ bool adopt(const unsigned long long* pay, unsigned int tag) {
    bool ok = true;
#pragma unroll 1
    for (int j = 0; j < 2 * CPL; ++j) {   // wait until every word of the post has landed ...
      unsigned int dummy;
      ok = ll_load(pay + 2 * ((j >> 1) * 32 + lane) + (j & 1), tag, dummy) && ok;
    }
#pragma unroll
    for (int k = 0; k < CPL; ++k) {       // ... then take them
      const unsigned long long lo = ld_relaxed_sys(pay + 2 * (k * 32 + lane)), hi = ld_relaxed_sys(pay + 2 * (k * 32 + lane) + 1);
      x[k] = bits_to_double((hi << 32) | (lo & 0xffffffffull));
    }
    return __all_sync(PGN_FULL_MASK, ok);
  }
  // the team leader hands the adopted state to the other warps of the team through shared memory
  __device__ void share_state(double* sm) const {
#pragma unroll
    for (int k = 0; k < CPL; ++k) sm[k * 32 + lane] = x[k];
  }
  __device__ void load_state(const double* sm) {
#pragma unroll
    for (int k = 0; k < CPL; ++k) x[k] = sm[k * 32 + lane];
  }
  __device__ void write_trace(double* row) const {
#pragma unroll
    for (int k = 0; k < CPL; ++k)
      if (valid(k)) row[k * 32 + lane] = x[k];
  }
  __device__ void online_fit() {
    on_n += 1;
    const double g = div_<kCompact>(1.0, (double)on_n);
#pragma unroll
    for (int k = 0; k < CPL; ++k) {
      double mu_old = on_mu[k];
      double mu = mu_old + g * (x[k] - mu_old);
      on_s2[k] = on_s2[k] + g * ((x[k] - mu) * (x[k] - mu_old) - on_s2[k]);
      on_mu[k] = mu;
    }
  }
  __device__ void store_online() const {
#pragma unroll
    for (int k = 0; k < CPL; ++k)
      if (valid(k)) { P->online_mean[k * 32 + lane] = on_mu[k]; P->online_s2[k * 32 + lane] = on_s2[k]; }
    if (lane == 0) *P->online_n = on_n;
  }
  // per-replica recorders: the target-chain online statistics belong to the replica that produced them
  __device__ void flush_online(OnEntry* row) const {
#pragma unroll
    for (int k = 0; k < CPL; ++k)
      if (valid(k)) row[k * 32 + lane] = OnEntry{on_n, on_mu[k], on_s2[k]};
  }
  __device__ void load_online(const OnEntry* row) {
    // L2 loads: between the two target chains of a two-leg ladder the row was written by ANOTHER warp (scan_kernel)
    on_n = __ldcg(&row[0].n);
#pragma unroll
    for (int k = 0; k < CPL; ++k) {   // rows are d_pad wide and the padding stays zero: no test on the loads' results
      on_mu[k] = __ldcg(&row[k * 32 + lane].mu); on_s2[k] = __ldcg(&row[k * 32 + lane].s2);
    }
  }

  // ---- densities -----------------------------------------------------------
  __device__ __forceinline__ double toy_precision(double b) const { return (1.0 - b) * P->p[0] + b * P->p[1]; }

  // log_potential callable (InterpolatedLogPotential.jl:10-17 / ScaledPrecisionNormalPath.jl:19-20)
  __device__ __forceinline__ double lp_call(double b, double a0, double a1) const {
    if (TK == PGN_TARGET_TOY_MVN) return -0.5 * toy_precision(b) * a0;
    if (b == 0.0) return a0;
    if (b == 1.0) return a1;
    return (1.0 - b) * a0 + b * a1;
  }
  // LogDensityProblems.logdensity of the AD wrapper (BufferedAD.jl:89-94)
  __device__ __forceinline__ double lp_ad(double b, double a0, double a1) const {
    if (TK == PGN_TARGET_TOY_MVN) return -0.5 * toy_precision(b) * a0;
    return (1.0 - b) * a0 + b * a1;
  }

  // exp(a[m] - M) for the K <= 8 mixture components: lane m evaluates component m (one SIMT
  // pass instead of K redundant ones), then the values are broadcast — same bits as K calls.
  __device__ __forceinline__ void mode_exps(const double (&a)[KMAX_MODES], double M, int K, double (&w)[KMAX_MODES]) const {
    double mine = a[0];
#pragma unroll
    for (int m = 1; m < KMAX_MODES; ++m) mine = ((lane & 7) == m) ? a[m] : mine;
    const double e = exp_(mine - M);
#pragma unroll
    for (int m = 0; m < KMAX_MODES; ++m) w[m] = __shfl_sync(PGN_FULL_MASK, e, m);
    (void)K;
  }

  // ---- mixed Bool / Integer / Float product target (PGN_TARGET_MIXED; the state of test/test_slice_sampler.jl:56-75)
  __device__ __forceinline__ int coord_kind(int c) const {   // 0 Bool, 1 Integer, 2 Float
    if (TK != PGN_TARGET_MIXED) return 2;
    const int nb = (int)P->p[0], ni = (int)P->p[1];
    return c < nb ? 0 : (c < nb + ni ? 1 : 2);
  }
  // log density of coordinate c at value v under side 0 (reference) or 1 (target); Distributions.logpdf of a discrete
  // distribution at a non-integer or out-of-support point is -Inf
  __device__ __forceinline__ double mixed_term(int c, double v, int side) const {
    const double* t = P->means;
    const int n = (int)P->p[2];
    const int kind = coord_kind(c);
    if (kind == 0) return v == 1.0 ? t[2 * side] : (v == 0.0 ? t[2 * side + 1] : -PGN_INF);
    if (kind == 1) {
      if (!(v >= 0.0 && v <= (double)n) || v != floor(v)) return -PGN_INF;
      const int k = (int)v;
      return (t[10 + k] + (double)k * t[4 + 2 * side]) + (double)(n - k) * t[5 + 2 * side];
    }
    if (side == 1) return -(v * v + PGN_LOG2PI) * 0.5;
    return -(v * v * P->p[5] + PGN_LOG2PI) * 0.5 - P->p[4];
  }

  // reference term / gradient of coordinate slot k (valid) at xv: the fixed N(0, s^2 I) (DistributionLogPotential) or, on
  // the variational leg, gaussian_logdensity and its gradient (GaussianReference.jl:45-53, 72-80)
  __device__ __forceinline__ double var_t(int j, int k) const { return __ldg(P->var_tab + (size_t)j * P->d_pad + k * 32 + lane); }
  __device__ __forceinline__ double ref_term(int k, double xv, double ivr, double lsr) const {
    if constexpr (VAR) {
      if (var_ref) { const double dx = xv - var_t(0, k); return var_t(2, k) - var_t(3, k) * (dx * dx); }
    }
    return -(xv * xv * ivr + PGN_LOG2PI) * 0.5 - lsr;
  }
  __device__ __forceinline__ double ref_grad(int k, double xv, double ivr) const {
    if constexpr (VAR) {
      if (var_ref) return -(var_t(4, k) * (xv - var_t(0, k)));
    }
    return -xv * ivr;
  }

  // component densities at xx
  __device__ void eval(const double (&xx)[CPL], double& a0, double& a1) {
    n_points += 1;
    if (TK == PGN_TARGET_TOY_MVN) {
      double acc = 0.0;
#pragma unroll
      for (int k = 0; k < CPL; ++k) if (valid(k)) acc = acc + xx[k] * xx[k];
      a0 = warp_sum(acc); a1 = 0.0;
    } else if (TK == PGN_TARGET_FUNNEL) {
      const double y = __shfl_sync(PGN_FULL_MASK, xx[0], 0);
      const double e = exp_(-y);
      const double sy = P->p[0], lsy = P->p[1], ivr = P->p[5], lsr = P->p[4];
      double v[2] = {0.0, 0.0};
#pragma unroll
      for (int k = 0; k < CPL; ++k) if (valid(k)) {
        const double xv = xx[k];
        v[0] = v[0] + ref_term(k, xv, ivr, lsr);
        if (k == 0 && lane == 0) { double zy = y / sy; v[1] = v[1] + (-(zy * zy + PGN_LOG2PI) * 0.5 - lsy); }
        else { double t = xv * xv * e; v[1] = v[1] + (-(t + PGN_LOG2PI) * 0.5 - 0.5 * y); }
      }
      warp_sum_n<2>(v);
      a0 = v[0]; a1 = v[1];
    } else if (TK == PGN_TARGET_UNID) {   // unid_log_potential on the unit square, Uniform(0,1)^2 reference (warp-uniform)
      const double x0 = __shfl_sync(PGN_FULL_MASK, xx[0], 0), x1 = __shfl_sync(PGN_FULL_MASK, xx[0], 1);
      const bool in0 = x0 >= 0.0 && x0 <= 1.0, in1 = x1 >= 0.0 && x1 <= 1.0;
      a0 = (in0 ? 0.0 : -PGN_INF) + (in1 ? 0.0 : -PGN_INF);
      if constexpr (VAR) {   // variational leg: the Gaussian reference instead of the uniform prior
        if (var_ref) {
          double acc = 0.0;
#pragma unroll
          for (int k = 0; k < CPL; ++k) if (valid(k)) acc = acc + ref_term(k, xx[k], 0.0, 0.0);
          a0 = warp_sum(acc);
        }
      }
      if (in0 && in1) {
        const double pr = x0 * x1;
        a1 = P->p[1] * log_(pr) + (P->p[0] - P->p[1]) * log1p_<false>(-pr);
      } else a1 = -PGN_INF;
    } else if (TK == PGN_TARGET_MIXED) {
      double v[2] = {0.0, 0.0};
#pragma unroll
      for (int k = 0; k < CPL; ++k) if (valid(k)) {
        v[0] = v[0] + mixed_term(k * 32 + lane, xx[k], 0);
        v[1] = v[1] + mixed_term(k * 32 + lane, xx[k], 1);
      }
      warp_sum_n<2>(v);
      a0 = v[0]; a1 = v[1];
    } else {   // GMM
      const double ivr = P->p[5], lsr = P->p[4], ivm = P->p[2], cst = P->p[1];
      const int K = P->n_modes;
      double v[KMAX_MODES + 1];
#pragma unroll
      for (int m = 0; m <= KMAX_MODES; ++m) v[m] = 0.0;
#pragma unroll
      for (int k = 0; k < CPL; ++k) if (valid(k)) {
        const double xv = xx[k];
        v[KMAX_MODES] = v[KMAX_MODES] + ref_term(k, xv, ivr, lsr);
#pragma unroll
        for (int m = 0; m < KMAX_MODES; ++m) {
          double t = xv - sm_means[m * P->d_pad + k * 32 + lane];
          v[m] = v[m] + t * t;
        }
      }
      warp_sum_n<KMAX_MODES + 1>(v);
      a0 = v[KMAX_MODES];
      const double* lw = sm_means + (size_t)KMAX_MODES * P->d_pad;
      double a[KMAX_MODES];
      double M = -PGN_INF;
#pragma unroll
      for (int m = 0; m < KMAX_MODES; ++m) {
        a[m] = lw[m] - 0.5 * v[m] * ivm - cst;
        if (a[m] > M) M = a[m];
      }
      double wexp[KMAX_MODES];
      mode_exps(a, M, K, wexp);
      double s = 0.0;
#pragma unroll
      for (int m = 0; m < KMAX_MODES; ++m) s = s + wexp[m];
      a1 = M + log_(s);
    }
  }

  // densities + beta-combined raw gradient (BufferedAD.jl:98-111); `extra` is a
  // lane partial that rides along in the same butterfly (in: partial, out: sum)
  __device__ void eval_grad(const double (&xx)[CPL], double b, double& a0, double& a1, double (&g)[CPL], double& extra) {
    n_points += 1;
    if (TK == PGN_TARGET_MIXED || TK == PGN_TARGET_UNID) {   // no gradient-based explorer is ever selected for these targets
      a0 = a1 = 0.0;
#pragma unroll
      for (int k = 0; k < CPL; ++k) g[k] = 0.0;
      extra = warp_sum(extra);
    } else if (TK == PGN_TARGET_TOY_MVN) {
      double v[2] = {0.0, extra};
#pragma unroll
      for (int k = 0; k < CPL; ++k) if (valid(k)) v[0] = v[0] + xx[k] * xx[k];
      warp_sum_n<2>(v);
      a0 = v[0]; a1 = 0.0; extra = v[1];
      const double prec = toy_precision(b);
#pragma unroll
      for (int k = 0; k < CPL; ++k) g[k] = valid(k) ? -prec * xx[k] : 0.0;
    } else if (TK == PGN_TARGET_FUNNEL) {
      const double y = __shfl_sync(PGN_FULL_MASK, xx[0], 0);
      const double e = exp_(-y);
      const double sy = P->p[0], lsy = P->p[1], ivy = P->p[2], ivr = P->p[5], lsr = P->p[4];
      double v[4] = {0.0, 0.0, 0.0, extra};
#pragma unroll
      for (int k = 0; k < CPL; ++k) if (valid(k)) {
        const double xv = xx[k];
        v[0] = v[0] + ref_term(k, xv, ivr, lsr);
        if (k == 0 && lane == 0) {
          double zy = y / sy;
          v[1] = v[1] + (-(zy * zy + PGN_LOG2PI) * 0.5 - lsy);
          v[2] = v[2] + 0.0;
        } else {
          double t = xv * xv * e;
          v[1] = v[1] + (-(t + PGN_LOG2PI) * 0.5 - 0.5 * y);
          v[2] = v[2] + (0.5 * (xv * xv * e) - 0.5);
        }
      }
      warp_sum_n<4>(v);
      a0 = v[0]; a1 = v[1]; extra = v[3];
      const double T = v[2];
#pragma unroll
      for (int k = 0; k < CPL; ++k) {
        if (!valid(k)) { g[k] = 0.0; continue; }
        const double xv = xx[k];
        double gr = ref_grad(k, xv, ivr);
        double gt = (k == 0 && lane == 0) ? (-y * ivy + T) : (-xv * e);
        double acc = gr * (1.0 - b);
        g[k] = acc + gt * b;
      }
    } else {   // GMM
      const double ivr = P->p[5], lsr = P->p[4], ivm = P->p[2], cst = P->p[1];
      const int K = P->n_modes;
      double v[KMAX_MODES + 2];
#pragma unroll
      for (int m = 0; m < KMAX_MODES + 2; ++m) v[m] = 0.0;
      v[KMAX_MODES + 1] = extra;
#pragma unroll
      for (int k = 0; k < CPL; ++k) if (valid(k)) {
        const double xv = xx[k];
        v[KMAX_MODES] = v[KMAX_MODES] + ref_term(k, xv, ivr, lsr);
#pragma unroll
        for (int m = 0; m < KMAX_MODES; ++m) {
          double t = xv - sm_means[m * P->d_pad + k * 32 + lane];
          v[m] = v[m] + t * t;
        }
      }
      warp_sum_n<KMAX_MODES + 2>(v);
      a0 = v[KMAX_MODES]; extra = v[KMAX_MODES + 1];
      const double* lw = sm_means + (size_t)KMAX_MODES * P->d_pad;
      double w[KMAX_MODES];
      double M = -PGN_INF;
#pragma unroll
      for (int m = 0; m < KMAX_MODES; ++m) {
        w[m] = lw[m] - 0.5 * v[m] * ivm - cst;
        if (w[m] > M) M = w[m];
      }
      {
        double wexp[KMAX_MODES];
        mode_exps(w, M, K, wexp);
#pragma unroll
        for (int m = 0; m < KMAX_MODES; ++m) w[m] = wexp[m];
      }
      double s = 0.0;
#pragma unroll
      for (int m = 0; m < KMAX_MODES; ++m) s = s + w[m];
      a1 = M + log_(s);
#pragma unroll
      for (int k = 0; k < CPL; ++k) {
        if (!valid(k)) { g[k] = 0.0; continue; }
        const double xv = xx[k];
        double acc = 0.0;
#pragma unroll
        for (int m = 0; m < KMAX_MODES; ++m) acc = acc + w[m] * (sm_means[m * P->d_pad + k * 32 + lane] - xv);
        double gt = (acc / s) * ivm;
        double gr = ref_grad(k, xv, ivr);
        double t = gr * (1.0 - b);
        g[k] = t + gt * b;
      }
    }
  }

  // ---- sample_iid! / ToyExplorer ---------------------------------------------
  __device__ void sample_iid(double b) {
    if (TK == PGN_TARGET_MIXED) {   // rand! of the product reference: one tick per Bool / Float coordinate, n per Binomial
      const int nb = (int)P->p[0], ni = (int)P->p[1], n = (int)P->p[2];
      const double p0 = P->means[8], q0 = P->means[9];
#pragma unroll
      for (int k = 0; k < CPL; ++k) {
        if (!valid(k)) continue;
        const int c = k * 32 + lane;
        const int kind = coord_kind(c);
        const unsigned long long t = rng.ctr + (unsigned long long)(kind == 0 ? c : (kind == 1 ? nb + (c - nb) * n : nb + ni * n + (c - nb - ni)));
        if (kind == 0) x[k] = uniform_at(rng, t) < p0 ? 1.0 : 0.0;
        else if (kind == 1) {
          int cnt = 0;
          for (int j = 0; j < n; ++j) cnt += uniform_at(rng, t + (unsigned long long)j) < q0 ? 1 : 0;
          x[k] = (double)cnt;
        } else x[k] = P->p[3] * normal_at(rng, t);
      }
      rng.ctr += (unsigned long long)(nb + ni * n + (d - nb - ni));
      return;
    }
    if (TK == PGN_TARGET_UNID && !(VAR && var_ref)) {   // rand!(rng, product_distribution([Uniform(), Uniform()]), x)
#pragma unroll
      for (int k = 0; k < CPL; ++k)
        if (valid(k)) x[k] = uniform_at(rng, rng.ctr + (unsigned long long)(k * 32 + lane));
      rng.ctr += (unsigned long long)d;
      return;
    }
#pragma unroll
    for (int k = 0; k < CPL; ++k) {
      if (!valid(k)) continue;
      double z = normal_tick(rng.ctr + (unsigned long long)(k * 32 + lane));
      if (VAR && var_ref) { x[k] = z * var_t(1, k) + var_t(0, k); continue; }   // GaussianReference.jl:30-37
      if (TK == PGN_TARGET_TOY_MVN) x[k] = z / sqrt(toy_precision(b));   // toy_mvn_target.jl:18-21
      else x[k] = P->p[3] * z;                                           // rand!(rng, MvNormal(0, s^2 I), x)
    }
    rng.ctr += (unsigned long long)d;
  }

  // ---- SliceSampler (src/explorers/SliceSampler.jl:24-237) --------------------
  template <int KO>
  __device__ __forceinline__ double lp_at(int lo, double v) {
    double xx[CPL];
#pragma unroll
    for (int k = 0; k < CPL; ++k) xx[k] = (k == KO && lane == lo) ? v : x[k];
    double a0, a1;
    eval(xx, a0, a1);
    n_ref += 1;
    return lp_call(beta, a0, a1);
  }
  static __device__ __forceinline__ bool isapprox(double a, double b) {
    const double rtol = bits_to_double(0x3e50000000000000ULL);   // sqrt(eps(Float64)) = 2^-26
    if (a == b) return true;
    if (!(is_finite(a) && is_finite(b))) return false;
    double aa = fabs(a), ab = fabs(b);
    return fabs(a - b) <= rtol * (aa > ab ? aa : ab);
  }
  template <int KO>
  __device__ bool slice_accept(int lo, double old_position, double new_position, double z, double L, double R,
                               double lp_L, double lp_R) {   // :192-237
    const double w = P->slice_w;
    double Lhat = L, Rhat = R;
    bool Rstale = false, Lstale = false, D = false;
    while (Rhat - Lhat > 1.1 * w) {
      double M = (Lhat + Rhat) / 2.0;
      if (((old_position < M) && (new_position >= M)) || ((old_position >= M) && (new_position < M))) D = true;
      if (new_position < M) { Rhat = M; Rstale = true; } else { Lhat = M; Lstale = true; }
      if (D) {
        if (Lstale) { lp_L = lp_at<KO>(lo, Lhat); Lstale = false; }
        if (Rstale) { lp_R = lp_at<KO>(lo, Rhat); Rstale = false; }
        if ((z >= lp_L) && (z >= lp_R)) { expl_acc.fit(0.0); return false; }
      }
    }
    expl_acc.fit(1.0);
    return true;
  }
  // rand(rng, a:b) for integers a <= b held in doubles: a + floor(u (b - a + 1)), u the replica's next uniform
  __device__ __forceinline__ double rand_int_range(double a, double b) {
    const double span = (b - a) + 1.0;
    double k = floor(draw_uniform() * span);
    if (k > b - a) k = b - a;
    return a + k;
  }
  template <int KO>
  __device__ double slice_coord(int lo, double cached_lp) {   // :89-186
    const double w = P->slice_w;
    const double cur = __shfl_sync(PGN_FULL_MASK, x[KO], lo);
    const int kind = coord_kind(KO * 32 + lo);
    if (TK == PGN_TARGET_MIXED && kind == 0) {   // Bool: sample from the full conditional, one density evaluation (:65-86)
      double lp0, lp1;
      if (cur != 0.0) { lp1 = cached_lp; lp0 = lp_at<KO>(lo, 0.0); }
      else { lp0 = cached_lp; lp1 = lp_at<KO>(lo, 1.0); }
      const double prob_ratio = exp_(lp1 - lp0);
      const double prob_zero = 1.0 / (1.0 + prob_ratio);
      const bool zero = draw_uniform() < prob_zero;
      if (lane == lo) x[KO] = zero ? 0.0 : 1.0;
      return zero ? lp0 : lp1;
    }
    const bool integer = TK == PGN_TARGET_MIXED && kind == 1;
    const double z = cached_lp - draw_exponential();
    // slice_double :97-126
    double L, R;
    if (integer) {   // initialize_slice_endpoints for integers :136-142
      if (w != floor(w)) { err = PGN_ERR_INVALID; return 0.0; }
      L = cur - rand_int_range(0.0, ceil(w));
      R = L + ceil(w);
    } else {
      L = cur - w * draw_uniform();
      R = L + w;
    }
    int K = P->slice_p;
    double lp_L = lp_at<KO>(lo, L);
    double lp_R = lp_at<KO>(lo, R);
    while (K > 0 && ((z < lp_L) || (z < lp_R))) {
      double V = draw_uniform();
      if (V <= 0.5) { L = L - (R - L); lp_L = lp_at<KO>(lo, L); }
      else { R = R + (R - L); lp_R = lp_at<KO>(lo, R); }
      K -= 1;
    }
    n_steps += P->slice_p - K;
    // slice_shrink! :144-186
    double Lbar = L, Rbar = R;
    int n = 1;
    while (n <= P->slice_max_iter) {
      double new_position = integer ? rand_int_range(Lbar, Rbar) : Lbar + draw_uniform() * (Rbar - Lbar);   // :188-189
      double new_lp = lp_at<KO>(lo, new_position);
      bool consider = z < new_lp;
      if (consider && slice_accept<KO>(lo, cur, new_position, z, L, R, lp_L, lp_R)) {
        if (lane == lo) x[KO] = new_position;
        n_steps += n;
        return new_lp;
      }
      if (new_position < cur) Lbar = new_position; else Rbar = new_position;
      if (integer ? (Lbar == Rbar) : isapprox(Lbar, Rbar)) {   // isapprox of two Integers is ==
        n_steps += n;
        return lp_at<KO>(lo, cur);
      }
      n += 1;
    }
    err = PGN_ERR_SLICE_MAX_ITER;
    return 0.0;
  }
  template <int KO>
  __device__ double slice_block(double cached_lp) {
    if (KO * 32 >= d) return cached_lp;
    const int hi = (d - KO * 32) < 32 ? (d - KO * 32) : 32;
    for (int lo = 0; lo < hi; ++lo) {
      cached_lp = slice_coord<KO>(lo, cached_lp);
      if (err) return cached_lp;
      if (!is_finite(cached_lp)) { err = PGN_ERR_BAD_DENSITY; return cached_lp; }   // :52-59
    }
    return cached_lp;
  }
  __device__ void slice_step() {   // :24-30
    double cached_lp = -PGN_INF;
    for (int pass = 0; pass < P->slice_n_passes; ++pass) {
      if (cached_lp == -PGN_INF) {   // cached_log_potential :32-41
        double a0, a1;
        eval(x, a0, a1);
        n_ref += 1;
        double result = lp_call(beta, a0, a1);
        if (result == -PGN_INF) { err = PGN_ERR_BAD_DENSITY; return; }
        cached_lp = result;
      }
      cached_lp = slice_block<0>(cached_lp); if (err) return;
      if (CPL > 1) { cached_lp = slice_block<(CPL > 1 ? 1 : 0)>(cached_lp); if (err) return; }
      if (CPL > 2) { cached_lp = slice_block<(CPL > 2 ? 2 : 0)>(cached_lp); if (err) return; }
      if (CPL > 3) { cached_lp = slice_block<(CPL > 3 ? 3 : 0)>(cached_lp); if (err) return; }
    }
    eval(x, e0, e1);   // densities at the final state, consumed by the swap
  }

  // ---- autoMALA (src/explorers/AutoMALA.jl:84-275, hamiltonian_dynamics.jl:28-84) ----
  struct Trial {
    double x1[CPL], p1[CPL], g1c[CPL];
    double a0, a1, lp1, h_after, eps;
  };
  // one leap_frog! + log_joint from (xs, ps) with conditioned gradient gs at xs;
  // returns h_after - h_before (log_joint_difference_function :250-275).
  // pre_one: the preconditioner is exactly the identity, x / 1.0 == x bit for bit.
  __device__ __forceinline__ double run_trial(const double (&xs)[CPL], const double (&ps)[CPL],
                                              const double (&gs)[CPL], const double (&pre)[CPL], bool pre_one,
                                              double eps, double h_before, Trial& T) {
    double ph[CPL];
    double pp = 0.0;
    const double half_eps = eps * 0.5;   // == eps / 2 for every double
#pragma unroll
    for (int k = 0; k < CPL; ++k) {
      ph[k] = ps[k] + half_eps * gs[k];
      T.x1[k] = xs[k] + eps * (pre_one ? ph[k] : ph[k] / pre[k]);
      pp = valid(k) ? pp + ph[k] * ph[k] : pp;
    }
    double graw[CPL];
    eval_grad(T.x1, beta, T.a0, T.a1, graw, pp);
    T.lp1 = lp_ad(beta, T.a0, T.a1);
#pragma unroll
    for (int k = 0; k < CPL; ++k) T.g1c[k] = pre_one ? graw[k] : graw[k] / pre[k];
    const double cur = T.lp1 - 0.5 * pp;
    double s2;
    if (!is_finite(cur)) {   // hamiltonian_dynamics.jl:56-59: early return, no last half-step
#pragma unroll
      for (int k = 0; k < CPL; ++k) T.p1[k] = ph[k];
      s2 = pp;
    } else {
      double q = 0.0;
#pragma unroll
      for (int k = 0; k < CPL; ++k) {
        T.p1[k] = ph[k] + half_eps * T.g1c[k];
        q = valid(k) ? q + T.p1[k] * T.p1[k] : q;
      }
      s2 = warp_sum(q);
    }
    T.h_after = T.lp1 - 0.5 * s2;
    T.eps = eps;
    return T.h_after - h_before;
  }
  __device__ bool build_preconditioner(double (&pre)[CPL], int precond_kind, double mix_p0, double mix_p01) {   // Preconditioner.jl:57-77; returns "is identity"
    const bool have = P->std_devs != nullptr;
    if (!have || precond_kind == PGN_PRECOND_IDENTITY) {
#pragma unroll
      for (int k = 0; k < CPL; ++k) pre[k] = 1.0;
      return true;
    }
    double sd[CPL];
#pragma unroll
    for (int k = 0; k < CPL; ++k) sd[k] = valid(k) ? P->std_devs[k * 32 + lane] : 0.0;
    if (precond_kind == PGN_PRECOND_DIAGONAL) {
#pragma unroll
      for (int k = 0; k < CPL; ++k) pre[k] = sd[k] == 0.0 ? 1.0 : div_<kCompact>(1.0, sd[k]);
      return false;
    }
    const double u = draw_uniform();
    if (u <= mix_p0) {
#pragma unroll
      for (int k = 0; k < CPL; ++k) pre[k] = sd[k] == 0.0 ? 1.0 : div_<kCompact>(1.0, sd[k]);
      return false;
    } else if (u <= mix_p01) {
#pragma unroll
      for (int k = 0; k < CPL; ++k) pre[k] = 1.0;
      return true;
    } else {
      const double mix = draw_uniform();
      const double rmix = 1.0 - mix;
#pragma unroll
      for (int k = 0; k < CPL; ++k) pre[k] = sd[k] == 0.0 ? 1.0 : mix + div_<kCompact>(rmix, sd[k]);
      return false;
    }
  }
  // auto_mala! :106-182 with the step-size searches (auto_step_size :184-248) evaluated by the
  // chain's TEAM of W warps.  The reference walks the candidate steps eps0*2^k one at a time
  // (k = 0, then -1, -2, ... or +1, +2, ...); every candidate is an independent trial from the same
  // start point, so a team round evaluates W of them at once — warp w takes offset r1_k(w) in the
  // first round and the next W offsets of the chosen direction afterwards — publishes the results
  // in shared memory, and then every warp replays the reference's sequential decisions on the
  // published log-joint differences.  Results (exponent, chosen trial, every statistic) are the
  // reference's; only the wall-clock order of the evaluations changes.  The trial at the chosen
  // step is taken from the published results (the reference re-runs leap_frog! there, :144-151),
  // and the reversed search stops at its exponent (:160-163 use nothing else).
  // The forward and the reversed searches share ONE inlined copy of run_trial; the density and
  // gradient at the scan's starting state go through the same copy (pass i = -1: a "trial" with
  // zero momentum and eps = 0 lands exactly on x), so the kernel contains a single instance of the
  // target's density code.  n_refresh_eff = 0 gives the densities only (reference chain after
  // sample_iid!).
  // Buffers: round r writes buffer r mod 3 and reads r mod 3 and (r-1) mod 3 after its barrier, so a
  // warp that has run ahead into round r+1 never touches what a slower team-mate still reads.
  // First-round candidates: offset 0, n_neg offsets below it and W - 1 - n_neg above it.  Which side gets
  // the speculative slots is a guess from the exponent this chain's previous forward search ended at
  // (neighbouring refreshments of a chain want similar steps); the guess changes which trials are
  // evaluated early, never which one is chosen.
  __device__ __forceinline__ int r1_n_neg() const {
    const int S = W - 1;
    // a search that ends at exponent e < 0 needs the offsets -1..e, one that ends at e >= 0 needs +1..e+1
    int neg = e_prev < 0 ? 1 - e_prev : 1, pos = e_prev >= 0 ? e_prev + 2 : 1;   // one spare on the predicted side, one slot on the other
    if (e_prev < 0) { neg = neg < S ? neg : S; pos = pos < S - neg ? pos : S - neg; }
    else { pos = pos < S ? pos : S; neg = neg < S - pos ? neg : S - pos; }
    const int left = S - neg - pos;
    return neg + (e_prev < 0 ? (left + 1) / 2 : left / 2);
  }
  __device__ __forceinline__ int r1_k(int w, int n_neg) const { return w == 0 ? 0 : (w <= n_neg ? -w : w - n_neg); }
  __device__ __forceinline__ int r1_warp(int k, int n_neg) const { return k < 0 ? -k : n_neg + k; }
  __device__ __forceinline__ void publish(int buf, const Trial& T, double diff) {
    double* s = slot(buf, tw);
#pragma unroll
    for (int k = 0; k < CPL; ++k) {
      s[k * 32 + lane] = T.x1[k];
      s[(CPL + k) * 32 + lane] = T.p1[k];
      s[(2 * CPL + k) * 32 + lane] = T.g1c[k];
    }
    if (lane == 0) {
      double* h = s + 3 * CPL * 32;
      h[0] = T.a0; h[1] = T.a1; h[2] = T.lp1; h[3] = T.h_after; h[4] = diff; h[5] = T.eps;
    }
  }
  __device__ __forceinline__ void fetch(int buf, int w, Trial& T) const {
    const double* s = slot(buf, w);
#pragma unroll
    for (int k = 0; k < CPL; ++k) {
      T.x1[k] = s[k * 32 + lane];
      T.p1[k] = s[(CPL + k) * 32 + lane];
      T.g1c[k] = s[(2 * CPL + k) * 32 + lane];
    }
    const double* h = s + 3 * CPL * 32;
    T.a0 = h[0]; T.a1 = h[1]; T.lp1 = h[2]; T.h_after = h[3]; T.eps = h[5];
  }
  // step s of a Compose / Mix program (s >= 0), or the explorer's own parameters (s < 0)
  __device__ void automala(bool use_mh, int n_refresh_eff, int s = -1) {
    double pre[CPL];
    bool pre_one = true;
    // Mix (src/explorers/Mix.jl:20-21): one tick of the replica's stream picks the autoMALA variant of this step
    double step0 = P->step_size, v_p0 = P->mix_p0, v_p01 = P->mix_p01;
    int v_precond = P->precond_kind;
    if (s >= 0) {
      step0 = P->mix_step_size[s]; v_precond = P->mix_precond_kind[s]; v_p0 = P->mix_variant_p0[s]; v_p01 = P->mix_variant_p01[s];
    } else if (n_refresh_eff > 0 && P->n_mix > 1) {
      int v = (int)(draw_uniform() * (double)P->n_mix);
      v = v >= P->n_mix ? P->n_mix - 1 : v;
      n_refresh_eff = P->mix_n_refresh[v]; step0 = P->mix_step_size[v]; v_precond = P->mix_precond_kind[v];
      v_p0 = P->mix_variant_p0[v]; v_p01 = P->mix_variant_p01[v];
    }
    if (n_refresh_eff > 0) pre_one = build_preconditioner(pre, v_precond, v_p0, v_p01);
    else {
#pragma unroll
      for (int k = 0; k < CPL; ++k) pre[k] = 1.0;
    }
    double g0[CPL];
#pragma unroll
    for (int k = 0; k < CPL; ++k) g0[k] = 0.0;
    double lp0 = 0.0;
    if (!(step0 > 0)) { err = PGN_ERR_INVALID; return; }
    // A scan's draws are fixed ticks of the replica's stream once the preconditioner has taken its own:
    // refreshment i uses ticks [c0 + i*stride, +d) for the momentum and the next 2 (3 with MH) for a, b, (u).
    // The team draws them all now, each warp a share of the SIMT passes, instead of every warp drawing
    // every one of them again at its refreshment.
    const bool pooled = rng_pool != nullptr && W > 1 && n_refresh_eff > 0 && n_refresh_eff <= P->pool_refresh;
    if (pooled) {
      const unsigned long long c0 = rng.ctr, stride = (unsigned long long)d + (use_mh ? 3ull : 2ull);
      for (int ir = tw; ir < n_refresh_eff; ir += W) {   // team warp w draws for refreshments w, w + W, ...
        double* rp = rng_pool + (size_t)ir * POOL_DOUBLES;
        const unsigned long long cr = c0 + (unsigned long long)ir * stride;
        double pp = 0.0;
#pragma unroll
        for (int k = 0; k < CPL; ++k) {
          const double z = valid(k) ? normal_tick(cr + (unsigned long long)(k * 32 + lane)) : 0.0;
          rp[k * 32 + lane] = z;
          pp = valid(k) ? pp + z * z : pp;
        }
        const double uu = uniform_at(rng, cr + (unsigned long long)(d + (lane < 3 ? lane : 0)));
        const double lu = log_(uu);
        const double spp = warp_sum(pp);                 // the momentum's squared norm, same tree as at the refreshment
        if (lane < 3) { rp[CPL * 32 + lane] = uu; rp[CPL * 32 + 4 + lane] = lu; }
        if (lane == 3) rp[CPL * 32 + 3] = spp;
      }
      team_barrier();
    }
    Trial T;
#pragma unroll 1
    for (int i = -1; i < n_refresh_eff; ++i) {
      double sx[CPL], sp[CPL], sg[CPL];     // start point of the current search
      double fx[CPL], fg[CPL];              // forward proposal (kept across the reversed search)
      double f_a0 = 0.0, f_a1 = 0.0, f_lp = 0.0, h_rev = 0.0;
      double lower = 0.0, upper = 0.0, u_mh = 0.0, init_joint = 0.0;
      int expo_fwd = 0, expo_rev = 0;
      if (i >= 0) {
        double p[CPL];
        double pp = 0.0;
        double a, b, la, lb, spp = 0.0;
        if (pooled) {   // this refreshment's draws were made at the start of the scan (same ticks, same values)
          const double* rp = rng_pool + (size_t)i * POOL_DOUBLES;
#pragma unroll
          for (int k = 0; k < CPL; ++k) p[k] = rp[k * 32 + lane];
          a = rp[CPL * 32]; b = rp[CPL * 32 + 1]; u_mh = rp[CPL * 32 + 2]; spp = rp[CPL * 32 + 3];
          la = rp[CPL * 32 + 4]; lb = rp[CPL * 32 + 5];
          rng.ctr += (unsigned long long)d + (use_mh ? 3ull : 2ull);
        } else {
#pragma unroll
        for (int k = 0; k < CPL; ++k) {
          p[k] = valid(k) ? normal_tick(rng.ctr + (unsigned long long)(k * 32 + lane)) : 0.0;   // randn!(rng, momentum)
          pp = valid(k) ? pp + p[k] * p[k] : pp;
        }
        rng.ctr += (unsigned long long)d;
        // a, b (:132-133) and the MH uniform (:173) are the next three ticks of this replica's
        // stream; lanes 0..2 draw them (and the logs of a, b) in one SIMT pass
        double mine = uniform_at(rng, rng.ctr + (unsigned long long)(lane < 3 ? lane : 0));
        double lmine = log_(mine);
        rng.ctr += use_mh ? 3ull : 2ull;
        a = __shfl_sync(PGN_FULL_MASK, mine, 0); b = __shfl_sync(PGN_FULL_MASK, mine, 1);
        la = __shfl_sync(PGN_FULL_MASK, lmine, 0); lb = __shfl_sync(PGN_FULL_MASK, lmine, 1);
        u_mh = __shfl_sync(PGN_FULL_MASK, mine, 2);
        }
        lower = a < b ? la : lb;     // log(min(a, b))
        upper = a < b ? lb : la;     // log(max(a, b))
        init_joint = lp0 - 0.5 * (pooled ? spp : warp_sum(pp));
        if (!is_finite(init_joint)) { err = PGN_ERR_NOT_POSITIVE; return; }
        if (!(lower < upper)) { err = PGN_ERR_INVALID; return; }
#pragma unroll
        for (int k = 0; k < CPL; ++k) { sx[k] = x[k]; sp[k] = p[k]; sg[k] = g0[k]; fx[k] = 0.0; fg[k] = 0.0; }
      } else {
#pragma unroll
        for (int k = 0; k < CPL; ++k) { sx[k] = x[k]; sp[k] = 0.0; sg[k] = 0.0; fx[k] = 0.0; fg[k] = 0.0; }
      }
      double h_before = init_joint;
      const int n_dir = (i >= 0 && use_mh) ? 2 : 1;
#pragma unroll 1
      for (int dir = 0; dir < n_dir; ++dir) {
        // phase 0: first team round; 1: later rounds in direction sgn; 2: one trial evaluated by every
        // warp for itself (densities at x, or the chosen step when it is not among the candidates)
        int phase = i >= 0 ? 0 : 2;
        int sgn = 0, m = 0, exponent = 0, nst = 0;
        int prev_buf = 0, prev_w = 0;        // slot of the candidate examined last
        const int n_neg = r1_n_neg();
        double eps_m = step0;         // step of the farthest candidate examined in direction sgn
        double eps = 0.0;
        if (i >= 0) {
          const int k0 = r1_k(tw, n_neg);
          // the reference halves / doubles the step one factor at a time (:216-248); while every intermediate is
          // a normal number that is the exact product with 2^k0, otherwise walk the same way it does
          eps = step0 * pow2i(k0);
          const unsigned int ebits = (unsigned int)(double_to_bits(eps) >> 52) & 0x7ffu;
          if (ebits == 0u || ebits == 0x7ffu) {
            eps = step0;
            for (int j = 0; j < (k0 < 0 ? -k0 : k0); ++j) eps = eps * (k0 < 0 ? 0.5 : 2.0);   // x * 0.5 == x / 2.0 for every double
          }
        }
#pragma unroll 1
        while (true) {
          const long long tc0 = clock64();
          const double diff = run_trial(sx, sp, sg, pre, pre_one, eps, h_before, T);
          const long long tc1 = clock64();
          t_trial += tc1 - tc0;
          if (phase == 2) break;
          const int buf = gen;
          gen = gen == 2 ? 0 : gen + 1;
          publish(buf, T, diff);
          team_barrier();
          const long long tc2 = clock64();
          t_barrier += tc2 - tc1;
          // Replay of the reference's sequential walk over the published candidates: lane l holds candidate l of
          // this round, a ballot finds the first one that stops the walk.
          bool decided = false;
          int win_buf = buf, win_w = 0;
          const double* hh = slot(buf, lane < W ? lane : 0) + 3 * CPL * 32;
          const double dn = hh[4], en = hh[5];
          const unsigned int team_mask = (1u << W) - 1u;
          int first_w, last_w;                 // candidates examined this round, in walk order
          if (phase == 0) {
            const double d0 = __shfl_sync(PGN_FULL_MASK, dn, 0);
            if (!is_finite(d0) || d0 < lower) sgn = -1;            // auto_step_size :203-209
            else if (d0 > upper) sgn = 1;
            else decided = true;
            prev_buf = buf; prev_w = 0;
            first_w = sgn < 0 ? 1 : n_neg + 1;                      // offsets -1..-n_neg sit in warps 1..n_neg, +1.. after them
            last_w = decided ? 0 : (sgn < 0 ? n_neg : W - 1);
          } else {
            first_w = 0; last_w = W - 1;
          }
          if (!decided && last_w >= first_w) {
            // shrink_step_size :228-248 stops at eps == 0 (error) or diff > lower; grow_step_size :216-226 at a
            // non-finite diff or diff < upper
            const bool stop = sgn < 0 ? (en == 0.0 || dn > lower) : (!is_finite(dn) || dn < upper);
            const unsigned int range = (team_mask >> first_w << first_w) & (0xffffffffu >> (31 - last_w));
            const unsigned int hit = __ballot_sync(PGN_FULL_MASK, stop) & range;
            if (hit != 0u) {
              const int w = __ffs((int)hit) - 1;
              const int n = m + (w - first_w) + 1;
              if (sgn < 0) {
                if (__shfl_sync(PGN_FULL_MASK, en, w) == 0.0) { err = PGN_ERR_STEP_UNDERFLOW; return; }
                nst = n; exponent = -n; win_buf = buf; win_w = w;
              } else {
                nst = n; exponent = n - 1;
                if (w == first_w) { win_buf = prev_buf; win_w = prev_w; } else { win_buf = buf; win_w = w - 1; }
              }
              decided = true;
            } else {
              m += last_w - first_w + 1;
              prev_buf = buf; prev_w = last_w;
              eps_m = __shfl_sync(PGN_FULL_MASK, en, last_w);
            }
          }
          t_decide += clock64() - tc2;
          if (decided) {
            if (dir == 1) break;                                   // reversed search: only the exponent is used (:160-163)
            const double eps_final = step0 * pow2(exponent);   // leap_frog! at the chosen step :144-151
            if (slot(win_buf, win_w)[3 * CPL * 32 + 5] == eps_final) {
              if (!(win_buf == buf && win_w == tw)) fetch(win_buf, win_w, T);
              break;
            }
            phase = 2; eps = eps_final;
          } else {
            phase = 1;
            eps = eps_m;
            for (int j = 0; j <= tw; ++j) eps = eps * (sgn < 0 ? 0.5 : 2.0);
          }
        }
        if (i < 0) break;
        n_steps += 1 + nst;
        if (tw == own(1)) am.template fit<kCompact>(pow2(exponent));
        if (dir == 0) expo_fwd = exponent; else expo_rev = exponent;
        if (dir == 0) {
          e_prev = exponent;
          n_ref += 1 + 1 + 3 * (1 + nst) + 2;
          h_rev = T.h_after;      // log_joint at (x1, -p1): same partial sums as h_after
          h_before = h_rev;
          f_a0 = T.a0; f_a1 = T.a1; f_lp = T.lp1;
#pragma unroll
          for (int k = 0; k < CPL; ++k) {
            fx[k] = T.x1[k]; fg[k] = T.g1c[k];
            sx[k] = T.x1[k]; sp[k] = T.p1[k] * -1.0; sg[k] = T.g1c[k];   // momentum .*= -1.0 :155
          }
        } else {
          n_ref += 1 + 3 * (1 + nst);
        }
      }
      if (i < 0) {   // densities and conditioned gradient at the scan's starting state
        e0 = T.a0; e1 = T.a1; lp0 = T.lp1;
#pragma unroll
        for (int k = 0; k < CPL; ++k) g0[k] = T.g1c[k];
        continue;
      }
      bool accept = true;
      if (use_mh) {
        const bool passed = (expo_rev == expo_fwd);
        if (tw == own(2)) rev.template fit<kCompact>(passed ? 1.0 : 0.0);
        double prob = 0.0;
        if (passed) { double e = exp_(h_rev - init_joint); prob = 1.0 < e ? 1.0 : e; n_ref += 1; }
        if (tw == own(3)) expl_acc.template fit<kCompact>(prob);
        accept = u_mh < prob;
      }
      if (accept) {
#pragma unroll
        for (int k = 0; k < CPL; ++k) { x[k] = fx[k]; g0[k] = fg[k]; }
        e0 = f_a0; e1 = f_a1; lp0 = f_lp;
      }
    }
  }

  // mala! (src/explorers/MALA.jl:74-97): one leapfrog at the fixed step size + MH, n_refresh times
  __device__ void mala(int s = -1) {
    double pre[CPL];
    const int n_refresh = s >= 0 ? P->mix_n_refresh[s] : P->n_refresh;
    const double step_size = s >= 0 ? P->mix_step_size[s] : P->step_size;
    const bool pre_one = s >= 0 ? build_preconditioner(pre, P->mix_precond_kind[s], P->mix_variant_p0[s], P->mix_variant_p01[s])
                                : build_preconditioner(pre, P->precond_kind, P->mix_p0, P->mix_p01);
    double g0[CPL];
    {
      double dummy = 0.0;
      double graw[CPL];
      eval_grad(x, beta, e0, e1, graw, dummy);
#pragma unroll
      for (int k = 0; k < CPL; ++k) g0[k] = pre_one ? graw[k] : graw[k] / pre[k];
    }
    double lp0 = lp_ad(beta, e0, e1);
    Trial T;
    for (int i = 0; i < n_refresh; ++i) {
      double p[CPL];
      double pp = 0.0;
#pragma unroll
      for (int k = 0; k < CPL; ++k) {
        p[k] = valid(k) ? normal_tick(rng.ctr + (unsigned long long)(k * 32 + lane)) : 0.0;
        pp = valid(k) ? pp + p[k] * p[k] : pp;
      }
      rng.ctr += (unsigned long long)d;
      const double init_joint = lp0 - 0.5 * warp_sum(pp);
      if (!is_finite(init_joint)) { err = PGN_ERR_NOT_POSITIVE; return; }
      run_trial(x, p, g0, pre, pre_one, step_size, init_joint, T);
      const double e = exp_(T.h_after - init_joint);   // final_joint_log at (x1, -p1) has the same partial sums
      const double prob = 1.0 < e ? 1.0 : e;
      expl_acc.fit(prob);
      n_ref += 4;
      if (draw_uniform() < prob) {
#pragma unroll
        for (int k = 0; k < CPL; ++k) { x[k] = T.x1[k]; g0[k] = T.g1c[k]; }
        e0 = T.a0; e1 = T.a1; lp0 = T.lp1;
      }
      n_steps += 1;
    }
  }

  // ---- explore!(pt, replica, explorer) (src/pt/pigeons.jl:101-132) --------------
  __device__ void explore(long long scan, bool is_reference) {
    if (EX == PGN_EXPLORER_COMPOSE) {
      // Compose (Compose.jl:16-19): every explorer of the program in turn; Mix (Mix.jl:20-21): one of them, drawn with one
      // tick of the replica's stream.  Steps that are not autoMALA run on every warp of the team, on identical data.
      if (is_reference) { sample_iid(beta); automala(scan != 1, 0); return; }   // densities of the fresh state only
      int first = 0, last = P->n_steps;
      if (P->program_is_mix) {
        int v = (int)(draw_uniform() * (double)P->n_steps);
        v = v >= P->n_steps ? P->n_steps - 1 : v;
        first = v; last = v + 1;
      }
      for (int s = first; s < last; ++s) {
        const int kind = P->step_kind[s];
        if (kind == PGN_EXPLORER_TOY) { sample_iid(beta); eval(x, e0, e1); }
        else if (kind == PGN_EXPLORER_SLICE) slice_step();
        else if (kind == PGN_EXPLORER_MALA) mala(s);
        else automala(scan != 1, P->mix_n_refresh[s], s);   // AutoMALA.jl:87,102
        if (err) return;
      }
      return;
    }
    if (EX == PGN_EXPLORER_AUTOMALA) {
      if (is_reference) sample_iid(beta);
      automala(scan != 1, is_reference ? 0 : P->n_refresh);   // AutoMALA.jl:87,102
      return;
    }
    if (is_reference) { sample_iid(beta); eval(x, e0, e1); return; }
    if (EX == PGN_EXPLORER_TOY) { sample_iid(beta); eval(x, e0, e1); }
    else if (EX == PGN_EXPLORER_SLICE) slice_step();
    else mala();
  }
  // log_unnormalized_ratio (src/log_potentials/log_potentials.jl:43-51)
  __device__ double log_ratio(double beta_partner) const {
    return lp_call(beta_partner, e0, e1) - lp_call(beta, e0, e1);
  }
  // Pooled uniforms: the replica's draws are consecutive ticks of one Philox stream, so the next
  // 32 are produced in one SIMT pass (lane l: tick pool_base + l) and handed out by shuffle —
  // same stream, values and order as one Philox per draw, without the 10 rounds on the serial path.
  __device__ __forceinline__ double draw_uniform() {
    unsigned long long idx = rng.ctr - pool_base;
    if (!pool_valid || idx >= 32ull) {
      pool_base = rng.ctr;
      pool = uniform_at(rng, pool_base + (unsigned long long)lane);
      pool_valid = true;
      idx = 0;
    }
    rng.ctr += 1;
    return __shfl_sync(PGN_FULL_MASK, pool, (int)idx);
  }
  __device__ __forceinline__ double draw_exponential() { return -log_(1.0 - draw_uniform()); }
  __device__ __forceinline__ void on_replica_changed() { pool_valid = false; }
};

// ===========================================================================
// Ising chain (examples/ising.jl): lane i holds row i of the bit-packed lattice
// ===========================================================================
// TABLE = true: the Metropolis ratios of the round are tabulated in shared memory (16.7 KB per chain: at most ~1.9 K
// chains per GPU).  TABLE = false: the same expression is evaluated where it is needed (about twice the instructions per
// site), no shared memory and a 64-register budget, so that up to 32 warps = 32 chains share an SM (4.7 K chains per GPU:
// BASELINE config 4 on ONE GPU).
template <bool TABLE>
struct IsingChainT {
  static constexpr bool kTestSwapper = false;
  static constexpr bool kTeam = false;
  static constexpr bool kIsing = true;
  static constexpr bool kMixed = false;
  static constexpr int kMaxThreads = TABLE ? 256 : 64;
  static constexpr int kMinBlocksPerSM = TABLE ? 1 : 16;
  static constexpr bool kCompact = false;
  const Params* P;
  int lane, L;
  double beta;
  unsigned int row;
  int S;               // sum_pair_products (examples/ising.jl:19)
  Rng rng;
  MeanAcc expl_acc, am, rev;
  long long n_steps, n_points, n_ref;
  int err;
  // The replica's uniforms are consecutive ticks of one Philox stream, so the next 32 of them
  // can be generated in ONE SIMT pass (lane l computes tick pool_base + l) and handed out with
  // a shuffle: the 10-round Philox leaves the serial per-site dependency chain.  Same stream,
  // same values, same order as drawing them one by one.
  double pool;
  unsigned long long pool_base;
  bool pool_valid;

  // A chain keeps its beta for the whole round, and a single flip changes S by dS in {0, +-4, +-8}
  // with S = 2 L^2 - 4 m, so the Metropolis ratio exp(lp(S + dS) - lp(S)) takes (L^2 + 1) x 4 values
  // per round.  The two dS that LOWER lp (the only ones whose ratio can be < 1: lp is monotone in S and
  // exp_ of a non-negative argument is >= 1) are tabulated in shared memory when the round starts —
  // the very expression the reference evaluates per site (examples/ising.jl:107-109), evaluated once
  // per (S, dS) instead of once per site.
  const double* tbl;   // [(L^2 + 1)][2]: |dS| = 4, 8
  int sig;             // sign of the dS that lowers lp
  int S0;              // 2 L^2
  static constexpr int TBL_PAD = 10;   // rows on either side for the wrong guesses of the speculative sweep (5 sites x +-2)
  static __host__ __device__ int table_doubles(int L_) { return TABLE ? (L_ * L_ + 1 + 2 * TBL_PAD) * 2 : 0; }
  // exp(lp(S + dS) - lp(S)) for table index idx = 2 * row + q (row: S = S0 - 4 row; q: |dS| = 4 (q + 1)); rows outside
  // [0, L^2] belong to refuted guesses of the speculative sweep only
  __device__ __forceinline__ double ratio_at(int idx) const {
    if (TABLE) return tbl[idx];
    const int row_m = idx >> 1;
    if (row_m < 0 || row_m > L * L) return 1.0;
    const int s_old = S0 - 4 * row_m;
    const int s_new = s_old + sig * 4 * ((idx & 1) + 1);
    return exp_(lp(beta, s_new) - lp(beta, s_old));
  }
  static __device__ void stage_shared(const Params&, double*) {}
  __device__ __forceinline__ int own(int) const { return 0; }

  __device__ __forceinline__ double draw_uniform() {
    unsigned long long idx = rng.ctr - pool_base;
    if (!pool_valid || idx >= 32ull) {
      pool_base = rng.ctr;
      pool = uniform_at(rng, pool_base + (unsigned long long)lane);
      pool_valid = true;
      idx = 0;
    }
    rng.ctr += 1;
    return __shfl_sync(PGN_FULL_MASK, pool, (int)idx);
  }
  __device__ __forceinline__ void on_replica_changed() { pool_valid = false; }

  __device__ __forceinline__ int sgn(unsigned int r, int j) const { return ((r >> j) & 1u) ? 1 : -1; }
  __device__ void recompute_S() {   // examples/ising.jl:28-36
    const unsigned int up = __shfl_sync(PGN_FULL_MASK, row, (lane + L - 1) % L);
    const unsigned int dn = __shfl_sync(PGN_FULL_MASK, row, (lane + 1) % L);
    int s = 0;
    if (lane < L)
      for (int j = 0; j < L; ++j) {
        int jl = j == 0 ? L - 1 : j - 1, jr = j == L - 1 ? 0 : j + 1;
        s += sgn(row, j) * (sgn(up, j) + sgn(dn, j) + sgn(row, jl) + sgn(row, jr));
      }
    S = __reduce_add_sync(PGN_FULL_MASK, s) / 2;
  }
  __device__ void init(const Params& Pr, const double*, int wl, int lane_, int replica_index) {
    P = &Pr; lane = lane_; L = (int)Pr.p[1];
    beta = Pr.beta[Pr.first_chain + wl - 1];
    const unsigned int* rows = reinterpret_cast<const unsigned int*>(Pr.x + (size_t)wl * Pr.d_pad);
    row = lane < L ? rows[lane] : 0u;
    recompute_S();
    rng.key0 = Pr.seed_lo; rng.key1 = (unsigned int)replica_index; rng.c2 = Pr.seed_hi; rng.c3 = 0u;
    rng.ctr = Pr.rng_ctr[wl];
    expl_acc = MeanAcc{0, 0.0}; am = MeanAcc{0, 0.0}; rev = MeanAcc{0, 0.0};
    n_steps = n_points = n_ref = 0; err = 0;
    pool = 0.0; pool_base = 0; pool_valid = false;
  }
  __device__ void build_table(double* smem) {
    double* t = smem + (size_t)(threadIdx.x >> 5) * table_doubles(L);
    const double c = P->p[0];
    const bool decreasing = (beta > 0.0 && c < 0.0) || (beta < 0.0 && c > 0.0);   // lp decreasing in S
    sig = decreasing ? 1 : -1;
    S0 = 2 * L * L;
    if (!TABLE) { tbl = nullptr; return; }
    for (int idx = lane; idx < table_doubles(L); idx += 32) {
      const int row_m = (idx >> 1) - TBL_PAD;
      double v = 1.0;                                    // padding rows are never used by the true trajectory
      if (row_m >= 0 && row_m <= L * L) {
        const int s_old = S0 - 4 * row_m;
        const int s_new = s_old + sig * 4 * ((idx & 1) + 1);
        v = exp_(lp(beta, s_new) - lp(beta, s_old));
      }
      t[idx] = v;
    }
    t += 2 * TBL_PAD;
    tbl = t;
    __syncwarp();
  }
  __device__ void store(int wl) {
    unsigned int* rows = reinterpret_cast<unsigned int*>(P->x + (size_t)wl * P->d_pad);
    if (lane < L) rows[lane] = row;
  }
  __device__ void post(unsigned long long* pay, unsigned int tag) const { ll_store(pay + lane, row, tag); }
  __device__ bool adopt(const unsigned long long* pay, unsigned int tag) {
    const bool ok = ll_load(pay + lane, tag, row);
    recompute_S();
    return __all_sync(PGN_FULL_MASK, ok);
  }
  __device__ void share_state(double*) const {}
  __device__ void load_state(const double*) {}
  __device__ void write_trace(double* out) const {
    if (lane < L)
      for (int j = 0; j < L; ++j) out[lane * L + j] = ((row >> j) & 1u) ? 1.0 : 0.0;
  }
  __device__ void online_fit() {}
  __device__ void store_online() const {}
  __device__ void flush_online(OnEntry*) const {}
  __device__ void load_online(const OnEntry*) {}

  // InterpolatedLogPotential over two IsingLogPotential's (examples/ising.jl:74,77)
  __device__ __forceinline__ double lp(double b, int s) const {
    const double ref = 0.0 * (double)s;
    const double tgt = P->p[0] * (double)s;
    if (b == 0.0) return ref;
    if (b == 1.0) return tgt;
    return (1.0 - b) * ref + b * tgt;
  }
  __device__ void sample_iid() {   // examples/ising.jl:49-58
    const unsigned int mask = L >= 32 ? 0xffffffffu : ((1u << L) - 1u);
    row = lane < L ? (bits32_at(rng, rng.ctr + (unsigned long long)lane) & mask) : 0u;
    rng.ctr += (unsigned long long)L;
    recompute_S();
  }
  // One raster-order sweep site: flip (i, j) unless the Metropolis test rejects it.  `a` = neighbours equal
  // to the site (me * sum(neighbours) = 2a - 4, so S_new - S = 8 - 4a and the table row moves by a - 2);
  // `low` says the move lowers lp (only then can accept_ratio be < 1), `q` picks |dS| = 8.  Returns "flipped".
  __device__ __forceinline__ bool site(int a, bool low, int q, int& m, int& k) {
    if (k == 32) {   // next block of this replica's uniforms (same stream, see draw_uniform)
      pool_base += 32ull;
      pool = uniform_at(rng, pool_base + (unsigned long long)lane);
      k = 0;
    }
    const double accept_ratio = ratio_at(2 * m + q);
    const double u = __shfl_sync(PGN_FULL_MASK, pool, k);
    const bool draws = low && accept_ratio < 1;           // rand(rng) only if accept_ratio < 1 (examples/ising.jl:110)
    const bool reject = draws && u > accept_ratio;
    k += draws ? 1 : 0;
    m += reject ? 0 : a - 2;
    return !reject;
  }
  __device__ void metropolis() {   // examples/ising.jl:98-117
    // the sweep's draws come from the pooled stream; k = index of the next one inside the pool
    if (!pool_valid || rng.ctr - pool_base > 32ull) {
      pool_base = rng.ctr;
      pool = uniform_at(rng, pool_base + (unsigned long long)lane);
      pool_valid = true;
    }
    int k = (int)(rng.ctr - pool_base);
    int m = (S0 - S) >> 2;                                 // table row of the current S = S0 - 4 m
    const int jl_me = lane == 0 ? L - 1 : lane - 1, jr_me = lane >= L - 1 ? 0 : lane + 1;
    const bool down = sig < 0;                             // lp falls when S falls
    for (int sweep = 0; sweep < P->ising_n_steps; ++sweep)
      for (int i = 0; i < L; ++i) {
        const unsigned int up = __shfl_sync(PGN_FULL_MASK, row, (i + L - 1) % L);
        const unsigned int dn = __shfl_sync(PGN_FULL_MASK, row, (i + 1) % L);
        const unsigned int cur0 = __shfl_sync(PGN_FULL_MASK, row, i);
        // Lane j prepares site (i, j) from the row as it is now: its vertical neighbours are final, its right
        // neighbour is visited later, and its left neighbour is either untouched (f = 0) or flipped (f = 1) by
        // the time the sweep gets there — both cases are encoded, the sequential pass below only selects.
        unsigned int code;
        {
          const unsigned int me = (cur0 >> lane) & 1u;
          const int v = (int)((((cur0 ^ up) >> lane) & 1u) ^ 1u) + (int)((((cur0 ^ dn) >> lane) & 1u) ^ 1u);
          const int eL = (int)((((cur0 >> jl_me) & 1u) ^ me) ^ 1u), eR = (int)((((cur0 >> jr_me) & 1u) ^ me) ^ 1u);
          const int a0 = v + eL + eR, a1 = v + (1 - eL) + eR;
          const unsigned int c0 = (unsigned int)a0 | ((down ? a0 >= 3 : a0 <= 1) ? 8u : 0u) | ((a0 & 3) == 0 ? 16u : 0u);
          const unsigned int c1 = (unsigned int)a1 | ((down ? a1 >= 3 : a1 <= 1) ? 8u : 0u) | ((a1 & 3) == 0 ? 16u : 0u);
          code = c0 | (c1 << 8);
        }
        // The sites of a row are visited in order, each depending on whether its left neighbour just flipped, on
        // the running S and on how many uniforms were consumed so far.  Five sites at a time, lane h GUESSES the
        // five flip outcomes (bit t of h), walks the block under that guess — table row, draw index and
        // neighbour state all follow from the guess, so nothing waits for a comparison — and checks every
        // outcome against the guess.  Exactly one lane is consistent (the first wrong bit of any other guess is
        // refuted by the true outcome at that site): its state is the sequential sweep's.
        unsigned int flips = 0u;
        bool f = false;                                     // did the site to the left of the block flip?
        int j0 = 0;
        for (; j0 + 5 <= L - 1; j0 += 5) {
          if (k > 27) {   // room for five draws: restart the pool at the current tick of the stream
            pool_base += (unsigned long long)k;
            pool = uniform_at(rng, pool_base + (unsigned long long)lane);
            k = 0;
          }
          // straight-line on purpose: the five table reads and the five draws of a guess are independent
          unsigned int cc[5];
#pragma unroll
          for (int t = 0; t < 5; ++t) cc[t] = __shfl_sync(PGN_FULL_MASK, code, j0 + t);
          int row_at[5];
          bool low[5];
          int mm = m;                  // the guess's table row, walked through the block
          {
            bool fl = f;
#pragma unroll
            for (int t = 0; t < 5; ++t) {
              const unsigned int c = fl ? (cc[t] >> 8) : cc[t];
              const bool g = ((lane >> t) & 1) != 0;
              row_at[t] = 2 * mm + (int)((c >> 4) & 1u);
              low[t] = (c & 8u) != 0u;
              mm += g ? (int)(c & 7u) - 2 : 0;
              fl = g;
            }
          }
          double ratio[5];
#pragma unroll
          for (int t = 0; t < 5; ++t) ratio[t] = ratio_at(row_at[t]);
          int kk = k;
          bool ok = true;
#pragma unroll
          for (int t = 0; t < 5; ++t) {
            const bool draws = low[t] & (ratio[t] < 1);               // rand(rng) only if accept_ratio < 1 (examples/ising.jl:110)
            const double u = __shfl_sync(PGN_FULL_MASK, pool, kk);
            const bool flipped = !(draws & (u > ratio[t]));
            ok = ok & (flipped == (((lane >> t) & 1) != 0));
            kk += draws ? 1 : 0;
          }
          const int win = __ffs((int)__ballot_sync(PGN_FULL_MASK, ok)) - 1;
          m = __shfl_sync(PGN_FULL_MASK, mm, win);
          k = __shfl_sync(PGN_FULL_MASK, kk, win);
          flips |= (unsigned int)win << j0;
          f = ((win >> 4) & 1) != 0;
        }
        for (; j0 < L - 1; ++j0) {   // the sites left over by the blocks of five, one at a time
          unsigned int c = __shfl_sync(PGN_FULL_MASK, code, j0);
          c = f ? (c >> 8) : c;
          f = site((int)(c & 7u), (c & 8u) != 0u, (int)((c >> 4) & 1u), m, k);
          flips |= f ? (1u << j0) : 0u;
        }
        {   // last site of the row: its right neighbour is site 0 of the same row, already visited
          const int j = L - 1;
          const unsigned int cur = cur0 ^ flips;
          const int jl = j == 0 ? L - 1 : j - 1, jr = 0;
          const unsigned int me = (cur >> j) & 1u;
          const int a = (int)(((((cur0 ^ up) >> j) & 1u) ^ 1u) + ((((cur0 ^ dn) >> j) & 1u) ^ 1u) + ((((cur >> jl) ^ me) & 1u) ^ 1u) +
                              ((((cur >> jr) ^ me) & 1u) ^ 1u));
          f = site(a, down ? a >= 3 : a <= 1, (a & 3) == 0 ? 1 : 0, m, k);
          flips |= f ? (1u << j) : 0u;
        }
        if (lane == i) row = cur0 ^ flips;
      }
    S = S0 - 4 * m;
    rng.ctr = pool_base + (unsigned long long)k;
    n_ref += 2LL * P->ising_n_steps * L * L;
    n_points += (long long)P->ising_n_steps * L * L;
  }
  __device__ void explore(long long, bool is_reference) {
    if (is_reference) sample_iid(); else metropolis();
  }
  __device__ double log_ratio(double beta_partner) const { return lp(beta_partner, S) - lp(beta, S); }
};
using IsingChain = IsingChainT<true>;
using IsingChainLite = IsingChainT<false>;

// ===========================================================================
// TestSwapper (src/swap/pair_swapper.jl:100-149): no state, constant acceptance
// ===========================================================================
struct TestSwapperChain {
  static constexpr bool kTestSwapper = true;
  static constexpr bool kTeam = false;
  static constexpr bool kIsing = false;
  static constexpr bool kMixed = false;
  static constexpr int kMaxThreads = 256;
  static constexpr int kMinBlocksPerSM = 1;
  static constexpr bool kCompact = false;
  const Params* P;
  Rng rng;
  MeanAcc expl_acc, am, rev;
  long long n_steps, n_points, n_ref;
  int err;
  static __device__ void stage_shared(const Params&, double*) {}
  __device__ __forceinline__ int own(int) const { return 0; }
  __device__ void init(const Params& Pr, const double*, int wl, int, int replica_index) {
    P = &Pr;
    rng.key0 = Pr.seed_lo; rng.key1 = (unsigned int)replica_index; rng.c2 = Pr.seed_hi; rng.c3 = 0u;
    rng.ctr = Pr.rng_ctr[wl];
    expl_acc = MeanAcc{0, 0.0}; am = MeanAcc{0, 0.0}; rev = MeanAcc{0, 0.0};
    n_steps = n_points = n_ref = 0; err = 0;
  }
  __device__ void store(int) {}
  __device__ void post(unsigned long long*, unsigned int) const {}
  __device__ bool adopt(const unsigned long long*, unsigned int) { return true; }
  __device__ void share_state(double*) const {}
  __device__ void load_state(const double*) {}
  __device__ void write_trace(double*) const {}
  __device__ void online_fit() {}
  __device__ void store_online() const {}
  __device__ void flush_online(OnEntry*) const {}
  __device__ void load_online(const OnEntry*) {}
  __device__ void explore(long long, bool) {}
  __device__ double log_ratio(double) const { return 0.0; }
  __device__ __forceinline__ double draw_uniform() { return next_uniform(rng); }
  __device__ __forceinline__ void on_replica_changed() {}
};

// ===========================================================================
// The scan kernel
// ===========================================================================
template <class Chain>
__global__ void __launch_bounds__(Chain::kMaxThreads, Chain::kMinBlocksPerSM) scan_kernel(const __grid_constant__ Params P) {
  extern __shared__ double smem[];
  Chain::stage_shared(P, smem);
  __syncthreads();
  const int lane = threadIdx.x & 31;
  // team kernels (autoMALA): block = the W warps serving one chain; otherwise one warp per chain
  int tw = Chain::kTeam ? (int)(threadIdx.x >> 5) : 0;
  int wl = Chain::kTeam ? (int)blockIdx.x : (int)(blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5));
  int tW = Chain::kTeam ? (int)(blockDim.x >> 5) : 1;
  int team_slot = 0;   // mixed blocks: which of the block's two team regions in shared memory
  if constexpr (Chain::kMixed) {
    const int warp = (int)(threadIdx.x >> 5);
    const int c0 = P.block_map[2 * blockIdx.x], c1 = P.block_map[2 * blockIdx.x + 1];
    if (c1 == -2) { wl = c0; tw = warp; tW = 2; }                        // one chain, a team of two
    else { wl = warp == 0 ? c0 : c1; tw = 0; tW = 1; team_slot = warp; }   // two chains, one warp each
    if (wl < 0) return;
  }
  if (wl >= P.n_local) return;
  // team control words: [2] evaluation counter, [3] payload status, [8..23] partner header (two alternating
  // sets of 8), [24..] the adopted state for the other warps of the team
  long long* team_ctl = nullptr;
  const int N = P.n_chains;
  const int chain = P.first_chain + wl;   // 1-based global chain index
  const int last_local = P.first_chain + P.n_local - 1;
  int replica_index = P.replica_index[wl];
  int rt_state = 0;   // recorders are emptied at every round (recorders.jl:113-118)
  Chain ch;
  ch.init(P, smem, wl, lane, replica_index);
  if constexpr (Chain::kIsing) ch.build_table(smem);
  if constexpr (Chain::kTeam) {
    // mixed blocks reserve two single-warp regions; a team of two uses both (CTL + 6 slots <= 2 (CTL + 3 slots))
    double* team_base = smem + Chain::target_smem_doubles(P.d_pad) +
                        (size_t)team_slot * (Chain::TEAM_CTL_DOUBLES + 3 * Chain::SLOT_DOUBLES);
    team_ctl = reinterpret_cast<long long*>(team_base);
    double* slots = team_base + Chain::TEAM_CTL_DOUBLES;
    // the momentum pool follows the trial slots when the launch reserved room for it (P.pool_refresh > 0)
    ch.set_team(tw, tW, slots, P.pool_refresh > 0 ? slots + (size_t)3 * tW * Chain::SLOT_DOUBLES : nullptr);
    if (tw == 0 && lane == 0) team_ctl[2] = 0;
    ch.team_barrier();
  }
  MeanAcc swap_acc{0, 0.0};
  LogSumAcc ls_fwd{0, -PGN_INF}, ls_bwd{0, -PGN_INF};
  long long n_restarts = 0, n_trips = 0;
  long long explore_cycles = 0, wait_cycles = 0;
  const bool is_ref = chain_is_reference(P, chain);   // DEO.jl:13, VariationalDEO.jl:19
  const bool is_tgt = chain_is_target(P, chain);      // DEO.jl:14, VariationalDEO.jl:20
  // two legs: the two target chains are neighbours; a replica crossing between them keeps fitting the SAME online
  // accumulators (its own recorder, pigeons.jl:110-131), so the row travels A -> table -> B under the hand-shake's order
  int err = 0;

  for (long long scan = 1; scan <= P.n_scans; ++scan) {
    // ---------------- explore ----------------
    const long long t_explore0 = clock64();
    ch.explore(scan, is_ref);
    explore_cycles += clock64() - t_explore0;
    if (ch.err) { err = ch.err; break; }
    if (is_tgt) {   // pigeons.jl:110-131
      if (tw == ch.own(4)) ch.online_fit();
      if (P.target_trace && tw == 0)   // two legs: [scan][2][d], chain n_var then chain n_var + 1
        ch.write_trace(P.target_trace + (two_legs(P) ? (size_t)(scan - 1) * 2 + (chain == P.n_var + 1 ? 1 : 0) : (size_t)(scan - 1)) * P.d);
    }
    // ---------------- swap ----------------
    const bool even = (scan & 1LL) == 0;                                   // DEO.jl:12
    int partner = chain + ((((chain & 1) == 0) == even) ? 1 : -1);         // OddEven.jl:23-31
    if (partner == 0) partner = 1;
    if (partner == N + 1) partner = N;
    double lr = 0.0;
    if (!Chain::kTestSwapper) {
      lr = ch.log_ratio(P.beta[partner - 1]);                              // swap_stat, pair_swapper.jl:42-47
      ch.n_ref += 2;
      if (lr != lr) { err = PGN_ERR_NAN_RATIO; break; }
    }
    const double u = ch.draw_uniform();
    const size_t log_at = (size_t)(scan - 1) * P.n_local + wl;
    if (lane == 0 && tw == 0) {   // recorded before the swap (swap.jl:110-111)
      if (P.index_process) P.index_process[log_at] = replica_index;
      if (P.swap_lr) P.swap_lr[log_at] = lr;
      if (P.swap_u) P.swap_u[log_at] = u;
    }
    if (rt_state == 0 && is_ref) rt_state = 1;                             // RoundTripRecorder.jl:43-54
    else if (rt_state == 1 && is_tgt) { rt_state = 2; n_restarts += 1; }
    else if (rt_state == 2 && is_ref) { rt_state = 1; n_trips += 1; }

    bool accepted = false;
    if (partner != chain) {
      const int ring = (int)((P.epoch & 1u) * 4u + (unsigned int)(scan & 3LL));
      const unsigned int tag = P.tag_base + (unsigned int)scan;
      const bool remote = partner < P.first_chain || partner > last_local;
      // ---- where I post my SwapStat + replica, and where my partner posts theirs
      char* dst;
      const char* src;
      if (!remote) {
        dst = P.mail + ((size_t)(2 + wl) * MAIL_RINGS + ring) * P.slot_bytes;
        src = P.mail + ((size_t)(2 + (partner - P.first_chain)) * MAIL_RINGS + ring) * P.slot_bytes;
      } else if (partner > chain) {   // partner lives on the right neighbour: I am its LEFT ghost
        dst = P.mail_right + ((size_t)0 * MAIL_RINGS + ring) * P.slot_bytes;
        src = P.mail + ((size_t)1 * MAIL_RINGS + ring) * P.slot_bytes;
      } else {
        dst = P.mail_left + ((size_t)1 * MAIL_RINGS + ring) * P.slot_bytes;
        src = P.mail + ((size_t)0 * MAIL_RINGS + ring) * P.slot_bytes;
      }
      unsigned long long* dstw = reinterpret_cast<unsigned long long*>(dst);
      const unsigned long long* srcw = reinterpret_cast<const unsigned long long*>(src);
      int status = 0;
      double lr_p = 0.0, u_p = 0.0;
      unsigned long long ctr_p = 0ull;
      int ri_p = 0, rt_p = 0;
      const bool mid = two_legs(P) && ((chain == P.n_var && partner == chain + 1) || (chain == P.n_var + 1 && partner == chain - 1));
      if (mid && P.on_table != nullptr && P.d > 0) {
        // the outgoing replica's online row must be in memory before the partner can see this scan's header:
        // rows -> (team barrier) -> fence -> tagged words; the partner fences after it has seen the header
        if (tw == ch.own(4)) ch.flush_online(P.on_table + (size_t)(replica_index - 1) * P.d_pad);
        if constexpr (Chain::kTeam) ch.team_barrier(); else __syncwarp();
        if (tw == 0) fence_acq_rel_gpu();
      }
      const long long t_wait0 = clock64();
      if (tw == 0) {
        {   // post: header words from lanes 0..7, then the replica
          const unsigned long long lb = double_to_bits(lr), ub = double_to_bits(u), cb = ch.rng.ctr;
          unsigned int hv = (unsigned int)lb;
          hv = lane == 1 ? (unsigned int)(lb >> 32) : hv;
          hv = lane == 2 ? (unsigned int)ub : hv;
          hv = lane == 3 ? (unsigned int)(ub >> 32) : hv;
          hv = lane == 4 ? (unsigned int)cb : hv;
          hv = lane == 5 ? (unsigned int)(cb >> 32) : hv;
          hv = lane == 6 ? (unsigned int)replica_index : hv;
          hv = lane == 7 ? (unsigned int)rt_state : hv;
          if (lane < LL_HDR_WORDS) ll_store(dstw + lane, hv, tag);
          ch.post(dstw + LL_HDR_WORDS, tag);
        }
        // ---- wait for the partner's header
        unsigned long long w = 0ull;
        unsigned long long t0 = 0, seen = 0;
        unsigned int it = 0;
        while (true) {
          if (lane < LL_HDR_WORDS) w = ld_relaxed_sys(srcw + lane);
          const bool ok = lane >= LL_HDR_WORDS || (unsigned int)(w >> 32) == tag;
          if (__all_sync(PGN_FULL_MASK, ok)) break;
          ++it;
          if (it > 4u) __nanosleep(it < 64u ? 32 : 256);   // a sleeping warp leaves its scheduler's issue slots to the working ones
          if ((it & 255u) == 0u) {
            if (lane == 0) {
              if (*reinterpret_cast<volatile int*>(P.error_flag) != 0) status = 1;
              const unsigned long long now = globaltimer_ns();
              const unsigned long long prog = *reinterpret_cast<volatile unsigned long long*>(P.progress);
              if (t0 == 0 || prog != seen) { t0 = now; seen = prog; }   // some chain of the shard finished a scan: not stuck
              else if (now - t0 > P.timeout_ns) status = 2;
            }
            status = __shfl_sync(PGN_FULL_MASK, status, 0);
            if (status != 0) break;
          }
        }
        const unsigned int lo = (unsigned int)w;
        const unsigned int h0 = __shfl_sync(PGN_FULL_MASK, lo, 0), h1 = __shfl_sync(PGN_FULL_MASK, lo, 1);
        const unsigned int h2 = __shfl_sync(PGN_FULL_MASK, lo, 2), h3 = __shfl_sync(PGN_FULL_MASK, lo, 3);
        const unsigned int h4 = __shfl_sync(PGN_FULL_MASK, lo, 4), h5 = __shfl_sync(PGN_FULL_MASK, lo, 5);
        ri_p = (int)__shfl_sync(PGN_FULL_MASK, lo, 6);
        rt_p = (int)__shfl_sync(PGN_FULL_MASK, lo, 7);
        lr_p = bits_to_double(((unsigned long long)h1 << 32) | h0);
        u_p = bits_to_double(((unsigned long long)h3 << 32) | h2);
        ctr_p = ((unsigned long long)h5 << 32) | h4;
        if (mid) { __syncwarp(); fence_acq_rel_gpu(); }   // the partner's online row (written before its header) is visible from here on
      }
      if constexpr (Chain::kTeam) {   // team warp 0 did the hand-shake and hands the header to the team
        long long* tc = team_ctl + 8 + 8 * (int)(scan & 1LL);
        if (tw == 0 && lane == 0) {
          tc[0] = status; tc[1] = (long long)double_to_bits(lr_p); tc[2] = (long long)double_to_bits(u_p);
          tc[3] = (long long)ctr_p; tc[4] = ri_p; tc[5] = rt_p;
        }
        ch.team_barrier();
        status = (int)tc[0];
        lr_p = bits_to_double((unsigned long long)tc[1]); u_p = bits_to_double((unsigned long long)tc[2]);
        ctr_p = (unsigned long long)tc[3]; ri_p = (int)tc[4]; rt_p = (int)tc[5];
      }
      wait_cycles += clock64() - t_wait0;
      if (status != 0) { err = status == 2 ? PGN_ERR_TIMEOUT : -1; break; }
      // ---- decision (identical on both sides; swap_decision pair_swapper.jl:81-88)
      const bool lower = chain < partner;
      double acceptance_pr;
      if (Chain::kTestSwapper) {
        acceptance_pr = P.p[0];
      } else {
        const double e = lower ? exp_(lr + lr_p) : exp_(lr_p + lr);
        acceptance_pr = 1.0 < e ? 1.0 : e;
        if (lower) {   // record_swap_stats! :59-66, by the replica holding the lower chain
          if (tw == ch.own(1)) swap_acc.template fit<Chain::kCompact>(acceptance_pr);
          if (tw == ch.own(2)) ls_fwd.template fit<Chain::kCompact>(lr);
          if (tw == ch.own(3)) ls_bwd.template fit<Chain::kCompact>(lr_p);
        }
      }
      accepted = (lower ? u : u_p) < acceptance_pr;
      if (accepted) {   // adopt the partner's replica (states move, chains stay)
        if (P.rec_table != nullptr) {
          // per-replica recorders: what this chain's warps accumulated belongs to the outgoing replica; continue with
          // what the incoming replica accumulated during its earlier visits of this chain (each statistic by its owner warp)
          RecEntry* eo = P.rec_table + (size_t)(replica_index - 1) * P.n_local + wl;
          const RecEntry* en = P.rec_table + (size_t)(ri_p - 1) * P.n_local + wl;
          if (tw == ch.own(1)) {
            if (lane == 0) { eo->am = ch.am; eo->swap_acc = swap_acc; }
            ch.am = en->am; swap_acc = en->swap_acc;
          }
          if (tw == ch.own(2)) {
            if (lane == 0) { eo->rev = ch.rev; eo->ls_fwd = ls_fwd; }
            ch.rev = en->rev; ls_fwd = en->ls_fwd;
          }
          if (tw == ch.own(3)) {
            if (lane == 0) { eo->expl_acc = ch.expl_acc; eo->ls_bwd = ls_bwd; }
            ch.expl_acc = en->expl_acc; ls_bwd = en->ls_bwd;
          }
          if (is_tgt && tw == ch.own(4) && P.d > 0) {
            if (!mid) ch.flush_online(P.on_table + (size_t)(replica_index - 1) * P.d_pad);   // mid: done before the post
            ch.load_online(P.on_table + (size_t)(ri_p - 1) * P.d_pad);
          }
        }
        replica_index = ri_p;
        rt_state = rt_p;
        ch.rng.ctr = ctr_p;
        ch.rng.key1 = (unsigned int)replica_index;
        ch.on_replica_changed();
        bool got = true;
        if (tw == 0) got = ch.adopt(srcw + LL_HDR_WORDS, tag);
        if constexpr (Chain::kTeam) {
          double* sx = reinterpret_cast<double*>(team_ctl + 24);
          if (tw == 0) { ch.share_state(sx); if (lane == 0) team_ctl[3] = got ? 1 : 0; }
          ch.team_barrier();
          if (tw != 0) ch.load_state(sx);
          got = team_ctl[3] != 0;
        }
        if (!got) { err = PGN_ERR_TIMEOUT; break; }
      }
    }
    if (lane == 0 && tw == 0) {
      if (P.swap_accept) P.swap_accept[log_at] = accepted ? 1 : 0;
      atomicAdd(P.progress, 1ull);
    }
  }

  if (err > 0 && lane == 0) atomicCAS(P.error_flag, 0, err);
  if (P.rec_table != nullptr) {   // per-replica recorders: the entry of the replica each chain holds at the end of the round
    RecEntry* eo = P.rec_table + (size_t)(replica_index - 1) * P.n_local + wl;
    if (lane == 0) {
      if (tw == ch.own(1)) { eo->am = ch.am; eo->swap_acc = swap_acc; }
      if (tw == ch.own(2)) { eo->rev = ch.rev; eo->ls_fwd = ls_fwd; }
      if (tw == ch.own(3)) { eo->expl_acc = ch.expl_acc; eo->ls_bwd = ls_bwd; }
    }
    if (is_tgt && tw == ch.own(4) && P.d > 0) ch.flush_online(P.on_table + (size_t)(replica_index - 1) * P.d_pad);
  }
  // ---------------- epilogue: replica back to HBM, statistics out ----------------
  if constexpr (Chain::kTeam) {   // collect what the other warps of the team hold: their statistics and evaluation counts
    double* ex = reinterpret_cast<double*>(team_ctl + 24);
    long long* exl = reinterpret_cast<long long*>(ex);
    if (is_tgt && tw == ch.own(4)) ch.store_online();
    if (lane == 0) {
      if (tw != 0) atomicAdd(reinterpret_cast<unsigned long long*>(team_ctl + 2), (unsigned long long)ch.n_points);
      if (tw == ch.own(1)) { exl[0] = ch.am.n; ex[1] = ch.am.mu; exl[2] = swap_acc.n; ex[3] = swap_acc.mu; }
      if (tw == ch.own(2)) { exl[4] = ch.rev.n; ex[5] = ch.rev.mu; exl[6] = ls_fwd.n; ex[7] = ls_fwd.value; }
      if (tw == ch.own(3)) { exl[8] = ch.expl_acc.n; ex[9] = ch.expl_acc.mu; exl[10] = ls_bwd.n; ex[11] = ls_bwd.value; }
    }
    ch.team_barrier();
    if (tw != 0) return;
    ch.n_points += team_ctl[2];
    ch.am.n = exl[0]; ch.am.mu = ex[1]; swap_acc.n = exl[2]; swap_acc.mu = ex[3];
    ch.rev.n = exl[4]; ch.rev.mu = ex[5]; ls_fwd.n = exl[6]; ls_fwd.value = ex[7];
    ch.expl_acc.n = exl[8]; ch.expl_acc.mu = ex[9]; ls_bwd.n = exl[10]; ls_bwd.value = ex[11];
  }
  ch.store(wl);
  if (is_tgt && !Chain::kTeam) ch.store_online();
  if (lane == 0) {
    P.replica_index[wl] = replica_index;
    P.rng_ctr[wl] = ch.rng.ctr;
    P.rt_state[wl] = rt_state;
    ChainStatsDev s;
    s.swap_n = swap_acc.n; s.swap_mean = swap_acc.mu; s.ls_fwd = ls_fwd.value; s.ls_bwd = ls_bwd.value;
    s.expl_acc_n = ch.expl_acc.n; s.expl_acc_mean = ch.expl_acc.mu; s.n_steps = ch.n_steps;
    s.am_n = ch.am.n; s.am_mean = ch.am.mu; s.rev_n = ch.rev.n; s.rev_mean = ch.rev.mu;
    s.n_restarts = n_restarts; s.n_round_trips = n_trips;
    s.n_points = ch.n_points; s.n_ref_evals = ch.n_ref;
    s.explore_cycles = explore_cycles; s.wait_cycles = wait_cycles;
    if constexpr (Chain::kTeam) { s.trial_cycles = ch.t_trial; s.barrier_cycles = ch.t_barrier; s.decide_cycles = ch.t_decide; }
    else { s.trial_cycles = s.barrier_cycles = s.decide_cycles = 0; }
    P.stats[wl] = s;
  }
}

// ===========================================================================
// Parity entry points (the other small helper kernels live in pgn_scan_misc.cu)
// ===========================================================================
// parity entry points: one warp per point
template <int TK, int CPL, bool VAR = false>
__global__ void eval_points_kernel(const __grid_constant__ Params P, const double* xs, const double* betas, int n_points, double* lp_out,
                                   double* ld_out, double* grad_out) {
  extern __shared__ double smem[];
  typedef VecChain<TK, CPL, PGN_EXPLORER_SLICE, VAR> Chain;
  Chain::stage_shared(P, smem);
  __syncthreads();
  const int lane = threadIdx.x & 31;
  const int w = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (w >= n_points) return;
  Chain ch;
  ch.P = &P; ch.sm_means = smem; ch.lane = lane; ch.d = P.d; ch.beta = betas[w];
  ch.n_points = 0; ch.n_ref = 0; ch.err = 0;
  ch.var_ref = VAR && P.var_tab != nullptr;   // the path of chain 1's leg
  double xx[CPL];
#pragma unroll
  for (int k = 0; k < CPL; ++k) xx[k] = (k * 32 + lane < P.d) ? xs[(size_t)w * P.d + k * 32 + lane] : 0.0;
  if (lp_out) {
    double a0, a1;
    ch.eval(xx, a0, a1);
    if (lane == 0) lp_out[w] = ch.lp_call(ch.beta, a0, a1);
  }
  if (ld_out) {
    double a0, a1, g[CPL], extra = 0.0;
    ch.eval_grad(xx, ch.beta, a0, a1, g, extra);
    // logdens = 0.0 + l_ref (1-b) + l_tgt b (BufferedAD.jl:99-108); toy: ScaledPrecisionNormalPath.jl:30-34
    if (lane == 0) ld_out[w] = (TK == PGN_TARGET_TOY_MVN) ? ch.lp_ad(ch.beta, a0, a1)
                                                          : ((0.0 + a0 * (1.0 - ch.beta)) + a1 * ch.beta);
#pragma unroll
    for (int k = 0; k < CPL; ++k)
      if (k * 32 + lane < P.d) grad_out[(size_t)w * P.d + k * 32 + lane] = g[k];
  }
}

// hamiltonian_dynamics! with the identity preconditioner: one warp per point, n_steps x VecChain::run_trial (= leap_frog!)
template <int TK, int CPL, bool VAR = false>
__global__ void leapfrog_kernel(const __grid_constant__ Params P, const double* xs, const double* ps, const double* betas,
                                const double* precond, double eps, int n_steps, int n_points, double* x_out, double* p_out) {
  extern __shared__ double smem[];
  typedef VecChain<TK, CPL, PGN_EXPLORER_MALA, VAR> Chain;
  Chain::stage_shared(P, smem);
  __syncthreads();
  const int lane = threadIdx.x & 31;
  const int w = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (w >= n_points) return;
  Chain ch;
  ch.P = &P; ch.sm_means = smem; ch.lane = lane; ch.d = P.d; ch.beta = betas[w];
  ch.n_points = 0; ch.n_ref = 0; ch.err = 0;
  ch.var_ref = VAR && P.var_tab != nullptr;
  double x[CPL], p[CPL], g[CPL], pre[CPL];
#pragma unroll
  for (int k = 0; k < CPL; ++k) {
    const bool v = k * 32 + lane < P.d;
    x[k] = v ? xs[(size_t)w * P.d + k * 32 + lane] : 0.0;
    p[k] = v ? ps[(size_t)w * P.d + k * 32 + lane] : 0.0;
    pre[k] = (v && precond != nullptr) ? precond[k * 32 + lane] : 1.0;
  }
  const bool pre_one = precond == nullptr;
  {   // conditioned gradient at the start point (hamiltonian_dynamics.jl:31-35, 45-47)
    double a0, a1, extra = 0.0;
    ch.eval_grad(x, ch.beta, a0, a1, g, extra);
#pragma unroll
    for (int k = 0; k < CPL; ++k) g[k] = pre_one ? g[k] : g[k] / pre[k];
  }
  typename Chain::Trial T;
  for (int s = 0; s < n_steps; ++s) {
    ch.run_trial(x, p, g, pre, pre_one, eps, 0.0, T);
#pragma unroll
    for (int k = 0; k < CPL; ++k) { x[k] = T.x1[k]; p[k] = T.p1[k]; g[k] = T.g1c[k]; }
    if (!is_finite(T.h_after)) break;   // :56-59, :80: the reference stops at a non-finite state
  }
#pragma unroll
  for (int k = 0; k < CPL; ++k)
    if (k * 32 + lane < P.d) { x_out[(size_t)w * P.d + k * 32 + lane] = x[k]; p_out[(size_t)w * P.d + k * 32 + lane] = p[k]; }
}

}  // namespace pgn
