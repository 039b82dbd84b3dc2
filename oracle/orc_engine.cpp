// ORACLE (test infrastructure only).  CPU restatement of the Pigeons.jl inner
// PT scan — run_one_round! = n_scans x { explore!(all replicas) ; swap! } —
// written to follow the reference's control flow statement by statement
// (including its redundant density evaluations), with the arithmetic of
// orc_math.hpp.  Each function cites the reference file:line it restates
// (paths relative to /root/reference).
//
// Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
// --impl reference leg may load this library.  It is never on the product path.
//
// parity unpinned: no golden vectors exist in the reference for this path and
// the reference cannot run here (no Julia).  What pins this file is
// (1) the reference's tolerance-based known answers (tests/test_oracle_known_answers.py)
// and (2) the exact-integer DEO/round-trip answer of test/test_round_trips.jl.
//
// Deliberate, documented deviations from Pigeons.jl:
//   * RNG: Philox4x32-10 keyed by (seed, replica_index), one block per draw,
//     instead of SplittableRandom + ziggurat (north star).
//   * sums over coordinates use the canonical 32-leaf tree (orc_math.hpp)
//     instead of Julia's pairwise/SIMD `sum` (unreproducible).
//   * (recorder_order = PGN_RECORDERS_PER_CHAIN only) recorder statistics accumulated per
//     chain / per pair in scan order; the default follows the reference: per replica,
//     then the replica-index tree merge of reduce_recorders! (recorders.jl:88-120).
#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <memory>
#include <string>
#include <unordered_map>
#include <vector>

#include "../include/pigeons_b200.h"
#include "orc_math.hpp"

#ifdef _OPENMP
#include <omp.h>
#endif

using namespace orc;

namespace {

struct OrcError {
  int code;
  std::string msg;
};

// ---------------------------------------------------------------------------
// Replica record (src/replicas/Replica.jl:5-30)
// ---------------------------------------------------------------------------
struct Replica {
  std::vector<double> x;     // state
  std::vector<uint32_t> rows;  // Ising: bit-packed rows (bit j of rows[i] = spin (i,j))
  int ising_S = 0;           // Ising: cached sum_pair_products (examples/ising.jl:19)
  int chain = 0;             // 1-based
  int replica_index = 0;     // 1-based
  Philox rng{};
  uint64_t ctr = 0;          // draws consumed
  int rt_state = 0;          // RoundTripRecorder.state
  // per-scan SwapStat (src/swap/pair_swapper.jl:8-11)
  double lr = 0.0, u = 0.0;
  // autoMALA buffers (src/explorers/Augmentation.jl:65-71)
  std::vector<double> momentum, precond, start_state, state_before, momentum_before, grad, g1, g2;
  // counters local to the replica during explore (merged after the parallel loop)
  int64_t ref_equiv_evals = 0;

  double uniform() { return uniform_at(rng, ctr++); }
  double exponential() { return exponential_at(rng, ctr++); }
};

struct MeanAcc {   // OnlineStatsBase.Mean with EqualWeight
  int64_t n = 0;
  double mu = 0.0;
  void fit(double x) { n += 1; mu = mu + (1.0 / (double)n) * (x - mu); }
  // OnlineStatsBase._merge!(::Mean, ::Mean): o.n += o2.n; o.mu = smooth(o.mu, o2.mu, o2.n / o.n), smooth(a, b, g) = a + g (b - a).
  // A statistic that was never fitted does not exist in the reference's GroupBy: merging with it is a copy.
  void merge(const MeanAcc& o) {
    if (o.n == 0) return;
    if (n == 0) { *this = o; return; }
    n += o.n;
    mu = mu + ((double)o.n / (double)n) * (o.mu - mu);
  }
};
struct LogSumAcc {  // src/recorders/LogSum.jl:1-24
  int64_t n = 0;
  double value = -INF;
  void fit(double y) { value = logaddexp_(value, y); n += 1; }
  void merge(const LogSumAcc& o) {   // LogSum.jl:15-18
    if (o.n == 0) return;
    if (n == 0) { *this = o; return; }
    value = logaddexp_(value, o.value);
    n += o.n;
  }
};
struct VarAcc {   // OnlineStatsBase.Variance with EqualWeight
  int64_t n = 0;
  double mu = 0.0, s2 = 0.0;
  void fit(double x) {
    double mu_old = mu;
    n += 1;
    double g = 1.0 / (double)n;
    mu = mu_old + g * (x - mu_old);
    s2 = s2 + g * ((x - mu) * (x - mu_old) - s2);
  }
  double value() const { return n > 1 ? s2 * ((double)n / (double)(n - 1)) : 1.0; }
  // OnlineStatsBase._merge!(::Variance, ::Variance): g = o2.n / (o.n += o2.n); delta = o2.mu - o.mu;
  // o.s2 = smooth(o.s2, o2.s2, g) + delta^2 g (1 - g); o.mu = smooth(o.mu, o2.mu, g)
  void merge(const VarAcc& o) {
    if (o.n == 0) return;
    if (n == 0) { *this = o; return; }
    n += o.n;
    const double g = (double)o.n / (double)n;
    const double delta = o.mu - mu;
    s2 = (s2 + g * (o.s2 - s2)) + ((delta * delta) * g) * (1.0 - g);
    mu = mu + g * (o.mu - mu);
  }
};

struct ChainStats {
  MeanAcc swap_acc;           // pair (c, c+1), stored at the lower chain
  LogSumAcc ls_fwd, ls_bwd;   // keys (c,c+1) and (c+1,c)
  MeanAcc expl_acc;           // explorer_acceptance_pr
  int64_t n_steps = 0;        // explorer_n_steps (Sum)
  MeanAcc am;                 // am_factors
  MeanAcc rev;                // reversibility_rate
  void merge(const ChainStats& o) {   // merge of the GroupBy entries of one key (Sum: integer addition)
    swap_acc.merge(o.swap_acc); ls_fwd.merge(o.ls_fwd); ls_bwd.merge(o.ls_bwd);
    expl_acc.merge(o.expl_acc); n_steps += o.n_steps; am.merge(o.am); rev.merge(o.rev);
  }
};

// The recorders ONE replica carries during a round (src/recorders/recorders.jl:1-12: "each recorders object keeps track
// of only the statistics for one replica"), restricted to the statistics of this path: the GroupBy recorders keyed by
// chain (pair statistics under the pair's lower chain) and the target-chain online statistics.
struct ReplicaRecorders {
  std::unordered_map<int, ChainStats> by_chain;   // key: chain - 1
  std::vector<VarAcc> online;                     // _transformed_online; empty until the replica visits the target chain
};
// merge_recorders (recorders.jl:122-131) -> Base.merge of every recorder: GroupBy merges entry by entry, inserting the
// keys the left side lacks; OnlineStateRecorder.merge (OnlineStateRecorder.jl:44-58) copies when one side is empty.
void merge_recorders(ReplicaRecorders& a, const ReplicaRecorders& b) {
  for (const auto& kv : b.by_chain) {
    auto it = a.by_chain.find(kv.first);
    if (it == a.by_chain.end()) a.by_chain.emplace(kv.first, kv.second);
    else it->second.merge(kv.second);
  }
  if (a.online.empty()) a.online = b.online;
  else if (!b.online.empty())
    for (size_t c = 0; c < a.online.size(); ++c) a.online[c].merge(b.online[c]);
}

// Explore-phase events are produced inside the (possibly threaded) replica
// loop; they only touch the stats slot of the replica's own chain, and every
// chain is held by exactly one replica, so there is no sharing.

struct Engine {
  pgn_config cfg{};
  std::vector<double> means, log_w;
  std::vector<double> data_x, data_y;   // LOGREG: X [n_data][d] row-major, y [n_data]
  pgn_explorer_params ep{};
  std::vector<double> std_devs;
  bool have_std = false;
  std::vector<double> beta;            // schedule, size N
  std::vector<Replica> replicas;       // sorted by chain between scans
  std::vector<ChainStats> stats;       // per chain (index chain-1): the reduced recorders
  std::vector<ReplicaRecorders> rec;   // per replica (index replica_index-1), PGN_RECORDERS_PER_REPLICA
  bool per_replica() const { return cfg.recorder_order == PGN_RECORDERS_PER_REPLICA; }
  // the recorder entry the replica writes to: its own (keyed by its current chain), or the chain's shared one
  ChainStats& st_of(const Replica& r) {
    return per_replica() ? rec[r.replica_index - 1].by_chain[r.chain - 1] : stats[r.chain - 1];
  }
  int64_t n_restarts = 0, n_round_trips = 0;
  std::vector<VarAcc> online;          // per dim, target chain
  int scan = 0;
  int n_threads = 1;

  int N() const { return cfg.n_chains; }
  int d() const { return cfg.dim; }
  int L() const { return (int)cfg.p[1]; }

  // ---- legs (StabilizedPT.jl:37-66, 96-116; VariationalDEO.jl:19-20): chains 1..n_var are the variational leg
  // (reference -> target), chains n_var+1..N the fixed leg (target -> reference)
  int n_var() const { return cfg.n_chains_variational; }
  bool two_legs() const { return n_var() > 0 && n_var() < N(); }
  bool is_reference(int chain) const { return (chain == 1 && N() > 1) || (two_legs() && chain == N()); }   // DEO.jl:13
  bool is_target(int chain) const { return two_legs() ? (chain == n_var() || chain == n_var() + 1) : chain == N(); }
  // GaussianReference (GaussianReference.jl:4-54), once activated (:16-18), replaces the reference of the variational leg
  bool var_active = false;
  std::vector<double> var_mean, var_sd, var_t0, var_t1, var_t2;
  bool uses_var(int chain) const { return var_active && chain <= n_var(); }
  // gaussian_logdensity :45-53: sum_i -0.5 log(2 pi sd_i^2) - 1/(2 sd_i^2) (x_i - mean_i)^2; the two coordinate
  // constants are tabulated when the reference is set
  double var_density(const double* x) const {
    return tree_sum(d(), [&](int c) { double dx = x[c] - var_mean[c]; return var_t0[c] - var_t1[c] * (dx * dx); });
  }

  // ------------------------------------------------------------------ densities
  // Component log densities.  FUNNEL: test/supporting/dimensional-analysis.jl:33-47
  // with Distributions.logpdf(Normal(0,s),x) = -(z^2+log2pi)/2 - log(s) and
  // s = exp(y/2) => z^2 = x^2 exp(-y), log s = y/2.
  double ref_density(const Replica& r) const {
    if (uses_var(r.chain)) return var_density(r.x.data());
    return fixed_ref_density(r);
  }
  double fixed_ref_density(const Replica& r) const {
    const double* x = r.x.data();
    switch (cfg.target_kind) {
      case PGN_TARGET_FUNNEL:
      case PGN_TARGET_GMM:
      case PGN_TARGET_LOGREG: {
        const double iv = cfg.p[5], ls = cfg.p[4];
        return tree_sum(d(), [&](int c) { return -(x[c] * x[c] * iv + LOG2PI) * 0.5 - ls; });
      }
      case PGN_TARGET_ISING:
        return 0.0 * (double)r.ising_S;   // IsingLogPotential(0.0, L) (examples/ising.jl:74,77)
      case PGN_TARGET_MIXED: return mixed_density(x, 0);
      case PGN_TARGET_UNID:   // logpdf(product_distribution([Uniform(), Uniform()]), x): 0 inside the unit square, -Inf outside
        return ((x[0] >= 0.0 && x[0] <= 1.0) ? 0.0 : -INF) + ((x[1] >= 0.0 && x[1] <= 1.0) ? 0.0 : -INF);
      default: return QNAN;
    }
  }
  // ---- mixed Bool / Integer / Float product target (the state of test/test_slice_sampler.jl:56-75): side 0 = reference
  // Bernoulli(p0) x Binomial(n, q0) x Normal(0, s), side 1 = target Bernoulli(p1) x Binomial(n, q1) x Normal(0, 1).
  // Distributions.logpdf of a discrete distribution at a non-integer or out-of-support point is -Inf.
  int mixed_nb() const { return (int)cfg.p[0]; }
  int mixed_ni() const { return (int)cfg.p[1]; }
  int mixed_kind(int c) const { return c < mixed_nb() ? 0 : (c < mixed_nb() + mixed_ni() ? 1 : 2); }   // 0 Bool, 1 Integer, 2 Float
  double mixed_term(int c, double v, int side) const {
    const double* t = means.data();
    const int n = (int)cfg.p[2];
    switch (mixed_kind(c)) {
      case 0: return v == 1.0 ? t[2 * side] : (v == 0.0 ? t[2 * side + 1] : -INF);
      case 1: {
        if (!(v >= 0.0 && v <= (double)n) || v != std::floor(v)) return -INF;
        const int k = (int)v;
        return (t[10 + k] + (double)k * t[4 + 2 * side]) + (double)(n - k) * t[5 + 2 * side];
      }
      default:
        if (side == 1) return -(v * v + LOG2PI) * 0.5;
        return -(v * v * cfg.p[5] + LOG2PI) * 0.5 - cfg.p[4];
    }
  }
  double mixed_density(const double* x, int side) const {
    return tree_sum(d(), [&](int c) { return mixed_term(c, x[c], side); });
  }
  // ---- logistic regression (BASELINE config 5; analytic-gradient target in the style of
  // test/test_custom_gradient.jl:1-33): target = N(0, s^2 I) prior x Bernoulli(sigmoid(X theta)).
  // Summation orders are part of the arithmetic spec (they are what the device GEMMs do):
  //   z_n      : sequential fma over c = 0..d-1 starting from 0.0
  //   sum_n ll : rows in tiles of 128; inside a tile the canonical 32-lane tree
  //              (lane = row % 32); tile sums added in tile order
  //   G[c]     : rows in chunks of 4096, sequential fma inside a chunk starting from 0.0,
  //              chunk partials added in chunk order
  static constexpr int LR_TILE = 128, LR_CHUNK = 4096;
  int n_data() const { return (int)cfg.p[0]; }
  void logreg_z(const double* x, std::vector<double>& z) const {
    const int n = n_data(), dd = d();
    z.assign(n, 0.0);
    // each z_i is ONE sequential fma chain over c; eight rows are interleaved only to give the
    // CPU independent chains to pipeline (the per-row order is unchanged)
    int i = 0;
    for (; i + 8 <= n; i += 8) {
      const double* row = &data_x[(size_t)i * dd];
      double a[8] = {0, 0, 0, 0, 0, 0, 0, 0};
      for (int c = 0; c < dd; ++c) {
        const double xc = x[c];
        for (int j = 0; j < 8; ++j) a[j] = std::fma(row[(size_t)j * dd + c], xc, a[j]);
      }
      for (int j = 0; j < 8; ++j) z[i + j] = a[j];
    }
    for (; i < n; ++i) {
      const double* row = &data_x[(size_t)i * dd];
      double acc = 0.0;
      for (int c = 0; c < dd; ++c) acc = std::fma(row[c], x[c], acc);
      z[i] = acc;
    }
  }
  static void logreg_terms(double z, double y, double& ll, double& resid) {
    const double az = z < 0.0 ? -z : z;
    const double t = exp_(-az);
    const double sp = (z > 0.0 ? z : 0.0) + log1p_(t);         // softplus(z)
    const double sig = z >= 0.0 ? 1.0 / (1.0 + t) : t / (1.0 + t);
    ll = y * z - sp;
    resid = y - sig;
  }
  double logreg_lik(const double* x, double* grad_lik) const {
    const int n = n_data(), dd = d();
    std::vector<double> z, ll(n), rs(n);
    logreg_z(x, z);
    for (int i = 0; i < n; ++i) logreg_terms(z[i], data_y[i], ll[i], rs[i]);
    double total = 0.0;
    for (int t0 = 0; t0 < n; t0 += LR_TILE) {
      const int len = std::min(LR_TILE, n - t0);
      total = total + tree_sum(len, [&](int i) { return ll[t0 + i]; });
    }
    if (grad_lik) {
      // G[c] = sum over 4096-row chunks (in order) of the chunk's sequential fma chain over rows;
      // the loops are ordered row-major for the cache, each G[c] still sees its rows in ascending order
      std::vector<double> part(dd);
      for (int c = 0; c < dd; ++c) grad_lik[c] = 0.0;
      for (int k0 = 0; k0 < n; k0 += LR_CHUNK) {
        const int k1 = std::min(n, k0 + LR_CHUNK);
        std::fill(part.begin(), part.end(), 0.0);
        for (int i = k0; i < k1; ++i) {
          const double* row = &data_x[(size_t)i * dd];
          const double ri = rs[i];
          for (int c = 0; c < dd; ++c) part[c] = std::fma(row[c], ri, part[c]);
        }
        for (int c = 0; c < dd; ++c) grad_lik[c] = grad_lik[c] + part[c];
      }
    }
    return total;
  }

  double tgt_density(const Replica& r) const {
    const double* x = r.x.data();
    switch (cfg.target_kind) {
      case PGN_TARGET_LOGREG: return fixed_ref_density(r) + logreg_lik(x, nullptr);   // the prior is part of the posterior
      case PGN_TARGET_FUNNEL: {
        const double y = x[0];
        const double e = exp_(-y);
        const double sy = cfg.p[0], lsy = cfg.p[1];
        return tree_sum(d(), [&](int c) {
          if (c == 0) { double zy = y / sy; return -(zy * zy + LOG2PI) * 0.5 - lsy; }
          double t = x[c] * x[c] * e;
          return -(t + LOG2PI) * 0.5 - 0.5 * y;
        });
      }
      case PGN_TARGET_GMM: {
        const int K = cfg.n_modes;
        const double iv = cfg.p[2], cst = cfg.p[1];
        double a[64];
        double M = -INF;
        for (int k = 0; k < K; ++k) {
          const double* m = &means[(size_t)k * d()];
          double q = tree_sum(d(), [&](int c) { double t = x[c] - m[c]; return t * t; });
          a[k] = log_w[k] - 0.5 * q * iv - cst;
          if (a[k] > M) M = a[k];
        }
        double s = 0.0;
        for (int k = 0; k < K; ++k) s = s + exp_(a[k] - M);
        return M + log_(s);
      }
      case PGN_TARGET_ISING:
        return cfg.p[0] * (double)r.ising_S;   // examples/ising.jl:74
      case PGN_TARGET_MIXED: return mixed_density(x, 1);
      case PGN_TARGET_UNID: {   // unid_log_potential, test/test_DistributionLogPotential.jl:8-13
        if (!((x[0] >= 0.0 && x[0] <= 1.0) && (x[1] >= 0.0 && x[1] <= 1.0))) return -INF;
        const double pr = x[0] * x[1];
        return cfg.p[1] * log_(pr) + (cfg.p[0] - cfg.p[1]) * log1p_(-pr);
      }
      default: return QNAN;
    }
  }
  double ref_density_grad(const Replica& r, double* g) const {
    const double* x = r.x.data();
    if (uses_var(r.chain)) {   // GaussianReference.jl:72-80: -1/sd^2 (x - mean)
      for (int c = 0; c < d(); ++c) g[c] = -(var_t2[c] * (x[c] - var_mean[c]));
      return var_density(x);
    }
    const double iv = cfg.p[5];
    for (int c = 0; c < d(); ++c) g[c] = -x[c] * iv;
    return fixed_ref_density(r);
  }
  double tgt_density_grad(const Replica& r, double* g) const {
    const double* x = r.x.data();
    switch (cfg.target_kind) {
      case PGN_TARGET_LOGREG: {
        std::vector<double> gl(d());
        const double lik = logreg_lik(x, gl.data());
        const double iv = cfg.p[5];
        for (int c = 0; c < d(); ++c) g[c] = -x[c] * iv + gl[c];
        return fixed_ref_density(r) + lik;
      }
      case PGN_TARGET_FUNNEL: {
        const double y = x[0];
        const double e = exp_(-y);
        const double ivy = cfg.p[2];
        double T = tree_sum(d(), [&](int c) {
          if (c == 0) return 0.0;
          return 0.5 * (x[c] * x[c] * e) - 0.5;
        });
        g[0] = -y * ivy + T;
        for (int c = 1; c < d(); ++c) g[c] = -x[c] * e;
        return tgt_density(r);
      }
      case PGN_TARGET_GMM: {
        const int K = cfg.n_modes;
        const double iv = cfg.p[2], cst = cfg.p[1];
        double a[64];
        double M = -INF;
        for (int k = 0; k < K; ++k) {
          const double* m = &means[(size_t)k * d()];
          double q = tree_sum(d(), [&](int c) { double t = x[c] - m[c]; return t * t; });
          a[k] = log_w[k] - 0.5 * q * iv - cst;
          if (a[k] > M) M = a[k];
        }
        double w[64];
        double s = 0.0;
        for (int k = 0; k < K; ++k) { w[k] = exp_(a[k] - M); s = s + w[k]; }
        for (int c = 0; c < d(); ++c) {
          double acc = 0.0;
          for (int k = 0; k < K; ++k) acc = acc + w[k] * (means[(size_t)k * d() + c] - x[c]);
          g[c] = (acc / s) * iv;
        }
        return M + log_(s);
      }
      default: return QNAN;
    }
  }

  double toy_precision(double b) const {   // ScaledPrecisionNormalPath.jl:45-46
    return (1.0 - b) * cfg.p[0] + b * cfg.p[1];
  }
  double sqr_norm(const std::vector<double>& v) const {   // src/utils/misc.jl:10
    const double* p = v.data();
    return tree_sum((int)v.size(), [&](int c) { return p[c] * p[c]; });
  }

  // log_potential callable of chain with parameter b.
  //   toy MVN : ScaledPrecisionNormalPath.jl:19-20
  //   others  : InterpolatedLogPotential.jl:10-17 + InterpolatingPath.jl:26-27
  double log_potential(double b, Replica& r) {
    r.ref_equiv_evals += 1;
    if (cfg.target_kind == PGN_TARGET_TOY_MVN) return -0.5 * toy_precision(b) * sqr_norm(r.x);
    if (b == 0.0) return ref_density(r);
    if (b == 1.0) return tgt_density(r);
    return (1.0 - b) * ref_density(r) + b * tgt_density(r);
  }
  // LogDensityProblems.logdensity of the AD wrapper (BufferedAD.jl:89-94;
  // toy MVN: ScaledPrecisionNormalPath.jl:23)
  double logdensity(double b, Replica& r) {
    r.ref_equiv_evals += 1;
    if (cfg.target_kind == PGN_TARGET_TOY_MVN) return -0.5 * toy_precision(b) * sqr_norm(r.x);
    double l1 = ref_density(r);
    double l2 = tgt_density(r);
    return (1.0 - b) * l1 + b * l2;
  }
  // logdensity_and_gradient (BufferedAD.jl:98-111; toy: ScaledPrecisionNormalPath.jl:30-34)
  double logdensity_and_gradient(double b, Replica& r, double* g) {
    r.ref_equiv_evals += 1;
    const int dd = d();
    if (cfg.target_kind == PGN_TARGET_TOY_MVN) {
      double prec = toy_precision(b);
      double ld = -0.5 * prec * sqr_norm(r.x);
      for (int c = 0; c < dd; ++c) g[c] = -prec * r.x[c];
      return ld;
    }
    double logdens = 0.0;
    double l = ref_density_grad(r, r.g1.data());
    logdens = logdens + l * (1.0 - b);
    for (int c = 0; c < dd; ++c) g[c] = r.g1[c] * (1.0 - b);
    l = tgt_density_grad(r, r.g2.data());
    logdens = logdens + l * b;
    for (int c = 0; c < dd; ++c) g[c] = g[c] + r.g2[c] * b;
    return logdens;
  }

  // ------------------------------------------------------------------ init / iid
  void ising_recompute(Replica& r) const {   // examples/ising.jl:28-36 (each bond once)
    const int l = L();
    int s = 0;
    for (int i = 0; i < l; ++i)
      for (int j = 0; j < l; ++j) {
        int me = ((r.rows[i] >> j) & 1u) ? 1 : -1;
        s += me * ising_sum_neighbours(r, i, j);
      }
    r.ising_S = s / 2;
  }
  int ising_sum_neighbours(const Replica& r, int i, int j) const {   // examples/ising.jl:61-71
    const int l = L();
    auto sg = [&](int a, int b) { return ((r.rows[a] >> b) & 1u) ? 1 : -1; };
    int up = (i == 0) ? l - 1 : i - 1, dn = (i == l - 1) ? 0 : i + 1;
    int lf = (j == 0) ? l - 1 : j - 1, rt = (j == l - 1) ? 0 : j + 1;
    return sg(up, j) + sg(dn, j) + sg(i, lf) + sg(i, rt);
  }
  void ising_flip(Replica& r, int i, int j) const {   // examples/ising.jl:39-46
    int me = ((r.rows[i] >> j) & 1u) ? 1 : -1;
    int before = me * ising_sum_neighbours(r, i, j);
    r.rows[i] ^= (1u << j);
    int after = -me * ising_sum_neighbours(r, i, j);
    r.ising_S += after - before;
  }
  void ising_sync_x(Replica& r) const {
    const int l = L();
    for (int i = 0; i < l; ++i)
      for (int j = 0; j < l; ++j) r.x[(size_t)i * l + j] = ((r.rows[i] >> j) & 1u) ? 1.0 : 0.0;
  }

  // initialization(target, rng, replica_index)
  void initialization(Replica& r) {
    const int dd = d();
    r.x.assign(dd, 0.0);
    switch (cfg.target_kind) {
      case PGN_TARGET_TOY_MVN: {   // toy_mvn_target.jl:10-11
        double sq = std::sqrt(cfg.p[1]);
        for (int c = 0; c < dd; ++c) r.x[c] = normal_at(r.rng, r.ctr + c) / sq;
        r.ctr += dd;
        break;
      }
      case PGN_TARGET_ISING:       // examples/ising.jl:85 (all false)
        r.rows.assign(L(), 0u);
        ising_recompute(r);
        break;
      case PGN_TARGET_UNID:        // test/test_DistributionLogPotential.jl:14
        r.x.assign(dd, 0.5);
        break;
      default: break;              // zeros (dimensional-analysis.jl:24)
    }
  }
  // sample_iid!(reference_log_potential, replica, shared)
  void sample_iid(double b, Replica& r) {
    const int dd = d();
    if (uses_var(r.chain)) {   // GaussianReference.jl:30-37: randn * sd + mean, coordinate by coordinate
      for (int c = 0; c < dd; ++c) r.x[c] = normal_at(r.rng, r.ctr + c) * var_sd[c] + var_mean[c];
      r.ctr += dd;
      return;
    }
    switch (cfg.target_kind) {
      case PGN_TARGET_TOY_MVN: {   // toy_mvn_target.jl:15-21
        double sq = std::sqrt(toy_precision(b));
        for (int c = 0; c < dd; ++c) r.x[c] = normal_at(r.rng, r.ctr + c) / sq;
        r.ctr += dd;
        break;
      }
      case PGN_TARGET_FUNNEL:
      case PGN_TARGET_GMM:
      case PGN_TARGET_LOGREG: {    // DistributionLogPotential.jl:26-27 (rand!(rng, MvNormal(0, s^2 I), x))
        double sg = cfg.p[3];
        for (int c = 0; c < dd; ++c) r.x[c] = sg * normal_at(r.rng, r.ctr + c);
        r.ctr += dd;
        break;
      }
      case PGN_TARGET_ISING: {     // examples/ising.jl:49-58 (one 32-bit word per row)
        const int l = L();
        uint32_t mask = (l >= 32) ? 0xffffffffu : ((1u << l) - 1u);
        for (int i = 0; i < l; ++i) r.rows[i] = bits32_at(r.rng, r.ctr + i) & mask;
        r.ctr += l;
        ising_recompute(r);
        break;
      }
      case PGN_TARGET_UNID:        // rand!(rng, product_distribution([Uniform(), Uniform()]), x)
        for (int c = 0; c < dd; ++c) r.x[c] = uniform_at(r.rng, r.ctr + c);
        r.ctr += dd;
        break;
      case PGN_TARGET_MIXED: {     // rand! of the product reference: one tick per Bool / Float coordinate, n per Binomial
        const int n = (int)cfg.p[2];
        const double p0 = means[8], q0 = means[9];
        uint64_t t = r.ctr;
        for (int c = 0; c < dd; ++c) {
          switch (mixed_kind(c)) {
            case 0: r.x[c] = uniform_at(r.rng, t) < p0 ? 1.0 : 0.0; t += 1; break;
            case 1: {
              int k = 0;
              for (int j = 0; j < n; ++j) k += uniform_at(r.rng, t + j) < q0 ? 1 : 0;
              r.x[c] = (double)k; t += n; break;
            }
            default: r.x[c] = cfg.p[3] * normal_at(r.rng, t); t += 1; break;
          }
        }
        r.ctr = t;
        break;
      }
      default: break;              // TestSwapper: nothing (pair_swapper.jl:143)
    }
  }

  // ------------------------------------------------------------------ SliceSampler
  // src/explorers/SliceSampler.jl:24-237
  void slice_step(Replica& r, ChainStats& st) {
    const double b = beta[r.chain - 1];
    double cached_lp = -INF;
    for (int pass = 0; pass < ep.slice_n_passes; ++pass) cached_lp = slice_sample(r, st, b, cached_lp);
  }
  double slice_sample(Replica& r, ChainStats& st, double b, double cached_lp) {
    // cached_log_potential :32-41
    if (cached_lp == -INF) {
      double result = log_potential(b, r);
      if (result == -INF) throw OrcError{PGN_ERR_BAD_DENSITY, "SliceSampler must be initialized in the support"};
      cached_lp = result;
    }
    for (int c = 0; c < d(); ++c) {
      cached_lp = slice_sample_coord(r, st, b, c, cached_lp);
      if (!std::isfinite(cached_lp))
        throw OrcError{PGN_ERR_BAD_DENSITY, "invalid log density after updating state at index " + std::to_string(c)};
    }
    return cached_lp;
  }
  // coordinate types (SliceSampler.jl dispatches on typeof(pointer[])): only the MIXED target has non-Float coordinates
  int coord_kind(int c) const { return cfg.target_kind == PGN_TARGET_MIXED ? mixed_kind(c) : 2; }
  // rand(rng, a:b) for integers a <= b held in doubles: a + floor(u (b - a + 1)), u the replica's next uniform
  // (Julia's range sampler is not reproducible here; part of the documented RNG deviation)
  static double rand_int_range(Replica& r, double a, double b) {
    const double span = (b - a) + 1.0;
    double k = std::floor(r.uniform() * span);
    if (k > b - a) k = b - a;
    return a + k;
  }
  // Bool coordinates: sample from the full conditional, one density evaluation (SliceSampler.jl:65-86)
  double slice_sample_bool(Replica& r, double b, int c, double cached_lp) {
    double& ptr = r.x[c];
    double lp0, lp1;
    if (ptr != 0.0) { lp1 = cached_lp; ptr = 0.0; lp0 = log_potential(b, r); }
    else { lp0 = cached_lp; ptr = 1.0; lp1 = log_potential(b, r); }
    const double prob_ratio = exp_(lp1 - lp0);
    const double prob_zero = 1.0 / (1.0 + prob_ratio);
    if (r.uniform() < prob_zero) { ptr = 0.0; return lp0; }
    ptr = 1.0;
    return lp1;
  }
  double slice_sample_coord(Replica& r, ChainStats& st, double b, int c, double cached_lp) {   // :89-95
    if (coord_kind(c) == 0) return slice_sample_bool(r, b, c, cached_lp);
    double z = cached_lp - r.exponential();
    double Lq, Rq, lp_L, lp_R;
    slice_double(r, st, b, c, z, Lq, Rq, lp_L, lp_R);
    return slice_shrink(r, st, b, c, z, Lq, Rq, lp_L, lp_R);
  }
  void slice_double(Replica& r, ChainStats& st, double b, int c, double z, double& Lq, double& Rq,
                    double& potent_L, double& potent_R) {   // :97-126
    double& ptr = r.x[c];
    double old_position = ptr;
    if (coord_kind(c) == 1) {              // initialize_slice_endpoints for integers :136-142
      if (ep.slice_w != std::floor(ep.slice_w))
        throw OrcError{PGN_ERR_INVALID, "for integer variables, the width should be an integer"};
      const double width = std::ceil(ep.slice_w);
      Lq = ptr - rand_int_range(r, 0.0, width);
      Rq = Lq + width;
    } else {
      Lq = ptr - ep.slice_w * r.uniform();   // initialize_slice_endpoints :129-133
      Rq = Lq + ep.slice_w;
    }
    int K = ep.slice_p;
    ptr = Lq; potent_L = log_potential(b, r);
    ptr = Rq; potent_R = log_potential(b, r);
    while (K > 0 && ((z < potent_L) || (z < potent_R))) {
      double V = r.uniform();
      if (V <= 0.5) {
        Lq = Lq - (Rq - Lq);
        ptr = Lq; potent_L = log_potential(b, r);
      } else {
        Rq = Rq + (Rq - Lq);
        ptr = Rq; potent_R = log_potential(b, r);
      }
      K -= 1;
    }
    st.n_steps += ep.slice_p - K;
    ptr = old_position;
  }
  static bool isapprox(double x, double y) {   // Base.isapprox, rtol = sqrt(eps) = 2^-26, atol = 0
    const double rtol = from_bits(0x3e50000000000000ULL);
    if (x == y) return true;
    if (!(std::isfinite(x) && std::isfinite(y))) return false;
    double ax = std::fabs(x), ay = std::fabs(y);
    return std::fabs(x - y) <= rtol * (ax > ay ? ax : ay);
  }
  double slice_shrink(Replica& r, ChainStats& st, double b, int c, double z, double Lq, double Rq,
                      double lp_L, double lp_R) {   // :144-186
    double& ptr = r.x[c];
    double old_position = ptr;
    double Lbar = Lq, Rbar = Rq;
    double new_lp = 0.0;
    int n = 1;
    while (n <= ep.slice_max_iter) {
      const bool integer = coord_kind(c) == 1;
      double new_position = integer ? rand_int_range(r, Lbar, Rbar)          // draw_new_position(L::Integer, R::Integer) :189
                                    : Lbar + r.uniform() * (Rbar - Lbar);    // draw_new_position :188
      ptr = new_position;
      new_lp = log_potential(b, r);
      bool consider = z < new_lp;
      ptr = old_position;
      if (consider && slice_accept(r, st, b, c, new_position, z, Lq, Rq, lp_L, lp_R)) {
        ptr = new_position;
        st.n_steps += n;
        return new_lp;
      }
      if (new_position < ptr) Lbar = new_position; else Rbar = new_position;
      if (integer ? (Lbar == Rbar) : isapprox(Lbar, Rbar)) {   // isapprox of two Integers is ==
        ptr = old_position;
        st.n_steps += n;
        return log_potential(b, r);
      }
      n += 1;
    }
    throw OrcError{PGN_ERR_SLICE_MAX_ITER, "slice_shrink: maximum number of iterations reached"};
  }
  bool slice_accept(Replica& r, ChainStats& st, double b, int c, double new_position, double z, double Lq,
                    double Rq, double lp_L, double lp_R) {   // :192-237
    double& ptr = r.x[c];
    double old_position = ptr;
    double Lhat = Lq, Rhat = Rq;
    bool Rstale = false, Lstale = false;
    bool D = false;
    while (Rhat - Lhat > 1.1 * ep.slice_w) {
      double M = (Lhat + Rhat) / 2.0;
      if (((old_position < M) && (new_position >= M)) || ((old_position >= M) && (new_position < M))) D = true;
      if (new_position < M) { Rhat = M; Rstale = true; } else { Lhat = M; Lstale = true; }
      if (D) {
        if (Lstale) { ptr = Lhat; lp_L = log_potential(b, r); Lstale = false; }
        if (Rstale) { ptr = Rhat; lp_R = log_potential(b, r); Rstale = false; }
        if ((z >= lp_L) && (z >= lp_R)) {
          ptr = old_position;
          st.expl_acc.fit(0.0);
          return false;
        }
      }
    }
    ptr = old_position;
    st.expl_acc.fit(1.0);
    return true;
  }

  // ------------------------------------------------------------------ AutoMALA
  // src/explorers/hamiltonian_dynamics.jl:28-29
  double log_joint(double logp, const std::vector<double>& momentum) const { return logp - 0.5 * sqr_norm(momentum); }
  // hamiltonian_dynamics!(…, n_steps = 1)  (hamiltonian_dynamics.jl:39-84)
  bool leap_frog(double b, Replica& r, double step_size) {
    const int dd = d();
    double* grad = r.grad.data();
    // conditioned_target_gradient :31-35
    double logp = logdensity_and_gradient(b, r, grad);
    for (int c = 0; c < dd; ++c) grad[c] = grad[c] / r.precond[c];
    for (int c = 0; c < dd; ++c) r.momentum[c] = r.momentum[c] + (step_size / 2) * grad[c];
    // full position step
    for (int c = 0; c < dd; ++c) r.x[c] = r.x[c] + step_size * (r.momentum[c] / r.precond[c]);
    logp = logdensity_and_gradient(b, r, grad);
    for (int c = 0; c < dd; ++c) grad[c] = grad[c] / r.precond[c];
    double current_log_joint = log_joint(logp, r.momentum);
    if (!std::isfinite(current_log_joint)) return false;
    for (int c = 0; c < dd; ++c) r.momentum[c] = r.momentum[c] + (step_size / 2) * grad[c];
    if (!std::isfinite(sqr_norm(r.momentum))) return false;
    return true;
  }
  // log_joint_difference_function :250-275 — returns h_before, caller evaluates trials
  struct LJD { double h_before; };
  LJD ljd_begin(double b, Replica& r) {
    r.state_before = r.x;
    r.momentum_before = r.momentum;
    return LJD{log_joint(logdensity(b, r), r.momentum)};
  }
  double ljd_eval(double b, Replica& r, const LJD& f, double step_size) {
    leap_frog(b, r, step_size);
    double h_after = log_joint(logdensity(b, r), r.momentum);
    r.x = r.state_before;
    r.momentum = r.momentum_before;
    return h_after - f.h_before;
  }
  // auto_step_size :184-214 (+ grow :216-226, shrink :228-248)
  int auto_step_size(double b, Replica& r, ChainStats& st, double step_size, double lower_bound, double upper_bound) {
    if (!(step_size > 0)) throw OrcError{PGN_ERR_INVALID, "autoMALA: step_size must be > 0"};
    if (!(lower_bound < upper_bound)) throw OrcError{PGN_ERR_INVALID, "autoMALA: lower_bound < upper_bound violated"};
    LJD f = ljd_begin(b, r);
    double initial_difference = ljd_eval(b, r, f, step_size);
    int n_steps = 0, exponent = 0;
    if (!std::isfinite(initial_difference) || initial_difference < lower_bound) {
      int n = 1;
      double eps = step_size;
      while (true) {
        eps = eps / 2.0;
        double diff = ljd_eval(b, r, f, eps);
        if (eps == 0.0) throw OrcError{PGN_ERR_STEP_UNDERFLOW, "autoMALA: could not find a positive step size"};
        if (diff > lower_bound) { n_steps = n; exponent = -n; break; }
        n += 1;
      }
    } else if (initial_difference > upper_bound) {
      int n = 1;
      double eps = step_size;
      while (true) {
        eps = eps * 2.0;
        double diff = ljd_eval(b, r, f, eps);
        if (!std::isfinite(diff) || diff < upper_bound) { n_steps = n; exponent = n - 1; break; }
        n += 1;
      }
    }
    st.n_steps += 1 + n_steps;
    st.am.fit(pow2(exponent));
    return exponent;
  }
  static double pow2(int e) {   // 2.0^e, exact
    if (e > 1023) return INF;
    if (e < -1074) return 0.0;
    if (e >= -1022) return pow2i(e);
    return scale2(1.0, e);
  }
  // One explorer of a Mix (src/explorers/Mix.jl:20-21: step!(rand(replica.rng, explorer.explorers), ...)): the
  // variant is drawn uniformly with one tick of the replica's stream; a plain autoMALA / MALA is the single variant.
  struct Variant { int n_refresh; double step_size; int precond_kind; double p0, p01; };
  Variant pick_variant(Replica& r) {
    if (ep.n_mix <= 1) return Variant{ep.n_refresh, ep.step_size, ep.precond_kind, ep.mix_p0, ep.mix_p01};
    int v = (int)(r.uniform() * (double)ep.n_mix);
    if (v >= ep.n_mix) v = ep.n_mix - 1;
    return Variant{ep.mix_n_refresh[v], ep.mix_step_size[v], ep.mix_precond_kind[v], ep.mix_variant_p0[v], ep.mix_variant_p01[v]};
  }
  // build_preconditioner! (Preconditioner.jl:57-77)
  void build_preconditioner(Replica& r, const Variant& vr) {
    const int dd = d();
    if (!have_std || vr.precond_kind == PGN_PRECOND_IDENTITY) {
      for (int c = 0; c < dd; ++c) r.precond[c] = 1.0;
      return;
    }
    if (vr.precond_kind == PGN_PRECOND_DIAGONAL) {
      for (int c = 0; c < dd; ++c) r.precond[c] = std_devs[c] == 0.0 ? 1.0 : 1.0 / std_devs[c];
      return;
    }
    double u = r.uniform();
    if (u <= vr.p0) {
      for (int c = 0; c < dd; ++c) r.precond[c] = std_devs[c] == 0.0 ? 1.0 : 1.0 / std_devs[c];
    } else if (u <= vr.p01) {
      for (int c = 0; c < dd; ++c) r.precond[c] = 1.0;
    } else {
      double mix = r.uniform();
      double rmix = 1.0 - mix;
      for (int c = 0; c < dd; ++c) r.precond[c] = std_devs[c] == 0.0 ? 1.0 : mix + rmix / std_devs[c];
    }
  }
  // auto_mala! (AutoMALA.jl:106-182)
  void auto_mala(Replica& r, ChainStats& st, bool use_mh) { auto_mala(r, st, use_mh, pick_variant(r)); }
  void auto_mala(Replica& r, ChainStats& st, bool use_mh, const Variant& vr) {
    const int dd = d();
    const double b = beta[r.chain - 1];
    build_preconditioner(r, vr);
    for (int i = 0; i < vr.n_refresh; ++i) {
      r.start_state = r.x;
      for (int c = 0; c < dd; ++c) r.momentum[c] = normal_at(r.rng, r.ctr + c);   // randn!(rng, momentum)
      r.ctr += dd;
      double init_joint_log = log_joint(logdensity(b, r), r.momentum);
      if (!std::isfinite(init_joint_log))
        throw OrcError{PGN_ERR_NOT_POSITIVE, "AutoMALA can only be called on a configuration of positive density"};
      double a = r.uniform();
      double bb = r.uniform();
      double lower_bound = log_(a < bb ? a : bb);
      double upper_bound = log_(a < bb ? bb : a);
      int proposed_exponent = auto_step_size(b, r, st, vr.step_size, lower_bound, upper_bound);
      double proposed_step_size = vr.step_size * pow2(proposed_exponent);
      leap_frog(b, r, proposed_step_size);
      if (use_mh) {
        for (int c = 0; c < dd; ++c) r.momentum[c] = r.momentum[c] * -1.0;
        int reversed_exponent = auto_step_size(b, r, st, vr.step_size, lower_bound, upper_bound);
        bool reversibility_passed = reversed_exponent == proposed_exponent;
        st.rev.fit(reversibility_passed ? 1.0 : 0.0);
        double probability;
        if (reversibility_passed) {
          double final_joint_log = log_joint(logdensity(b, r), r.momentum);
          double e = exp_(final_joint_log - init_joint_log);
          probability = 1.0 < e ? 1.0 : e;   // min(1.0, e) (NaN propagates like Julia's min)
          if (e != e) probability = e;
        } else {
          probability = 0.0;
        }
        st.expl_acc.fit(probability);
        if (r.uniform() < probability) {
          // accept
        } else {
          r.x = r.start_state;
        }
      }
    }
  }

  // mala! (src/explorers/MALA.jl:74-97)
  void mala(Replica& r, ChainStats& st) { mala(r, st, Variant{ep.n_refresh, ep.step_size, ep.precond_kind, ep.mix_p0, ep.mix_p01}); }
  void mala(Replica& r, ChainStats& st, const Variant& vr) {
    const int dd = d();
    const double b = beta[r.chain - 1];
    build_preconditioner(r, vr);
    for (int i = 0; i < vr.n_refresh; ++i) {
      r.start_state = r.x;
      for (int c = 0; c < dd; ++c) r.momentum[c] = normal_at(r.rng, r.ctr + c);
      r.ctr += dd;
      double init_joint_log = log_joint(logdensity(b, r), r.momentum);
      if (!std::isfinite(init_joint_log))
        throw OrcError{PGN_ERR_NOT_POSITIVE, "MALA can only be called on a configuration of positive density"};
      leap_frog(b, r, vr.step_size);
      for (int c = 0; c < dd; ++c) r.momentum[c] = r.momentum[c] * -1.0;
      double final_joint_log = log_joint(logdensity(b, r), r.momentum);
      double e = exp_(final_joint_log - init_joint_log);
      double probability = 1.0 < e ? 1.0 : e;
      st.expl_acc.fit(probability);
      if (r.uniform() < probability) {
      } else {
        r.x = r.start_state;
      }
      st.n_steps += 1;
    }
  }

  // ------------------------------------------------------------------ IsingMetropolis
  // examples/ising.jl:98-117
  void ising_metropolis(Replica& r) {
    const double b = beta[r.chain - 1];
    const int l = L();
    for (int k = 0; k < ep.ising_n_steps; ++k)
      for (int i = 0; i < l; ++i)
        for (int j = 0; j < l; ++j) {
          double log_pr_before = log_potential(b, r);
          ising_flip(r, i, j);
          double log_pr_after = log_potential(b, r);
          double accept_ratio = exp_(log_pr_after - log_pr_before);
          if (accept_ratio < 1 && r.uniform() > accept_ratio) ising_flip(r, i, j);
        }
  }

  // one explorer of a Compose / Mix program
  void program_step(Replica& r, ChainStats& st, int s) {
    const Variant vr{ep.mix_n_refresh[s], ep.mix_step_size[s], ep.mix_precond_kind[s], ep.mix_variant_p0[s], ep.mix_variant_p01[s]};
    switch (ep.step_kind[s]) {
      case PGN_EXPLORER_TOY: sample_iid(beta[r.chain - 1], r); break;
      case PGN_EXPLORER_SLICE: slice_step(r, st); break;
      case PGN_EXPLORER_AUTOMALA: auto_mala(r, st, scan != 1, vr); break;
      case PGN_EXPLORER_MALA: mala(r, st, vr); break;
      default: throw OrcError{PGN_ERR_INVALID, "unsupported explorer inside Compose / Mix"};
    }
  }

  // ------------------------------------------------------------------ explore!
  // explore!(pt, replica, explorer)  (src/pt/pigeons.jl:101-132)
  void explore(Replica& r) {
    ChainStats& st = st_of(r);
    const bool is_reference = this->is_reference(r.chain);
    if (cfg.target_kind == PGN_TARGET_TEST_SWAPPER) return;
    if (is_reference) {
      sample_iid(beta[r.chain - 1], r);
    } else {
      switch (ep.kind) {
        case PGN_EXPLORER_TOY: sample_iid(beta[r.chain - 1], r); break;   // ToyExplorer.jl:7-12
        case PGN_EXPLORER_SLICE: slice_step(r, st); break;
        case PGN_EXPLORER_AUTOMALA: auto_mala(r, st, scan != 1); break;   // AutoMALA.jl:87,102
        case PGN_EXPLORER_COMPOSE:                                        // Compose.jl:16-19: every explorer in turn
          for (int s = 0; s < ep.n_steps; ++s) program_step(r, st, s);
          break;
        case PGN_EXPLORER_MIX: {                                          // Mix.jl:20-21: rand(rng, explorers) performs the step
          int v = (int)(r.uniform() * (double)ep.n_steps);
          if (v >= ep.n_steps) v = ep.n_steps - 1;
          program_step(r, st, v);
          break;
        }
        case PGN_EXPLORER_MALA: mala(r, st); break;
        case PGN_EXPLORER_ISING_METROPOLIS: ising_metropolis(r); break;
        default: break;
      }
    }
  }

  // ------------------------------------------------------------------ swap!
  int partner_chain(int chain) const {   // OddEven.jl:23-31, DEO.jl:12
    bool even = (scan % 2 == 0);
    int direction = ((chain % 2 == 0) == even) ? 1 : -1;
    int proposed = chain + direction;
    if (proposed == 0) return 1;
    if (proposed == N() + 1) return N();
    return proposed;
  }
  // swap_stat (pair_swapper.jl:42-47) + log_unnormalized_ratio (log_potentials.jl:43-51)
  void swap_stat(Replica& r, int partner) {
    if (cfg.target_kind == PGN_TARGET_TEST_SWAPPER) {   // pair_swapper.jl:114
      r.lr = 0.0;
      r.u = r.uniform();
      return;
    }
    double lp_num = log_potential(beta[partner - 1], r);
    double lp_den = log_potential(beta[r.chain - 1], r);
    double ans = lp_num - lp_den;
    if (ans != ans) throw OrcError{PGN_ERR_NAN_RATIO, "Got NaN log-unnormalized ratio"};
    r.lr = ans;
    r.u = r.uniform();
  }
  void record_round_trip(Replica& r) {   // RoundTripRecorder.jl:43-54
    bool is_ref = is_reference(r.chain), is_tgt = is_target(r.chain);
    if (r.rt_state == 0 && is_ref) r.rt_state = 1;
    else if (r.rt_state == 1 && is_tgt) { r.rt_state = 2; n_restarts += 1; }
    else if (r.rt_state == 2 && is_ref) { r.rt_state = 1; n_round_trips += 1; }
  }
};

struct RoundLogs {
  int32_t* index_process;
  double* swap_lr;
  double* swap_u;
  uint8_t* swap_accept;
  double* target_trace;
};

void run_round(Engine& E, int64_t n_scans, pgn_round_out* out) {
  const int N = E.N(), d = E.d();
  E.stats.assign(N, ChainStats{});
  E.rec.assign(E.per_replica() ? N : 0, ReplicaRecorders{});
  E.n_restarts = E.n_round_trips = 0;
  E.online.assign(d, VarAcc{});
  for (auto& r : E.replicas) { r.rt_state = 0; r.ref_equiv_evals = 0; }   // recorders are emptied every round (recorders.jl:113-118)
  int64_t total_ref_evals = 0;

  for (int64_t s = 1; s <= n_scans; ++s) {
    E.scan = (int)s;
    // ---- explore!(pt, explorer, multithreaded)  (pigeons.jl:82-97)
    OrcError first_err{0, ""};
#ifdef _OPENMP
#pragma omp parallel for schedule(static) num_threads(E.n_threads)
#endif
    for (int i = 0; i < N; ++i) {
      try {
        E.explore(E.replicas[i]);
      } catch (OrcError& e) {
#ifdef _OPENMP
#pragma omp critical
#endif
        { if (first_err.code == 0) first_err = e; }
      }
    }
    if (first_err.code) throw first_err;
    // target-chain recording (pigeons.jl:110-131): every replica at a target chain, into its own recorders
    {
      int t_idx = 0;
      for (int chain = 1; chain <= N; ++chain) {
        if (!E.is_target(chain)) continue;
        Replica& rt = E.replicas[chain - 1];
        if (E.cfg.target_kind == PGN_TARGET_ISING) E.ising_sync_x(rt);
        // online statistics are kept for vector states only (IsingState is not a
        // continuous-variable state; OnlineStateRecorder.jl:87-110 does not apply)
        if (E.cfg.target_kind != PGN_TARGET_ISING) {
          std::vector<VarAcc>* on = &E.online;
          if (E.per_replica()) {
            on = &E.rec[rt.replica_index - 1].online;
            if (on->empty()) on->assign(d, VarAcc{});
          }
          for (int c = 0; c < d; ++c) (*on)[c].fit(rt.x[c]);
        }
        const int n_tgt = E.two_legs() ? 2 : 1;
        if (out->target_trace) std::memcpy(out->target_trace + ((size_t)(s - 1) * n_tgt + t_idx) * d, rt.x.data(), sizeof(double) * d);
        t_idx += 1;
      }
    }
    // ---- swap!  (swap.jl:6-26)
    for (int my_chain = 1; my_chain <= N; ++my_chain) {
      Replica& mine = E.replicas[my_chain - 1];
      int partner = E.partner_chain(my_chain);
      if (partner >= my_chain) {
        Replica& other = E.replicas[partner - 1];
        E.swap_stat(mine, partner);
        if (partner != my_chain) E.swap_stat(other, my_chain);
        // _swap! both halves (swap.jl:106-126): recorders first
        Replica* both[2] = {&mine, partner != my_chain ? &other : nullptr};
        for (Replica* r : both) {
          if (!r) continue;
          if (out->index_process) out->index_process[(size_t)(s - 1) * N + (r->chain - 1)] = r->replica_index;
          if (out->swap_lr) out->swap_lr[(size_t)(s - 1) * N + (r->chain - 1)] = r->lr;
          if (out->swap_u) out->swap_u[(size_t)(s - 1) * N + (r->chain - 1)] = r->u;
          E.record_round_trip(*r);
        }
        bool do_swap = false;
        if (partner != my_chain) {
          double acceptance_pr;
          if (E.cfg.target_kind == PGN_TARGET_TEST_SWAPPER) {
            acceptance_pr = E.cfg.p[0];           // pair_swapper.jl:121-124 (no stats recorded :131)
          } else {
            double e = exp_(mine.lr + other.lr);  // swap_acceptance_probability :88
            acceptance_pr = 1.0 < e ? 1.0 : e;
            ChainStats& st = E.st_of(mine);           // record_swap_stats! :59-66, by the replica holding the lower chain (swap.jl:119-121)
            st.swap_acc.fit(acceptance_pr);
            st.ls_fwd.fit(mine.lr);
            st.ls_bwd.fit(other.lr);
          }
          do_swap = mine.u < acceptance_pr;       // swap_decision :81-85 (uniform of the lower chain)
        }
        if (out->swap_accept) {
          out->swap_accept[(size_t)(s - 1) * N + (my_chain - 1)] = do_swap ? 1 : 0;
          if (partner != my_chain) out->swap_accept[(size_t)(s - 1) * N + (partner - 1)] = do_swap ? 1 : 0;
        }
        if (do_swap) {
          mine.chain = partner;
          other.chain = my_chain;
          std::swap(E.replicas[my_chain - 1], E.replicas[partner - 1]);   // resort_replicas! (swap.jl:28-39)
        }
      }
    }
  }
  for (auto& r : E.replicas) total_ref_evals += r.ref_equiv_evals;

  // ---- reduce_recorders! (recorders.jl:88-120): replicas sorted by replica_index, then the binary tree of
  // reduce_deterministically (Entangler.jl:214-277): at the level with spacing s, entry i absorbs entry i + s for
  // i = 1, 1 + 2s, ...; an unpaired last entry waits for the next level.
  if (E.per_replica()) {
    std::vector<ReplicaRecorders>& work = E.rec;
    for (int spacing = 1; spacing < N; spacing *= 2)
      for (int i = 0; i + spacing < N; i += 2 * spacing) merge_recorders(work[i], work[i + spacing]);
    for (const auto& kv : work[0].by_chain) E.stats[kv.first] = kv.second;
    if (!work[0].online.empty()) E.online = work[0].online;
  }

  // ---- outputs
  for (int c = 0; c < N; ++c) {
    const ChainStats& st = E.stats[c];
    if (out->swap_n) out->swap_n[c] = st.swap_acc.n;
    if (out->swap_mean) out->swap_mean[c] = st.swap_acc.mu;
    if (out->logsum_fwd) out->logsum_fwd[c] = st.ls_fwd.value;
    if (out->logsum_bwd) out->logsum_bwd[c] = st.ls_bwd.value;
    if (out->expl_acc_n) out->expl_acc_n[c] = st.expl_acc.n;
    if (out->expl_acc_mean) out->expl_acc_mean[c] = st.expl_acc.mu;
    if (out->expl_n_steps) out->expl_n_steps[c] = st.n_steps;
    if (out->am_n) out->am_n[c] = st.am.n;
    if (out->am_mean) out->am_mean[c] = st.am.mu;
    if (out->rev_n) out->rev_n[c] = st.rev.n;
    if (out->rev_mean) out->rev_mean[c] = st.rev.mu;
  }
  out->n_tempered_restarts = E.n_restarts;
  out->n_round_trips = E.n_round_trips;
  const bool vec_state = E.cfg.target_kind != PGN_TARGET_ISING && d > 0;
  out->online_n = vec_state ? E.online[0].n : n_scans;
  for (int c = 0; c < d && vec_state; ++c) {
    if (out->online_mean) out->online_mean[c] = E.online[c].mu;
    if (out->online_var) out->online_var[c] = E.online[c].value();
  }
  out->n_ref_equiv_evals = total_ref_evals;
  out->n_density_points = 0;
  out->kernel_ms = 0.0;
  out->gemm_ms = 0.0;
  out->batch_steps = 0;
  out->n_launches = 0;
  out->active_columns = 0;
  out->gemm_columns = 0;
}

int fail(char** err, int code, const std::string& msg) {
  if (err) {
    *err = (char*)std::malloc(msg.size() + 1);
    std::memcpy(*err, msg.c_str(), msg.size() + 1);
  }
  return code;
}

}  // namespace

// ===========================================================================
// C API: same shapes as include/pigeons_b200.h with the prefix orc_, so the
// tests can drive the oracle and the CUDA engine through one adapter.
// ===========================================================================
extern "C" {

struct orc_handle { Engine E; };

int orc_create(const pgn_config* cfg, orc_handle** out, char** err) {
  if (!cfg || !out) return fail(err, PGN_ERR_INVALID, "null argument");
  if (cfg->abi_version != PGN_ABI_VERSION) return fail(err, PGN_ERR_INVALID, "ABI version mismatch");
  if (cfg->world_size != 1 || cfg->rank != 0) return fail(err, PGN_ERR_INVALID, "oracle is single-process");
  if (cfg->n_chains < 1) return fail(err, PGN_ERR_INVALID, "n_chains must be >= 1");
  if (cfg->n_chains_variational < 0 || cfg->n_chains_variational > cfg->n_chains)
    return fail(err, PGN_ERR_INVALID, "0 <= n_chains_variational <= n_chains");
  if (cfg->n_chains_variational > 0 && cfg->n_chains_variational < cfg->n_chains &&
      cfg->recorder_order != PGN_RECORDERS_PER_REPLICA)
    return fail(err, PGN_ERR_INVALID, "two legs need recorder_order = PGN_RECORDERS_PER_REPLICA");
  if (cfg->target_kind == PGN_TARGET_LOGREG && (!cfg->data_x || !cfg->data_y || cfg->p[0] < 1))
    return fail(err, PGN_ERR_INVALID, "LOGREG: data_x / data_y / n_data missing");
  if (cfg->target_kind == PGN_TARGET_GMM && (cfg->n_modes < 1 || cfg->n_modes > 64))
    return fail(err, PGN_ERR_INVALID, "GMM: 1 <= n_modes <= 64");
  if (cfg->target_kind == PGN_TARGET_ISING) {
    int L = (int)cfg->p[1];
    if (L < 2 || L > 32 || L * L != cfg->dim) return fail(err, PGN_ERR_INVALID, "ISING: 2 <= L <= 32 and dim == L*L");
  }
  auto h = new orc_handle();
  Engine& E = h->E;
  E.cfg = *cfg;
  if (cfg->target_kind == PGN_TARGET_GMM) {
    E.means.assign(cfg->means, cfg->means + (size_t)cfg->n_modes * cfg->dim);
    E.log_w.assign(cfg->log_weights, cfg->log_weights + cfg->n_modes);
  }
  if (cfg->target_kind == PGN_TARGET_UNID && (cfg->dim != 2 || !(cfg->p[0] >= cfg->p[1]) || cfg->p[1] < 0)) {
    delete h;
    return fail(err, PGN_ERR_INVALID, "UNID: dim == 2 and 0 <= n_successes <= n_trials");
  }
  if (cfg->target_kind == PGN_TARGET_MIXED) {
    if (!cfg->means || cfg->n_modes != 10 + (int)cfg->p[2] + 1 || cfg->p[0] < 0 || cfg->p[1] < 0 ||
        cfg->p[0] + cfg->p[1] > cfg->dim || cfg->p[2] < 1) {
      delete h;
      return fail(err, PGN_ERR_INVALID, "MIXED: parameter table / coordinate counts inconsistent");
    }
    E.means.assign(cfg->means, cfg->means + cfg->n_modes);
  }
  if (cfg->target_kind == PGN_TARGET_LOGREG) {
    const size_t n = (size_t)cfg->p[0];
    E.data_x.assign(cfg->data_x, cfg->data_x + n * cfg->dim);
    E.data_y.assign(cfg->data_y, cfg->data_y + n);
  }
  E.cfg.means = E.cfg.log_weights = E.cfg.data_x = E.cfg.data_y = nullptr;
  E.beta.assign(cfg->n_chains, 0.0);
  for (int i = 0; i < cfg->n_chains; ++i) E.beta[i] = cfg->n_chains == 1 ? 1.0 : (double)i / (cfg->n_chains - 1);
  E.ep = pgn_explorer_params{};
  E.ep.kind = PGN_EXPLORER_NONE;
#ifdef _OPENMP
  E.n_threads = omp_get_max_threads();
#endif
  *out = h;
  return PGN_OK;
}
int orc_destroy(orc_handle* h) { delete h; return PGN_OK; }
void orc_free_string(char* s) { std::free(s); }
int orc_set_threads(orc_handle* h, int n) { h->E.n_threads = n < 1 ? 1 : n; return PGN_OK; }
int orc_get_threads(orc_handle* h) { return h->E.n_threads; }

int orc_local_range(const orc_handle* h, int32_t* first_chain, int32_t* n_local) {
  *first_chain = 1; *n_local = h->E.cfg.n_chains; return PGN_OK;
}
int orc_set_schedule(orc_handle* h, const double* beta, int32_t n, char** err) {
  if (n != h->E.cfg.n_chains) return fail(err, PGN_ERR_INVALID, "schedule length != n_chains");
  h->E.beta.assign(beta, beta + n);
  return PGN_OK;
}
int orc_set_explorer(orc_handle* h, const pgn_explorer_params* ep, char** err) {
  Engine& E = h->E;
  if (ep->n_mix < 0 || ep->n_mix > PGN_MAX_MIX) return fail(err, PGN_ERR_INVALID, "n_mix out of range");
  if ((ep->kind == PGN_EXPLORER_COMPOSE || ep->kind == PGN_EXPLORER_MIX) && (ep->n_steps < 1 || ep->n_steps > PGN_MAX_MIX))
    return fail(err, PGN_ERR_INVALID, "Compose / Mix: n_steps out of range");
  if (E.cfg.target_kind == PGN_TARGET_UNID && ep->kind != PGN_EXPLORER_SLICE)
    return fail(err, PGN_ERR_INVALID, "UNID: SliceSampler only");
  E.ep = *ep;
  E.have_std = ep->std_devs != nullptr;
  if (E.have_std) E.std_devs.assign(ep->std_devs, ep->std_devs + E.cfg.dim);
  E.ep.std_devs = nullptr;
  (void)err;
  return PGN_OK;
}
int orc_init_replicas(orc_handle* h, char** err) {
  Engine& E = h->E;
  const int N = E.N(), d = E.d();
  E.replicas.assign(N, Replica{});
  for (int i = 0; i < N; ++i) {
    Replica& r = E.replicas[i];
    r.chain = i + 1;
    r.replica_index = i + 1;
    r.rng = Philox{(uint32_t)(uint64_t)E.cfg.seed, (uint32_t)(i + 1), (uint32_t)((uint64_t)E.cfg.seed >> 32), 0u};
    r.ctr = 0;
    r.momentum.assign(d, 0.0); r.precond.assign(d, 1.0); r.start_state.assign(d, 0.0);
    r.state_before.assign(d, 0.0); r.momentum_before.assign(d, 0.0);
    r.grad.assign(d, 0.0); r.g1.assign(d, 0.0); r.g2.assign(d, 0.0);
    E.initialization(r);
  }
  (void)err;
  return PGN_OK;
}
// update_reference! has run on the host (GaussianReference.jl:22-28); install its mean / standard deviation
int orc_set_variational(orc_handle* h, const double* mean, const double* sd, char** err) {
  Engine& E = h->E;
  if (!mean || !sd) { E.var_active = false; return PGN_OK; }
  const int tk = E.cfg.target_kind, d = E.d();
  if (E.n_var() < 1) return fail(err, PGN_ERR_INVALID, "set_variational: n_chains_variational is 0");
  if (tk != PGN_TARGET_FUNNEL && tk != PGN_TARGET_GMM && tk != PGN_TARGET_UNID)
    return fail(err, PGN_ERR_INVALID, "set_variational: FUNNEL, GMM and UNID targets");
  for (int c = 0; c < d; ++c)
    if (!(sd[c] > 0.0) || !std::isfinite(sd[c]) || !std::isfinite(mean[c]))
      return fail(err, PGN_ERR_INVALID, "set_variational: finite means and positive finite standard deviations");
  E.var_mean.assign(mean, mean + d); E.var_sd.assign(sd, sd + d);
  E.var_t0.resize(d); E.var_t1.resize(d); E.var_t2.resize(d);
  for (int c = 0; c < d; ++c) {
    const double s2 = sd[c] * sd[c];
    E.var_t0[c] = -0.5 * log_(6.283185307179586 * s2);   // 2.0 * pi
    E.var_t1[c] = 1.0 / (2.0 * s2);
    E.var_t2[c] = 1.0 / s2;
  }
  E.var_active = true;
  return PGN_OK;
}

// hamiltonian_dynamics! (hamiltonian_dynamics.jl:39-84) with the identity preconditioner: n_steps x leap_frog!
int orc_hamiltonian_dynamics(orc_handle* h, const double* x, const double* p, int32_t n_points, const double* beta,
                             const double* diag_precond, double step_size, int32_t n_steps, double* x_out, double* p_out,
                             char** err) {
  Engine& E = h->E;
  const int d = E.d(), tk = E.cfg.target_kind;
  if (tk != PGN_TARGET_TOY_MVN && tk != PGN_TARGET_FUNNEL && tk != PGN_TARGET_GMM && tk != PGN_TARGET_LOGREG)
    return fail(err, PGN_ERR_INVALID, "hamiltonian_dynamics: vector targets with a gradient");
  for (int i = 0; i < n_points; ++i) {
    Replica r;
    r.chain = 1; r.replica_index = 1;
    r.x.assign(x + (size_t)i * d, x + (size_t)(i + 1) * d);
    r.momentum.assign(p + (size_t)i * d, p + (size_t)(i + 1) * d);
    r.precond.assign(d, 1.0); r.grad.assign(d, 0.0); r.g1.assign(d, 0.0); r.g2.assign(d, 0.0);
    if (diag_precond) r.precond.assign(diag_precond, diag_precond + d);
    for (int s = 0; s < n_steps; ++s)
      if (!E.leap_frog(beta[i], r, step_size)) break;
    std::memcpy(x_out + (size_t)i * d, r.x.data(), sizeof(double) * d);
    std::memcpy(p_out + (size_t)i * d, r.momentum.data(), sizeof(double) * d);
  }
  return PGN_OK;
}

int orc_get_state(orc_handle* h, pgn_replica_state* out, char** err) {
  Engine& E = h->E;
  const int N = E.N(), d = E.d();
  for (int i = 0; i < N; ++i) {
    Replica& r = E.replicas[i];
    if (E.cfg.target_kind == PGN_TARGET_ISING) E.ising_sync_x(r);
    if (out->x && d > 0) std::memcpy(out->x + (size_t)i * d, r.x.data(), sizeof(double) * d);
    if (out->replica_index) out->replica_index[i] = r.replica_index;
    if (out->rng_counter) out->rng_counter[i] = r.ctr;
    if (out->round_trip_state) out->round_trip_state[i] = r.rt_state;
  }
  (void)err;
  return PGN_OK;
}
int orc_set_state(orc_handle* h, const pgn_replica_state* in, char** err) {
  Engine& E = h->E;
  const int N = E.N(), d = E.d();
  if ((int)E.replicas.size() != N) return fail(err, PGN_ERR_INVALID, "call init_replicas first");
  for (int i = 0; i < N; ++i) {
    Replica& r = E.replicas[i];
    r.chain = i + 1;
    if (in->x && d > 0) r.x.assign(in->x + (size_t)i * d, in->x + (size_t)(i + 1) * d);
    if (in->replica_index) {
      r.replica_index = in->replica_index[i];
      r.rng.key1 = (uint32_t)r.replica_index;
    }
    if (in->rng_counter) r.ctr = in->rng_counter[i];
    if (in->round_trip_state) r.rt_state = in->round_trip_state[i];
    if (E.cfg.target_kind == PGN_TARGET_ISING && in->x) {
      const int L = E.L();
      r.rows.assign(L, 0u);
      for (int a = 0; a < L; ++a)
        for (int b = 0; b < L; ++b)
          if (r.x[(size_t)a * L + b] != 0.0) r.rows[a] |= (1u << b);
      E.ising_recompute(r);
    }
  }
  return PGN_OK;
}
int orc_run_round(orc_handle* h, int64_t n_scans, pgn_round_out* out, char** err) {
  try {
    run_round(h->E, n_scans, out);
  } catch (OrcError& e) {
    return fail(err, e.code, e.msg);
  }
  return PGN_OK;
}
int orc_log_potential(orc_handle* h, const double* x, int32_t n_points, const double* beta, double* out, char** err) {
  Engine& E = h->E;
  const int d = E.d();
  Replica r;
  r.chain = 1;   // the path of chain 1's leg: with the Gaussian reference once one is installed
  r.x.assign(d, 0.0);
  for (int i = 0; i < n_points; ++i) {
    r.x.assign(x + (size_t)i * d, x + (size_t)(i + 1) * d);
    if (E.cfg.target_kind == PGN_TARGET_ISING) {
      const int L = E.L();
      r.rows.assign(L, 0u);
      for (int a = 0; a < L; ++a)
        for (int b = 0; b < L; ++b)
          if (r.x[(size_t)a * L + b] != 0.0) r.rows[a] |= (1u << b);
      E.ising_recompute(r);
    }
    out[i] = E.log_potential(beta[i], r);
  }
  (void)err;
  return PGN_OK;
}
int orc_logdensity_and_gradient(orc_handle* h, const double* x, int32_t n_points, const double* beta,
                                double* logdens, double* grad, char** err) {
  Engine& E = h->E;
  const int d = E.d();
  if (E.cfg.target_kind == PGN_TARGET_ISING || E.cfg.target_kind == PGN_TARGET_TEST_SWAPPER || E.cfg.target_kind == PGN_TARGET_UNID)
    return fail(err, PGN_ERR_INVALID, "target has no gradient");
  Replica r;
  r.chain = 1;
  r.g1.assign(d, 0.0); r.g2.assign(d, 0.0);
  for (int i = 0; i < n_points; ++i) {
    r.x.assign(x + (size_t)i * d, x + (size_t)(i + 1) * d);
    logdens[i] = E.logdensity_and_gradient(beta[i], r, grad + (size_t)i * d);
  }
  return PGN_OK;
}
int orc_test_math(int32_t device, int32_t op, const double* in, double* out, int64_t n, int64_t seed,
                  int32_t replica_index, char** err) {
  (void)device; (void)err;
  Philox g{(uint32_t)(uint64_t)seed, (uint32_t)replica_index, (uint32_t)((uint64_t)seed >> 32), 0u};
  for (int64_t i = 0; i < n; ++i) {
    switch (op) {
      case 0: out[i] = exp_(in[i]); break;
      case 1: out[i] = log_(in[i]); break;
      case 2: out[i] = cospi_(in[i]); break;
      case 3: out[i] = normal_at(g, (uint64_t)in[i]); break;
      case 4: out[i] = uniform_at(g, (uint64_t)in[i]); break;
      case 5: out[i] = exponential_at(g, (uint64_t)in[i]); break;
      case 6: out[i] = logaddexp_(in[2 * i], in[2 * i + 1]); break;
      case 7: out[i] = log1p_(in[i]); break;
      default: out[i] = QNAN;
    }
  }
  return PGN_OK;
}
void orc_philox(const uint32_t* ctr4, const uint32_t* key2, uint32_t* out4) {
  philox4x32_10(ctr4[0], ctr4[1], ctr4[2], ctr4[3], key2[0], key2[1], out4);
}

}  // extern "C"
