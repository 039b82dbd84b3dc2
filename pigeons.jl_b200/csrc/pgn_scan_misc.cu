// pgn_scan_misc.cu — Ising / TestSwapper scan kernels and the small helper kernels
// (initialisation, Ising parity entry point, numerics self-test).
#include "pgn_host.hpp"

namespace pgn {

// initialization(target, rng, replica_index) for toy MVN (toy_mvn_target.jl:10-11):
// x = randn(rng, dim) / sqrt(precision1); one warp per replica.
__global__ void init_toy_kernel(const __grid_constant__ Params P) {
  const int lane = threadIdx.x & 31;
  const int wl = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (wl >= P.n_local) return;
  Rng rng{P.seed_lo, (unsigned int)(P.first_chain + wl), P.seed_hi, 0u, 0ull};
  const double sq = sqrt(P.p[1]);
  for (int c = lane; c < P.d; c += 32) P.x[(size_t)wl * P.d_pad + c] = normal_at(rng, (unsigned long long)c) / sq;
  if (lane == 0) P.rng_ctr[wl] = (unsigned long long)P.d;
}

__global__ void ising_lp_kernel(const __grid_constant__ Params P, const double* xs, const double* betas, int n_points, double* lp_out) {
  const int lane = threadIdx.x & 31;
  const int w = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (w >= n_points) return;
  IsingChain ch;
  ch.P = &P; ch.lane = lane; ch.L = (int)P.p[1];
  unsigned int r = 0u;
  if (lane < ch.L)
    for (int j = 0; j < ch.L; ++j)
      if (xs[(size_t)w * P.d + lane * ch.L + j] != 0.0) r |= (1u << j);
  ch.row = r;
  ch.recompute_S();
  if (lane == 0) lp_out[w] = ch.lp(betas[w], ch.S);
}

__global__ void test_math_kernel(int op, const double* in, double* out, long long n, unsigned int seed_lo,
                                 unsigned int seed_hi, unsigned int replica_index) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  Rng g{seed_lo, replica_index, seed_hi, 0u, 0ull};
  double r;
  switch (op) {
    case 0: r = exp_(in[i]); break;
    case 1: r = log_(in[i]); break;
    case 2: r = cospi_(in[i]); break;
    case 3: r = normal_at(g, (unsigned long long)in[i]); break;
    case 4: r = uniform_at(g, (unsigned long long)in[i]); break;
    case 5: r = exponential_at(g, (unsigned long long)in[i]); break;
    case 6: r = logaddexp_(in[2 * i], in[2 * i + 1]); break;
    case 7: r = log1p_<false>(in[i]); break;
    default: r = PGN_NAN;
  }
  out[i] = r;
}

// reduce_recorders! (src/recorders/recorders.jl:88-120) for the per-replica recorder tables: the binary tree of
// reduce_deterministically (src/mpi_utils/Entangler.jl:214-277) over the replica indices 1..N — at the level with
// spacing s, entry i absorbs entry i + s for i = 1, 1 + 2s, ...; an unpaired last entry waits for the next level —
// applied to the column of one local chain.  One warp per chain: the merges of a level are independent, lane l takes
// every 32nd of them.  The result lands in the chain's ChainStatsDev.
__global__ void merge_recorders_kernel(RecEntry* table, int n_replicas, int n_local, ChainStatsDev* out) {
  const int lane = threadIdx.x & 31;
  const int c = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (c >= n_local) return;
  for (int spacing = 1; spacing < n_replicas; spacing *= 2) {
    for (long long i = (long long)lane * 2 * spacing; i + spacing < n_replicas; i += 64LL * spacing) {
      RecEntry* a = table + (size_t)i * n_local + c;
      const RecEntry* b = table + (size_t)(i + spacing) * n_local + c;
      if ((b->expl_acc.n | b->am.n | b->rev.n | b->swap_acc.n | b->ls_fwd.n | b->ls_bwd.n) == 0) continue;   // replica never sat here
      RecEntry ea = *a;
      const RecEntry eb = *b;
      merge_mean(ea.expl_acc, eb.expl_acc); merge_mean(ea.am, eb.am); merge_mean(ea.rev, eb.rev);
      merge_mean(ea.swap_acc, eb.swap_acc);
      merge_logsum(ea.ls_fwd, eb.ls_fwd); merge_logsum(ea.ls_bwd, eb.ls_bwd);
      *a = ea;
    }
    __syncwarp();
  }
  if (lane == 0) {
    const RecEntry e = table[c];
    ChainStatsDev& o = out[c];
    o.swap_n = e.swap_acc.n; o.swap_mean = e.swap_acc.mu;
    o.ls_fwd = e.ls_fwd.n > 0 ? e.ls_fwd.value : -PGN_INF; o.ls_bwd = e.ls_bwd.n > 0 ? e.ls_bwd.value : -PGN_INF;
    o.expl_acc_n = e.expl_acc.n; o.expl_acc_mean = e.expl_acc.mu;
    o.am_n = e.am.n; o.am_mean = e.am.mu; o.rev_n = e.rev.n; o.rev_mean = e.rev.mu;
  }
}
// The same tree for the target-chain online statistics, one warp per coordinate (OnlineStatsBase._merge!(::Variance,
// ::Variance): g = n2 / (n += n2); delta = mu2 - mu; s2 = smooth(s2, s2', g) + delta^2 g (1 - g); mu = smooth(mu, mu2, g)).
__global__ void merge_online_kernel(OnEntry* table, int n_replicas, int d, int d_pad, double* mean, double* s2, long long* n_out) {
  const int lane = threadIdx.x & 31;
  const int c = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (c >= d) return;
  for (int spacing = 1; spacing < n_replicas; spacing *= 2) {
    for (long long i = (long long)lane * 2 * spacing; i + spacing < n_replicas; i += 64LL * spacing) {
      OnEntry* a = table + (size_t)i * d_pad + c;
      const OnEntry b = table[(size_t)(i + spacing) * d_pad + c];
      if (b.n == 0) continue;
      OnEntry ea = *a;
      if (ea.n == 0) { *a = b; continue; }
      ea.n += b.n;
      const double g = (double)b.n / (double)ea.n;
      const double delta = b.mu - ea.mu;
      ea.s2 = (ea.s2 + g * (b.s2 - ea.s2)) + ((delta * delta) * g) * (1.0 - g);
      ea.mu = ea.mu + g * (b.mu - ea.mu);
      *a = ea;
    }
    __syncwarp();
  }
  if (lane == 0) {
    const OnEntry e = table[c];
    mean[c] = e.mu; s2[c] = e.s2;
    if (c == 0) *n_out = e.n;
  }
}
// "absent" entries for a new round (recorders are emptied every round, recorders.jl:113-118)
__global__ void init_recorder_tables_kernel(RecEntry* table, size_t n) {
  RecEntry e;
  e.expl_acc = MeanAcc{0, 0.0}; e.am = e.expl_acc; e.rev = e.expl_acc; e.swap_acc = e.expl_acc;
  e.ls_fwd = LogSumAcc{0, -PGN_INF}; e.ls_bwd = e.ls_fwd;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) table[i] = e;
}
void launch_init_recorder_tables(cudaStream_t s, RecEntry* table, size_t n, int n_sms) {
  init_recorder_tables_kernel<<<n_sms * 4, 256, 0, s>>>(table, n);
}
void launch_merge_recorders(cudaStream_t s, RecEntry* table, int n_replicas, int n_local, ChainStatsDev* out) {
  merge_recorders_kernel<<<(n_local + 3) / 4, 128, 0, s>>>(table, n_replicas, n_local, out);
}
void launch_merge_online(cudaStream_t s, OnEntry* table, int n_replicas, int d, int d_pad, double* mean, double* s2, long long* n_out) {
  merge_online_kernel<<<(d + 3) / 4, 128, 0, s>>>(table, n_replicas, d, d_pad, mean, s2, n_out);
}

void* ising_scan_kernel() { return (void*)scan_kernel<IsingChain>; }
void* ising_lite_scan_kernel() { return (void*)scan_kernel<IsingChainLite>; }
void* test_swapper_scan_kernel() { return (void*)scan_kernel<TestSwapperChain>; }

void launch_init_toy(int grid, int block, cudaStream_t s, const Params& P) { init_toy_kernel<<<grid, block, 0, s>>>(P); }
void launch_ising_lp(int grid, int block, cudaStream_t s, const Params& P, const double* xs, const double* betas, int n,
                     double* lp) {
  ising_lp_kernel<<<grid, block, 0, s>>>(P, xs, betas, n, lp);
}
void launch_test_math(int grid, int block, int op, const double* in, double* out, long long n, unsigned int seed_lo,
                      unsigned int seed_hi, unsigned int replica_index) {
  test_math_kernel<<<grid, block>>>(op, in, out, n, seed_lo, seed_hi, replica_index);
}

// the vector-state families: the plain and the team kernels of a family are compiled separately
void* vec_plain_kernel_toy(int cpl, int ex);    void* vec_team_kernel_toy(int cpl, int ex);
void* vec_plain_kernel_funnel(int cpl, int ex); void* vec_team_kernel_funnel(int cpl, int ex);
void* vec_plain_kernel_gmm(int cpl, int ex);    void* vec_team_kernel_gmm(int cpl, int ex);
void* vec_plain_kernel_mixed(int cpl, int ex);
void* vec_plain_kernel_unid(int cpl, int ex);
void* vec_plain_kernel_unid_var(int cpl, int ex);
void* vec_plain_kernel_funnel_var(int cpl, int ex); void* vec_team_kernel_funnel_var(int cpl, int ex);
void* vec_plain_kernel_gmm_var(int cpl, int ex);    void* vec_team_kernel_gmm_var(int cpl, int ex);
static bool is_team_explorer(int ex) { return ex == PGN_EXPLORER_AUTOMALA || ex == PGN_EXPLORER_COMPOSE || ex == PGN_EXPLORER_MIX; }
void* vec_scan_kernel_toy(int cpl, int ex) { return is_team_explorer(ex) ? vec_team_kernel_toy(cpl, ex) : vec_plain_kernel_toy(cpl, ex); }
void* vec_scan_kernel_funnel(int cpl, int ex) { return is_team_explorer(ex) ? vec_team_kernel_funnel(cpl, ex) : vec_plain_kernel_funnel(cpl, ex); }
void* vec_scan_kernel_mixed(int cpl, int ex) { return ex == PGN_EXPLORER_SLICE ? vec_plain_kernel_mixed(cpl, ex) : nullptr; }
void* vec_scan_kernel_unid(int cpl, int ex) { return ex == PGN_EXPLORER_SLICE ? vec_plain_kernel_unid(cpl, ex) : nullptr; }
void* vec_scan_kernel_unid_var(int cpl, int ex) { return ex == PGN_EXPLORER_SLICE ? vec_plain_kernel_unid_var(cpl, ex) : nullptr; }
void* vec_scan_kernel_gmm(int cpl, int ex) { return is_team_explorer(ex) ? vec_team_kernel_gmm(cpl, ex) : vec_plain_kernel_gmm(cpl, ex); }
void* vec_scan_kernel_funnel_var(int cpl, int ex) { return is_team_explorer(ex) ? vec_team_kernel_funnel_var(cpl, ex) : vec_plain_kernel_funnel_var(cpl, ex); }
void* vec_scan_kernel_gmm_var(int cpl, int ex) { return is_team_explorer(ex) ? vec_team_kernel_gmm_var(cpl, ex) : vec_plain_kernel_gmm_var(cpl, ex); }

// GaussianReference tables from (mean, sd): tab = [5][d_pad] mean | sd | -0.5 log(2 pi sd^2) | 1/(2 sd^2) | 1/sd^2
// (gaussian_logdensity, GaussianReference.jl:45-53; 2.0 * pi is exact in binary)
__global__ void var_tables_kernel(double* tab, int d, int d_pad) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= d) return;
  const double sd = tab[(size_t)1 * d_pad + c];
  const double s2 = sd * sd;
  tab[(size_t)2 * d_pad + c] = -0.5 * log_(6.283185307179586 * s2);
  tab[(size_t)3 * d_pad + c] = 1.0 / (2.0 * s2);
  tab[(size_t)4 * d_pad + c] = 1.0 / s2;
}
void launch_var_tables(cudaStream_t s, double* tab, int d, int d_pad) {
  var_tables_kernel<<<(d + 127) / 128, 128, 0, s>>>(tab, d, d_pad);
}

}  // namespace pgn
