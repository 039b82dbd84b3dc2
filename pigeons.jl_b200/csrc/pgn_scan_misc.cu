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
    default: r = PGN_NAN;
  }
  out[i] = r;
}

void* ising_scan_kernel() { return (void*)scan_kernel<IsingChain>; }
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
void* vec_plain_kernel_toy(int cpl, int ex);    void* vec_team_kernel_toy(int cpl, int ex, int regcap);
void* vec_plain_kernel_funnel(int cpl, int ex); void* vec_team_kernel_funnel(int cpl, int ex, int regcap);
void* vec_plain_kernel_gmm(int cpl, int ex);    void* vec_team_kernel_gmm(int cpl, int ex, int regcap);
static bool is_team_explorer(int ex) { return ex == PGN_EXPLORER_AUTOMALA || ex == PGN_EXPLORER_SLICE_THEN_AUTOMALA; }
void* vec_scan_kernel_toy(int cpl, int ex, int regcap) { return is_team_explorer(ex) ? vec_team_kernel_toy(cpl, ex, regcap) : vec_plain_kernel_toy(cpl, ex); }
void* vec_scan_kernel_funnel(int cpl, int ex, int regcap) { return is_team_explorer(ex) ? vec_team_kernel_funnel(cpl, ex, regcap) : vec_plain_kernel_funnel(cpl, ex); }
void* vec_scan_kernel_gmm(int cpl, int ex, int regcap) { return is_team_explorer(ex) ? vec_team_kernel_gmm(cpl, ex, regcap) : vec_plain_kernel_gmm(cpl, ex); }

}  // namespace pgn
