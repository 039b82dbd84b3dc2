// pgn_scan_vec.cu — the register-resident scan kernels of one vector-state target family.
// Compiled once per (target family, part) by csrc/Makefile:
//   -DPGN_TK=1|2|3|7|8 PGN_TARGET_TOY_MVN | PGN_TARGET_FUNNEL | PGN_TARGET_GMM | PGN_TARGET_MIXED, PGN_TARGET_UNID (SliceSampler only)
//   -DPGN_PART=0     ToyExplorer / SliceSampler / MALA kernels + the parity entry points
//   -DPGN_PART=1     autoMALA team kernels (AutoMALA; Compose / Mix programs of ToyExplorer, SliceSampler, MALA, AutoMALA)
// so that the heavy template instantiations build in parallel.
#include "pgn_host.hpp"

#ifndef PGN_TK
#error "compile with -DPGN_TK=1|2|3|7|8"
#endif
#ifndef PGN_PART
#error "compile with -DPGN_PART=0|1"
#endif
#ifndef PGN_VAR
#define PGN_VAR 0   // 1: the kernels of a ladder whose variational leg uses a GaussianReference (FUNNEL, GMM)
#endif

namespace pgn {
namespace {

template <class Chain>
void* scan_kernel_ptr() { return (void*)scan_kernel<Chain>; }

#if PGN_PART == 0
template <int CPL>
void* plain_kernel_for(int ex) {
  switch (ex) {
#if PGN_TK == 1
    case PGN_EXPLORER_TOY: return scan_kernel_ptr<VecChain<PGN_TK, CPL, PGN_EXPLORER_TOY, PGN_VAR != 0>>();
#endif
    case PGN_EXPLORER_SLICE: return scan_kernel_ptr<VecChain<PGN_TK, CPL, PGN_EXPLORER_SLICE, PGN_VAR != 0>>();
#if PGN_TK < 7
    case PGN_EXPLORER_MALA: return scan_kernel_ptr<VecChain<PGN_TK, CPL, PGN_EXPLORER_MALA, PGN_VAR != 0>>();
#endif
    default: return nullptr;
  }
}
#if PGN_TK < 7
template <int CPL>
void leapfrog_launch(int grid, int block, size_t smem, cudaStream_t s, const Params& P, const double* xs, const double* ps,
                     const double* betas, const double* precond, double eps, int n_steps, int n, double* x_out, double* p_out) {
  leapfrog_kernel<PGN_TK, CPL, PGN_VAR != 0><<<grid, block, smem, s>>>(P, xs, ps, betas, precond, eps, n_steps, n, x_out, p_out);
}
#endif
template <int CPL>
void eval_points_launch(int grid, int block, size_t smem, cudaStream_t s, const Params& P, const double* xs, const double* betas,
                        int n, double* lp, double* ld, double* grad) {
  eval_points_kernel<PGN_TK, CPL, PGN_VAR != 0><<<grid, block, smem, s>>>(P, xs, betas, n, lp, ld, grad);
}
#else
#if !PGN_VAR && PGN_TK <= 3
// mixed blocks (two warps: one chain's team of two, or two single-warp chains): autoMALA on the gradient targets
template <int CPL>
void* mixed_team_kernel_for() { return scan_kernel_ptr<VecChain<PGN_TK, CPL, PGN_EXPLORER_AUTOMALA, false, true>>(); }
#endif
template <int CPL>
void* team_kernel_for(int ex) {
  switch (ex) {
    case PGN_EXPLORER_AUTOMALA: return scan_kernel_ptr<VecChain<PGN_TK, CPL, PGN_EXPLORER_AUTOMALA, PGN_VAR != 0>>();
#if !PGN_VAR   // Compose / Mix programs are not built for the variational leg
    case PGN_EXPLORER_COMPOSE: case PGN_EXPLORER_MIX: return scan_kernel_ptr<VecChain<PGN_TK, CPL, PGN_EXPLORER_COMPOSE>>();
#endif
    default: return nullptr;
  }
}
#endif

}  // namespace

#if PGN_TK == 1
#define PGN_FAMILY(name) name##_toy
#elif PGN_TK == 2 && PGN_VAR
#define PGN_FAMILY(name) name##_funnel_var
#elif PGN_TK == 3 && PGN_VAR
#define PGN_FAMILY(name) name##_gmm_var
#elif PGN_TK == 2
#define PGN_FAMILY(name) name##_funnel
#elif PGN_TK == 3
#define PGN_FAMILY(name) name##_gmm
#elif PGN_TK == 8 && PGN_VAR
#define PGN_FAMILY(name) name##_unid_var
#elif PGN_TK == 8
#define PGN_FAMILY(name) name##_unid
#else
#define PGN_FAMILY(name) name##_mixed
#endif

#if PGN_PART == 0
void* PGN_FAMILY(vec_plain_kernel)(int cpl, int ex) {
  switch (cpl) {
    case 1: return plain_kernel_for<1>(ex);
    case 2: return plain_kernel_for<2>(ex);
    case 4: return plain_kernel_for<4>(ex);
    default: return nullptr;
  }
}
void PGN_FAMILY(launch_eval_points)(int cpl, int grid, int block, size_t smem, cudaStream_t s, const Params& P, const double* xs,
                                    const double* betas, int n, double* lp, double* ld, double* grad) {
  switch (cpl) {
    case 1: eval_points_launch<1>(grid, block, smem, s, P, xs, betas, n, lp, ld, grad); break;
    case 2: eval_points_launch<2>(grid, block, smem, s, P, xs, betas, n, lp, ld, grad); break;
    case 4: eval_points_launch<4>(grid, block, smem, s, P, xs, betas, n, lp, ld, grad); break;
    default: throw CudaError{PGN_ERR_INVALID, "unsupported dimension"};
  }
}
#if PGN_TK < 7
void PGN_FAMILY(launch_leapfrog)(int cpl, int grid, int block, size_t smem, cudaStream_t s, const Params& P, const double* xs,
                                 const double* ps, const double* betas, const double* precond, double eps, int n_steps, int n, double* x_out, double* p_out) {
  switch (cpl) {
    case 1: leapfrog_launch<1>(grid, block, smem, s, P, xs, ps, betas, precond, eps, n_steps, n, x_out, p_out); break;
    case 2: leapfrog_launch<2>(grid, block, smem, s, P, xs, ps, betas, precond, eps, n_steps, n, x_out, p_out); break;
    case 4: leapfrog_launch<4>(grid, block, smem, s, P, xs, ps, betas, precond, eps, n_steps, n, x_out, p_out); break;
    default: throw CudaError{PGN_ERR_INVALID, "unsupported dimension"};
  }
}
#endif
#else
#if !PGN_VAR && PGN_TK <= 3
void* PGN_FAMILY(vec_mixed_team_kernel)(int cpl) {
  switch (cpl) {
    case 1: return mixed_team_kernel_for<1>();
    case 2: return mixed_team_kernel_for<2>();
    case 4: return mixed_team_kernel_for<4>();
    default: return nullptr;
  }
}
#endif
void* PGN_FAMILY(vec_team_kernel)(int cpl, int ex) {
  switch (cpl) {
    case 1: return team_kernel_for<1>(ex);
    case 2: return team_kernel_for<2>(ex);
    case 4: return team_kernel_for<4>(ex);
    default: return nullptr;
  }
}
#endif

}  // namespace pgn
