// pgn_scan_vec.cu — the register-resident scan kernels of one vector-state target family.
// Compiled once per (target family, part) by csrc/Makefile:
//   -DPGN_TK=1|2|3|7 PGN_TARGET_TOY_MVN | PGN_TARGET_FUNNEL | PGN_TARGET_GMM | PGN_TARGET_MIXED (SliceSampler only)
//   -DPGN_PART=0     ToyExplorer / SliceSampler / MALA kernels + the parity entry points
//   -DPGN_PART=1     autoMALA team kernels (AutoMALA; Compose / Mix programs of ToyExplorer, SliceSampler, MALA, AutoMALA)
// so that the heavy template instantiations build in parallel.
#include "pgn_host.hpp"

#ifndef PGN_TK
#error "compile with -DPGN_TK=1|2|3|7"
#endif
#ifndef PGN_PART
#error "compile with -DPGN_PART=0|1"
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
    case PGN_EXPLORER_TOY: return scan_kernel_ptr<VecChain<PGN_TK, CPL, PGN_EXPLORER_TOY>>();
#endif
    case PGN_EXPLORER_SLICE: return scan_kernel_ptr<VecChain<PGN_TK, CPL, PGN_EXPLORER_SLICE>>();
#if PGN_TK != 7
    case PGN_EXPLORER_MALA: return scan_kernel_ptr<VecChain<PGN_TK, CPL, PGN_EXPLORER_MALA>>();
#endif
    default: return nullptr;
  }
}
template <int CPL>
void eval_points_launch(int grid, int block, size_t smem, cudaStream_t s, const Params& P, const double* xs, const double* betas,
                        int n, double* lp, double* ld, double* grad) {
  eval_points_kernel<PGN_TK, CPL><<<grid, block, smem, s>>>(P, xs, betas, n, lp, ld, grad);
}
#else
template <int CPL>
void* team_kernel_for(int ex) {
  switch (ex) {
    case PGN_EXPLORER_AUTOMALA: return scan_kernel_ptr<VecChain<PGN_TK, CPL, PGN_EXPLORER_AUTOMALA>>();
    case PGN_EXPLORER_COMPOSE: case PGN_EXPLORER_MIX: return scan_kernel_ptr<VecChain<PGN_TK, CPL, PGN_EXPLORER_COMPOSE>>();
    default: return nullptr;
  }
}
#endif

}  // namespace

#if PGN_TK == 1
#define PGN_FAMILY(name) name##_toy
#elif PGN_TK == 2
#define PGN_FAMILY(name) name##_funnel
#elif PGN_TK == 3
#define PGN_FAMILY(name) name##_gmm
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
#else
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
