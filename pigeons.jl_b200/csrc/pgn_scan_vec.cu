// pgn_scan_vec.cu — the register-resident scan kernels of one vector-state target family.
// Compiled once per (target family, part) by csrc/Makefile:
//   -DPGN_TK=1|2|3   PGN_TARGET_TOY_MVN | PGN_TARGET_FUNNEL | PGN_TARGET_GMM
//   -DPGN_PART=0     ToyExplorer / SliceSampler / MALA kernels + the parity entry points
//   -DPGN_PART=1     autoMALA team kernels (AutoMALA, Compose(SliceSampler, AutoMALA))
// so that the heavy template instantiations build in parallel.
#include "pgn_host.hpp"

#ifndef PGN_TK
#error "compile with -DPGN_TK=1|2|3"
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
    case PGN_EXPLORER_TOY: return PGN_TK == PGN_TARGET_TOY_MVN ? scan_kernel_ptr<VecChain<PGN_TK, CPL, PGN_EXPLORER_TOY>>() : nullptr;
    case PGN_EXPLORER_SLICE: return scan_kernel_ptr<VecChain<PGN_TK, CPL, PGN_EXPLORER_SLICE>>();
    case PGN_EXPLORER_MALA: return scan_kernel_ptr<VecChain<PGN_TK, CPL, PGN_EXPLORER_MALA>>();
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
void* team_kernel_for(int ex, int regcap) {
  // four coordinates per lane at 128 registers per thread (instead of 255): twice the co-resident warps, i.e. room for
  // teams of two where the uncapped kernel fits one warp per chain (BASELINE config 3: 1024 chains of d = 128)
  if (CPL == 4 && PGN_TK != PGN_TARGET_FUNNEL && ex == PGN_EXPLORER_AUTOMALA && regcap == 128)
    return scan_kernel_ptr<CappedChain<VecChain<PGN_TK, (CPL == 4 ? 4 : 1), PGN_EXPLORER_AUTOMALA>, 128, 4>>();
  switch (ex) {
    case PGN_EXPLORER_AUTOMALA: return scan_kernel_ptr<VecChain<PGN_TK, CPL, PGN_EXPLORER_AUTOMALA>>();
    case PGN_EXPLORER_SLICE_THEN_AUTOMALA: return scan_kernel_ptr<VecChain<PGN_TK, CPL, PGN_EXPLORER_SLICE_THEN_AUTOMALA>>();
    default: return nullptr;
  }
}
#endif

}  // namespace

#if PGN_TK == 1
#define PGN_FAMILY(name) name##_toy
#elif PGN_TK == 2
#define PGN_FAMILY(name) name##_funnel
#else
#define PGN_FAMILY(name) name##_gmm
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
void* PGN_FAMILY(vec_team_kernel)(int cpl, int ex, int regcap) {
  switch (cpl) {
    case 1: return team_kernel_for<1>(ex, regcap);
    case 2: return team_kernel_for<2>(ex, regcap);
    case 4: return team_kernel_for<4>(ex, regcap);
    default: return nullptr;
  }
}
#endif

}  // namespace pgn
