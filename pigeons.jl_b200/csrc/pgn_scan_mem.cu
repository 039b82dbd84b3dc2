// pgn_scan_mem.cu — the memory-resident scan kernels (pgn_memchain.cuh): any dimension, any
// number of chains per GPU.
#include "pgn_host.hpp"
#include "pgn_memchain.cuh"

namespace pgn {
namespace {
template <int TK>
void* mem_kernel_for(int ex) {
  switch (ex) {
    case PGN_EXPLORER_TOY: return TK == PGN_TARGET_TOY_MVN ? (void*)scan_kernel_mem<TK, PGN_EXPLORER_TOY> : nullptr;
    case PGN_EXPLORER_SLICE: return (void*)scan_kernel_mem<TK, PGN_EXPLORER_SLICE>;
    case PGN_EXPLORER_AUTOMALA: return (void*)scan_kernel_mem<TK, PGN_EXPLORER_AUTOMALA>;
    case PGN_EXPLORER_MALA: return (void*)scan_kernel_mem<TK, PGN_EXPLORER_MALA>;
    default: return nullptr;
  }
}
}  // namespace

void* mem_scan_kernel(int target_kind, int ex) {
  switch (target_kind) {
    case PGN_TARGET_TOY_MVN: return mem_kernel_for<PGN_TARGET_TOY_MVN>(ex);
    case PGN_TARGET_FUNNEL: return mem_kernel_for<PGN_TARGET_FUNNEL>(ex);
    case PGN_TARGET_GMM: return mem_kernel_for<PGN_TARGET_GMM>(ex);
    default: return nullptr;
  }
}

void launch_eval_points_mem(int target_kind, int grid, int block, cudaStream_t s, const MemParams& MP, const double* xs,
                            const double* betas, int n, double* lp, double* ld, double* grad) {
  switch (target_kind) {
    case PGN_TARGET_TOY_MVN: eval_points_mem_kernel<PGN_TARGET_TOY_MVN><<<grid, block, 0, s>>>(MP, xs, betas, n, lp, ld, grad); break;
    case PGN_TARGET_FUNNEL: eval_points_mem_kernel<PGN_TARGET_FUNNEL><<<grid, block, 0, s>>>(MP, xs, betas, n, lp, ld, grad); break;
    case PGN_TARGET_GMM: eval_points_mem_kernel<PGN_TARGET_GMM><<<grid, block, 0, s>>>(MP, xs, betas, n, lp, ld, grad); break;
    default: throw CudaError{PGN_ERR_INVALID, "unsupported target"};
  }
}

}  // namespace pgn
