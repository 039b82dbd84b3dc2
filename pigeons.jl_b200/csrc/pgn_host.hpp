// pgn_host.hpp — host-side types shared by the translation units of libpigeons_b200.so:
// the engine handle, device buffers, error plumbing, and the functions through which
// pgn_engine.cu (the C ABI) reaches the kernels compiled in the other translation units.
//
// The library is built from several .cu files (csrc/Makefile) so that the heavy kernel
// families compile in parallel; every kernel is launched from the translation unit that
// defines it (no relocatable device code).
#pragma once
#include <cuda_runtime.h>

#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "pgn_kernels.cuh"
#include "pgn_logreg_types.cuh"
#include "pgn_memchain_types.cuh"

namespace pgn {

struct CudaError {
  int code;
  std::string msg;
};

#define CUDA_CHECK(expr)                                                                          \
  do {                                                                                            \
    cudaError_t e_ = (expr);                                                                      \
    if (e_ != cudaSuccess)                                                                        \
      throw ::pgn::CudaError{PGN_ERR_CUDA, std::string(#expr) + ": " + cudaGetErrorString(e_)};   \
  } while (0)

template <class T>
struct DevBuf {
  T* p = nullptr;
  size_t n = 0;
  void alloc(size_t count, bool zero = true) {
    release();
    n = count;
    if (count == 0) count = 1;
    CUDA_CHECK(cudaMalloc(&p, count * sizeof(T)));
    if (zero) CUDA_CHECK(cudaMemset(p, 0, count * sizeof(T)));
  }
  // stream-ordered allocation (first use inside a round: must not wait for another handle's kernel, see StreamBuf)
  void alloc_on(size_t count, cudaStream_t s) {
    release();
    n = count; async_stream = s; is_async = true;
    if (count == 0) count = 1;
    CUDA_CHECK(cudaMallocAsync(&p, count * sizeof(T), s));
    CUDA_CHECK(cudaMemsetAsync(p, 0, count * sizeof(T), s));
  }
  void release() {
    if (p) { if (is_async) cudaFreeAsync(p, async_stream); else cudaFree(p); }
    p = nullptr; n = 0; is_async = false;
  }
  cudaStream_t async_stream = nullptr;
  bool is_async = false;
  void upload(const T* h, size_t count) { CUDA_CHECK(cudaMemcpy(p, h, count * sizeof(T), cudaMemcpyHostToDevice)); }
  void download(T* h, size_t count) const { CUDA_CHECK(cudaMemcpy(h, p, count * sizeof(T), cudaMemcpyDeviceToHost)); }
  // the same on a stream of the caller's (followed by a wait for THAT stream only, not for the device)
  void upload(const T* h, size_t count, cudaStream_t s) {
    CUDA_CHECK(cudaMemcpyAsync(p, h, count * sizeof(T), cudaMemcpyHostToDevice, s));
    CUDA_CHECK(cudaStreamSynchronize(s));
  }
  void download(T* h, size_t count, cudaStream_t s) const {
    CUDA_CHECK(cudaMemcpyAsync(h, p, count * sizeof(T), cudaMemcpyDeviceToHost, s));
    CUDA_CHECK(cudaStreamSynchronize(s));
  }
  DevBuf() = default;
  DevBuf(const DevBuf&) = delete;
  DevBuf& operator=(const DevBuf&) = delete;
  ~DevBuf() { release(); }
};

// Per-round scratch (event logs, parity entry points), allocated and released in stream order on the handle's own
// stream (cudaMallocAsync / cudaFreeAsync): unlike cudaMalloc / cudaFree, which may wait for the whole device, these
// never wait for the kernel of ANOTHER handle — several handles of one process run their rounds concurrently and
// their kernels wait for each other's mailbox posts (pgn_peer_attach).
template <class T>
struct StreamBuf {
  T* p = nullptr;
  size_t n = 0;
  cudaStream_t s = nullptr;
  void alloc(size_t count, cudaStream_t stream, bool zero = false) {
    release();
    n = count; s = stream;
    if (count == 0) count = 1;
    CUDA_CHECK(cudaMallocAsync(&p, count * sizeof(T), stream));
    if (zero) CUDA_CHECK(cudaMemsetAsync(p, 0, count * sizeof(T), stream));
  }
  void release() {
    if (p) cudaFreeAsync(p, s);
    p = nullptr; n = 0;
  }
  void upload(const T* h, size_t count) {
    CUDA_CHECK(cudaMemcpyAsync(p, h, count * sizeof(T), cudaMemcpyHostToDevice, s));
    CUDA_CHECK(cudaStreamSynchronize(s));
  }
  void download(T* h, size_t count) const {
    CUDA_CHECK(cudaMemcpyAsync(h, p, count * sizeof(T), cudaMemcpyDeviceToHost, s));
    CUDA_CHECK(cudaStreamSynchronize(s));
  }
  StreamBuf() = default;
  StreamBuf(const StreamBuf&) = delete;
  StreamBuf& operator=(const StreamBuf&) = delete;
  ~StreamBuf() { release(); }
};

}  // namespace pgn

struct pgn_handle {
  pgn_config cfg{};
  pgn_explorer_params ep{};
  bool have_std = false;
  int first_chain = 1, n_local = 0;
  int cpl = 1, d_pad = 32, pay_doubles = 32;
  size_t slot_bytes = 0, mail_bytes = 0;
  unsigned int epoch = 0;
  unsigned long long scan_seq = 0;   // scans run so far (all rounds): base of the mailbox tags
  int n_sms = 0;
  pgn::DevBuf<double> beta, x, means, log_w, std_devs, online_mean, online_s2;
  pgn::DevBuf<int> replica_index, rt_state, error_flag;
  pgn::DevBuf<unsigned long long> rng_ctr;
  pgn::DevBuf<long long> online_n;
  pgn::DevBuf<unsigned long long> progress;
  pgn::DevBuf<pgn::ChainStatsDev> stats;
  pgn::DevBuf<char> mail;
  char* mail_left = nullptr;
  char* mail_right = nullptr;
  bool left_is_ipc = false, right_is_ipc = false;
  cudaStream_t stream = nullptr;
  cudaEvent_t ev0 = nullptr, ev1 = nullptr;
  bool initialised = false;
  unsigned long long timeout_ns = 20ull * 1000ull * 1000ull * 1000ull;
  // ---- memory-resident scan path (any d, any number of chains; pgn_memchain.cuh)
  bool force_mem = false;
  int recorder_order = PGN_RECORDERS_PER_REPLICA;
  std::vector<long long> last_explore;  // explore cycles per local chain in the previous round (mixed teams: who is slowest)
  int last_mixed_teams = 0;             // chains that ran as a team of two in the last round (diagnostics)
  pgn::DevBuf<double> var_tab;          // GaussianReference tables [5][d_pad] (pgn_set_variational)
  bool var_active = false;
  pgn::DevBuf<pgn::RecEntry> rec_table;   // per-replica recorders [n_chains][n_local] (PGN_RECORDERS_PER_REPLICA)
  pgn::DevBuf<pgn::OnEntry> on_table;     // target-chain online statistics per replica [n_chains][d_pad]
  pgn::DevBuf<pgn::MemRec> mem_rec;
  pgn::DevBuf<double> mem_vec[10];
  bool mem_allocated = false;
  // ---- logistic regression (batched GEMM path, pgn_logreg_host.cu)
  int lr_n_data = 0, lr_n_pad = 0, lr_r_pad = 0, lr_splits = 0;
  pgn::DevBuf<double> lr_Xr, lr_Xt, lr_y, lr_Theta, lr_Thetat, lr_LL, lr_Res, lr_lik, lr_Gp, lr_G;
  pgn::DevBuf<double> lr_P, lr_G0, lr_SX, lr_SP, lr_SG, lr_TP, lr_TG, lr_FX, lr_FG, lr_QX, lr_QP, lr_QG;
  pgn::DevBuf<pgn::LrChainState> lr_st;
  pgn::DevBuf<int> lr_cols;              // compacted list of the chains whose pending point is evaluated in this batch step
  pgn::DevBuf<pgn::LrControl> lr_ctl;    // device-side counters of the batched evaluation loop
  int lr_dmma_bn = 64;           // column tile of the tensor-core GEMM (PGN_DMMA_BN)
  bool lr_use_dmma = true;       // FP64 tensor-core GEMM (same summation order as the SIMT kernel, see pgn_logreg.cuh)
  double last_gemm_ms = 0.0;     // device time spent in the two GEMMs during the last round
  long long last_batch_steps = 0;
  long long last_launches = 0;      // kernels launched during the last round
  long long last_active_cols = 0;   // sum over batch steps of the number of chains that asked for an evaluation
  long long last_gemm_cols = 0;     // sum over batch steps of the number of columns the GEMMs multiplied
};

namespace pgn {

// ---- kernels compiled in the other translation units -------------------------------------------
// scan kernels (pgn_scan_vec.cu, one object per target family; pgn_scan_misc.cu; pgn_scan_mem.cu)
void* vec_scan_kernel_toy(int cpl, int ex);
void* vec_scan_kernel_funnel(int cpl, int ex);
void* vec_scan_kernel_gmm(int cpl, int ex);
void* vec_scan_kernel_mixed(int cpl, int ex);
// "mixed teams" variants of the autoMALA kernels (blocks of two warps: a team of two, or two single-warp chains)
void* vec_mixed_team_kernel_toy(int cpl);
void* vec_mixed_team_kernel_funnel(int cpl);
void* vec_mixed_team_kernel_gmm(int cpl);
void* vec_scan_kernel_unid(int cpl, int ex);
void* vec_scan_kernel_unid_var(int cpl, int ex);
void* vec_scan_kernel_funnel_var(int cpl, int ex);   // ladders whose variational leg uses a GaussianReference
void* vec_scan_kernel_gmm_var(int cpl, int ex);
void launch_var_tables(cudaStream_t s, double* tab, int d, int d_pad);
void* ising_scan_kernel();
void* ising_lite_scan_kernel();
void* test_swapper_scan_kernel();
void* mem_scan_kernel(int target_kind, int ex);
// parity entry points
void launch_eval_points_toy(int cpl, int grid, int block, size_t smem, cudaStream_t s, const Params& P, const double* xs,
                            const double* betas, int n, double* lp, double* ld, double* grad);
void launch_eval_points_funnel(int cpl, int grid, int block, size_t smem, cudaStream_t s, const Params& P, const double* xs,
                               const double* betas, int n, double* lp, double* ld, double* grad);
void launch_eval_points_gmm(int cpl, int grid, int block, size_t smem, cudaStream_t s, const Params& P, const double* xs,
                            const double* betas, int n, double* lp, double* ld, double* grad);
void launch_leapfrog_toy(int cpl, int grid, int block, size_t smem, cudaStream_t s, const Params& P, const double* xs, const double* ps,
                         const double* betas, const double* precond, double eps, int n_steps, int n, double* x_out, double* p_out);
void launch_leapfrog_funnel(int cpl, int grid, int block, size_t smem, cudaStream_t s, const Params& P, const double* xs, const double* ps,
                            const double* betas, const double* precond, double eps, int n_steps, int n, double* x_out, double* p_out);
void launch_leapfrog_gmm(int cpl, int grid, int block, size_t smem, cudaStream_t s, const Params& P, const double* xs, const double* ps,
                         const double* betas, const double* precond, double eps, int n_steps, int n, double* x_out, double* p_out);
void launch_eval_points_funnel_var(int cpl, int grid, int block, size_t smem, cudaStream_t s, const Params& P, const double* xs,
                                   const double* betas, int n, double* lp, double* ld, double* grad);
void launch_eval_points_gmm_var(int cpl, int grid, int block, size_t smem, cudaStream_t s, const Params& P, const double* xs,
                                const double* betas, int n, double* lp, double* ld, double* grad);
void launch_leapfrog_funnel_var(int cpl, int grid, int block, size_t smem, cudaStream_t s, const Params& P, const double* xs, const double* ps,
                                const double* betas, const double* precond, double eps, int n_steps, int n, double* x_out, double* p_out);
void launch_leapfrog_gmm_var(int cpl, int grid, int block, size_t smem, cudaStream_t s, const Params& P, const double* xs, const double* ps,
                             const double* betas, const double* precond, double eps, int n_steps, int n, double* x_out, double* p_out);
void launch_eval_points_unid_var(int cpl, int grid, int block, size_t smem, cudaStream_t s, const Params& P, const double* xs,
                                 const double* betas, int n, double* lp, double* ld, double* grad);
void launch_eval_points_unid(int cpl, int grid, int block, size_t smem, cudaStream_t s, const Params& P, const double* xs,
                             const double* betas, int n, double* lp, double* ld, double* grad);
void launch_eval_points_mixed(int cpl, int grid, int block, size_t smem, cudaStream_t s, const Params& P, const double* xs,
                              const double* betas, int n, double* lp, double* ld, double* grad);
void launch_eval_points_mem(int target_kind, int grid, int block, cudaStream_t s, const MemParams& MP, const double* xs,
                            const double* betas, int n, double* lp, double* ld, double* grad);
void launch_init_toy(int grid, int block, cudaStream_t s, const Params& P);
void launch_ising_lp(int grid, int block, cudaStream_t s, const Params& P, const double* xs, const double* betas, int n,
                     double* lp);
void launch_init_recorder_tables(cudaStream_t s, RecEntry* table, size_t n, int n_sms);
void launch_merge_recorders(cudaStream_t s, RecEntry* table, int n_replicas, int n_local, ChainStatsDev* out);
void launch_merge_online(cudaStream_t s, OnEntry* table, int n_replicas, int d, int d_pad, double* mean, double* s2, long long* n_out);
void launch_test_math(int grid, int block, int op, const double* in, double* out, long long n, unsigned int seed_lo,
                      unsigned int seed_hi, unsigned int replica_index);
// logistic regression (pgn_logreg_host.cu)
void logreg_allocate(pgn_handle* h, const pgn_config* cfg);
void logreg_fill_params(pgn_handle* h, LrParams& P);
void logreg_run_round(pgn_handle* h, int64_t n_scans, LrParams& P, std::vector<ChainStatsDev>& st_out, float& total_ms);
void logreg_points(pgn_handle* h, const double* x, int n_points, const double* beta, double* lp, double* ld, double* grad);
double logreg_measure_fp64_peak(int device);
void logreg_test_dmma(const double* a, const double* b, const double* c, double* d_out, int n_trials);

}  // namespace pgn
