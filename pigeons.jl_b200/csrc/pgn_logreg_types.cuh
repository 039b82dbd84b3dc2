// pgn_logreg_types.cuh — plain-data types of the logistic-regression path shared by the host code
// (pgn_host.hpp) and the kernels (pgn_logreg.cuh).
#pragma once
#include "pgn_kernels.cuh"

namespace pgn {

constexpr int LR_TILE = 128;     // rows per log-likelihood tile (canonical tree inside a tile)
constexpr int LR_CHUNK = 4096;   // rows per split-K chunk of the gradient GEMM
constexpr int GEMM_BM = 128, GEMM_BN = 128, GEMM_BK = 16, GEMM_THREADS = 256;
constexpr int DMMA_LD = 132;   // padded row stride (doubles): 264 words = 8 mod 32 -> conflict-free fragment loads

enum LrPhase { LR_SCAN_START = 0, LR_WAIT_X0 = 1, LR_WAIT_TRIAL = 2, LR_DONE = 3 };

struct LrChainState {
  int phase, refresh_i, dir, mode, n, exponent, nst, expo0;
  int pre_mode;   // 0: identity, 1: 1/sd, 2: mix + rmix/sd
  int err;
  double mix, rmix;
  double eps, h_before, init_joint, lower, upper, u_mh;
  double e0, e1, lp0;                 // densities at x
  double f_a0, f_a1, f_lp, h_rev;     // forward proposal
  double t_a0, t_a1, t_lp1, t_h_after, t_eps, pp;   // current trial
  double q_a0, q_a1, q_lp1, q_h_after, q_eps;       // previous trial of a growing search (its candidate one step back)
  // replica
  int replica_index, rt_state;
  unsigned long long ctr;
  // statistics (per chain, whole round)
  MeanAcc expl_acc, am, rev, swap_acc;
  LogSumAcc ls_fwd, ls_bwd;
  long long n_steps, n_points, n_ref, n_restarts, n_trips;
  // swap scratch
  double lr, u;
  int accepted;
};

constexpr int LR_STEP_CHUNK = 8;   // batch steps the host enqueues between two looks at the device-side counters

struct LrControl {   // device-side bookkeeping of the batched evaluation loop
  int n_cols;                   // columns of the current batch step (chains with a pending point)
  int hist[LR_STEP_CHUNK];      // n_cols of every step of the chunk in flight
  long long steps;              // batch steps of this round that evaluated at least one column
  long long sum_active;         // columns requested, summed over the steps
  long long sum_gemm_cols;      // columns multiplied (whole tiles of GEMM_BN), summed over the steps
};

struct LrParams {
  int d, d_pad, n_chains, first_chain, n_local, r_pad;
  int explorer_kind;
  long long scan;
  unsigned int seed_lo, seed_hi, epoch;
  double iv_ref, ls_ref, sigma_ref;
  int n_refresh; double step_size; int precond_kind; double mix_p0, mix_p01;
  const double* std_devs;
  const double* beta;
  LrChainState* st;
  double *X, *P, *G0, *SX, *SP, *SG, *TP, *TG, *FX, *FG, *TX;   // [r_pad][d_pad]
  double *QX, *QP, *QG;   // the previous trial's point, momentum and gradient (grow_step_size steps back to it)
  const double* lik;   // [r_pad]
  const double* G;     // [r_pad][d_pad] likelihood gradient at TX
  int* error_flag;
  // swap / logs
  char* mail; char* mail_left; char* mail_right; unsigned long long slot_bytes;
  double* online_mean; double* online_s2; long long* online_n;
  int* index_process; double* swap_lr; double* swap_u; unsigned char* swap_accept; double* target_trace;
  unsigned long long timeout_ns;
  RecEntry* rec_table; OnEntry* on_table;   // per-replica recorders (null: per chain), see pgn_kernels.cuh
};

}  // namespace pgn
