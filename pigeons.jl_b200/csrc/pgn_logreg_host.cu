// pgn_logreg_host.cu — host side of the logistic-regression path (BASELINE config 5): buffers,
// the batched-evaluation step (two FP64 GEMMs + Bernoulli terms + reductions), the round loop,
// the parity entry points, and the FP64 probes.  Kernels: pgn_logreg.cuh.
#include "pgn_host.hpp"
#include "pgn_logreg.cuh"

namespace pgn {

// ===========================================================================
// logistic regression: batched-GEMM engine (pgn_logreg.cuh)
// ===========================================================================
constexpr size_t GEMM_SMEM_BYTES = 2ull * 2 * GEMM_BK * GEMM_BM * sizeof(double);
constexpr int DMMA_BN = 64;   // column tile of the tensor-core GEMM (two 128 x 64 blocks per SM); PGN_DMMA_BN=128: one 128 x 128 block
constexpr size_t dmma_smem_bytes(int bn) { return 2ull * GEMM_BK * (DMMA_LD + bn + 4) * sizeof(double); }

void logreg_allocate(pgn_handle* h, const pgn_config* cfg) {
  const int d = cfg->dim, dp = h->d_pad;
  const int n = (int)cfg->p[0];
  const int np = (n + 127) / 128 * 128;
  const int rp = (h->n_local + 127) / 128 * 128;
  h->lr_n_data = n; h->lr_n_pad = np; h->lr_r_pad = rp;
  h->lr_splits = (np + LR_CHUNK - 1) / LR_CHUNK;
  // X row-major padded [np][dp] (K-major operand of the gradient GEMM) and its transpose [dp][np]
  {
    std::vector<double> xr((size_t)np * dp, 0.0);
    for (int i = 0; i < n; ++i) std::memcpy(&xr[(size_t)i * dp], cfg->data_x + (size_t)i * d, sizeof(double) * d);
    h->lr_Xr.alloc(xr.size(), false);
    h->lr_Xr.upload(xr.data(), xr.size());
  }
  h->lr_Xt.alloc((size_t)dp * np, false);
  {
    dim3 grid((dp + 31) / 32, (np + 31) / 32), block(32, 8);
    logreg_transpose_kernel<<<grid, block>>>(h->lr_Xr.p, np, dp, dp, h->lr_Xt.p, np);
    CUDA_CHECK(cudaGetLastError());
    CUDA_CHECK(cudaDeviceSynchronize());
  }
  h->lr_y.alloc(np);
  h->lr_y.upload(cfg->data_y, n);
  const size_t vec = (size_t)rp * dp;
  h->lr_Theta.alloc(vec); h->lr_Thetat.alloc(vec);
  h->lr_LL.alloc((size_t)np * rp, false); h->lr_Res.alloc((size_t)np * rp, false);
  h->lr_lik.alloc(rp);
  h->lr_Gp.alloc((size_t)h->lr_splits * dp * rp, false);
  h->lr_G.alloc(vec);
  h->lr_P.alloc(vec); h->lr_G0.alloc(vec); h->lr_SX.alloc(vec); h->lr_SP.alloc(vec); h->lr_SG.alloc(vec);
  h->lr_TP.alloc(vec); h->lr_TG.alloc(vec); h->lr_FX.alloc(vec); h->lr_FG.alloc(vec);
  h->lr_QX.alloc(vec); h->lr_QP.alloc(vec); h->lr_QG.alloc(vec);
  h->lr_st.alloc(h->n_local);
  h->lr_cols.alloc(rp);
  h->lr_ctl.alloc(1);
  {
    const char* g = std::getenv("PGN_GEMM");   // "simt" selects the DFMA kernel; default: FP64 tensor cores
    h->lr_use_dmma = !(g != nullptr && std::string(g) == "simt");
  }
  {
    const char* b = std::getenv("PGN_DMMA_BN");
    h->lr_dmma_bn = (b != nullptr && std::atoi(b) == 128) ? 128 : DMMA_BN;
  }
  CUDA_CHECK(cudaFuncSetAttribute(dgemm_km_dmma_kernel<0, 64>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)dmma_smem_bytes(64)));
  CUDA_CHECK(cudaFuncSetAttribute(dgemm_km_dmma_kernel<0, 128>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)dmma_smem_bytes(128)));
  CUDA_CHECK(cudaFuncSetAttribute(dgemm_km_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)GEMM_SMEM_BYTES));
  CUDA_CHECK(cudaFuncSetAttribute(dgemm_km_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)GEMM_SMEM_BYTES));
}

// column tile of the GEMMs of this handle: the compaction, the gather and the Bernoulli pass round n_cols up to it
static int logreg_col_tile(const pgn_handle* h) { return h->lr_use_dmma ? h->lr_dmma_bn : GEMM_BN; }

// Evaluate likelihood and its gradient at the rows cols[0..n_cols) of `theta` ([r_pad][d_pad]); results land in
// lr_lik[cols[j]] and lr_G[cols[j]][:].  n_cols and cols live on the device (LrControl / lr_cols, written by
// logreg_compact_kernel or by the parity entry point): the host launches the full grids, blocks of column tiles
// beyond n_cols return at once.  The four events bracket the two GEMM launches (not the Bernoulli / reduction passes).
void logreg_eval_batch(pgn_handle* h, const double* theta, cudaEvent_t* ev) {   // ev: null or 4 events bracketing the two GEMMs
  const int dp = h->d_pad, np = h->lr_n_pad, rp = h->lr_r_pad;
  const int* ncp = &h->lr_ctl.p->n_cols;
  const int* cols = h->lr_cols.p;
  const int bn = logreg_col_tile(h);
  {
    dim3 grid((dp + 31) / 32, (rp + 31) / 32), block(32, 8);
    logreg_gather_transpose_kernel<<<grid, block, 0, h->stream>>>(theta, dp, dp, ncp, cols, h->lr_Thetat.p, rp, bn);
  }
  auto gemm = [&](const double* A, int lda, const double* B, int k_total, int k_chunk, double* C, size_t split_stride, dim3 grid) {
    if (!h->lr_use_dmma)
      dgemm_km_kernel<0><<<grid, GEMM_THREADS, GEMM_SMEM_BYTES, h->stream>>>(A, lda, B, rp, k_total, k_chunk, C, nullptr, rp, split_stride,
                                                                             nullptr, 0, ncp);
    else if (bn == 64)
      dgemm_km_dmma_kernel<0, 64><<<grid, 128, dmma_smem_bytes(64), h->stream>>>(A, lda, B, rp, k_total, k_chunk, C, nullptr, rp,
                                                                                 split_stride, nullptr, 0, ncp);
    else
      dgemm_km_dmma_kernel<0, 128><<<grid, 256, dmma_smem_bytes(128), h->stream>>>(A, lda, B, rp, k_total, k_chunk, C, nullptr, rp,
                                                                                   split_stride, nullptr, 0, ncp);
  };
  if (ev) CUDA_CHECK(cudaEventRecord(ev[0], h->stream));
  gemm(h->lr_Xt.p, np, h->lr_Thetat.p, dp, dp, h->lr_LL.p, 0, dim3(np / GEMM_BM, rp / bn, 1));
  if (ev) CUDA_CHECK(cudaEventRecord(ev[1], h->stream));
  logreg_bernoulli_kernel<<<h->n_sms * 8, 256, 0, h->stream>>>(h->lr_LL.p, h->lr_Res.p, h->lr_y.p, rp, np, h->lr_n_data, ncp, bn);
  logreg_reduce_ll_kernel<<<(rp + 7) / 8, 256, 0, h->stream>>>(h->lr_LL.p, rp, h->lr_n_data, ncp, cols, h->lr_lik.p);
  if (ev) CUDA_CHECK(cudaEventRecord(ev[2], h->stream));
  gemm(h->lr_Xr.p, dp, h->lr_Res.p, np, LR_CHUNK, h->lr_Gp.p, (size_t)dp * rp, dim3(dp / GEMM_BM, rp / bn, h->lr_splits));
  if (ev) CUDA_CHECK(cudaEventRecord(ev[3], h->stream));
  {
    dim3 grid((dp + 31) / 32, (rp + 31) / 32), block(32, 8);
    logreg_finalize_grad_kernel<<<grid, block, 0, h->stream>>>(h->lr_Gp.p, h->lr_splits, (size_t)dp * rp, rp, dp, ncp, cols,
                                                              h->lr_G.p);
  }
  CUDA_CHECK(cudaGetLastError());
}
constexpr int LR_EVAL_LAUNCHES = 7;   // gather/transpose, GEMM, Bernoulli, reduce, GEMM, finalize + the compaction before them

void logreg_fill_params(pgn_handle* h, LrParams& P) {
  std::memset(&P, 0, sizeof(P));
  P.d = h->cfg.dim; P.d_pad = h->d_pad; P.n_chains = h->cfg.n_chains; P.first_chain = h->first_chain;
  P.n_local = h->n_local; P.r_pad = h->lr_r_pad;
  P.explorer_kind = h->ep.kind;
  P.seed_lo = (unsigned int)(unsigned long long)h->cfg.seed;
  P.seed_hi = (unsigned int)((unsigned long long)h->cfg.seed >> 32);
  P.epoch = h->epoch;
  P.sigma_ref = h->cfg.p[3]; P.ls_ref = h->cfg.p[4]; P.iv_ref = h->cfg.p[5];
  P.n_refresh = h->ep.n_refresh; P.step_size = h->ep.step_size; P.precond_kind = h->ep.precond_kind;
  P.mix_p0 = h->ep.mix_p0; P.mix_p01 = h->ep.mix_p01;

  P.std_devs = h->have_std ? h->std_devs.p : nullptr;
  P.beta = h->beta.p;
  P.st = h->lr_st.p;
  P.X = h->x.p; P.P = h->lr_P.p; P.G0 = h->lr_G0.p; P.SX = h->lr_SX.p; P.SP = h->lr_SP.p; P.SG = h->lr_SG.p;
  P.TP = h->lr_TP.p; P.TG = h->lr_TG.p; P.FX = h->lr_FX.p; P.FG = h->lr_FG.p; P.TX = h->lr_Theta.p;
  P.QX = h->lr_QX.p; P.QP = h->lr_QP.p; P.QG = h->lr_QG.p;
  P.lik = h->lr_lik.p; P.G = h->lr_G.p;
  P.error_flag = h->error_flag.p;
  P.mail = h->mail.p; P.mail_left = h->mail_left; P.mail_right = h->mail_right; P.slot_bytes = h->slot_bytes;
  P.online_mean = h->online_mean.p; P.online_s2 = h->online_s2.p; P.online_n = h->online_n.p;
  P.timeout_ns = h->timeout_ns * 30ull;   // scans take seconds here; neighbours may lag (default 600 s)
  P.rec_table = h->recorder_order == PGN_RECORDERS_PER_REPLICA ? h->rec_table.p : nullptr;
  P.on_table = h->recorder_order == PGN_RECORDERS_PER_REPLICA ? h->on_table.p : nullptr;
}

// run_one_round! for the logistic-regression target
void logreg_run_round(pgn_handle* h, int64_t n_scans, LrParams& P, std::vector<ChainStatsDev>& st_out, float& total_ms) {
  const int nl = h->n_local;
  if (h->ep.kind != PGN_EXPLORER_AUTOMALA && h->ep.kind != PGN_EXPLORER_MALA)
    throw CudaError{PGN_ERR_INVALID, "LOGREG supports the AutoMALA and MALA explorers"};
  // chain state from the replica arrays
  std::vector<int> ri(nl), rt(nl);
  std::vector<unsigned long long> ctr(nl);
  h->replica_index.download(ri.data(), nl, h->stream);
  h->rng_ctr.download(ctr.data(), nl, h->stream);
  std::vector<LrChainState> st(nl);
  std::memset(st.data(), 0, sizeof(LrChainState) * nl);
  for (int i = 0; i < nl; ++i) {
    st[i].phase = LR_SCAN_START;
    st[i].replica_index = ri[i]; st[i].ctr = ctr[i]; st[i].rt_state = 0;
    st[i].ls_fwd.value = -INFINITY; st[i].ls_bwd.value = -INFINITY;
  }
  h->lr_st.upload(st.data(), nl, h->stream);
  CUDA_CHECK(cudaMemsetAsync(h->online_mean.p, 0, sizeof(double) * h->d_pad, h->stream));
  CUDA_CHECK(cudaMemsetAsync(h->online_s2.p, 0, sizeof(double) * h->d_pad, h->stream));
  CUDA_CHECK(cudaMemsetAsync(h->online_n.p, 0, sizeof(long long), h->stream));
  // The batched evaluation loop.  Per batch step: controller (every chain consumes the evaluation of its pending point
  // and emits the next one) -> compaction of the chains that emitted a point -> the two GEMMs over those columns only.
  // The host enqueues LR_STEP_CHUNK steps at a time and looks at the device-side counters once per chunk: steps after
  // the last chain finished its scan find n_cols = 0 and every kernel of them returns at once.
  cudaEvent_t ge[4 * LR_STEP_CHUNK];
  for (auto& e : ge) CUDA_CHECK(cudaEventCreate(&e));
  h->last_gemm_ms = 0.0;
  h->last_batch_steps = 0;
  h->last_launches = 0;
  h->last_active_cols = 0;
  h->last_gemm_cols = 0;
  CUDA_CHECK(cudaMemsetAsync(h->lr_ctl.p, 0, sizeof(LrControl), h->stream));
  const int wpb = 4, grid = (nl + wpb - 1) / wpb;
  CUDA_CHECK(cudaEventRecord(h->ev0, h->stream));
  int flag = 0;
  LrControl ctl;
  for (int64_t scan = 1; scan <= n_scans && flag == 0; ++scan) {
    P.scan = scan;
    bool scan_done = false;
    while (!scan_done && flag == 0) {
      for (int k = 0; k < LR_STEP_CHUNK; ++k) {
        logreg_controller_kernel<<<grid, wpb * 32, 0, h->stream>>>(P);
        logreg_compact_kernel<<<1, 1024, 0, h->stream>>>(h->lr_st.p, nl, h->lr_cols.p, h->lr_ctl.p, k, logreg_col_tile(h));
        logreg_eval_batch(h, h->lr_Theta.p, &ge[4 * k]);
      }
      h->last_launches += (long long)LR_STEP_CHUNK * (1 + LR_EVAL_LAUNCHES);
      CUDA_CHECK(cudaMemcpyAsync(&ctl, h->lr_ctl.p, sizeof(LrControl), cudaMemcpyDeviceToHost, h->stream));
      CUDA_CHECK(cudaMemcpyAsync(&flag, h->error_flag.p, sizeof(int), cudaMemcpyDeviceToHost, h->stream));
      CUDA_CHECK(cudaStreamSynchronize(h->stream));
      for (int k = 0; k < LR_STEP_CHUNK; ++k) {
        if (ctl.hist[k] == 0) { scan_done = true; continue; }
        float ms1 = 0.f, ms2 = 0.f;
        CUDA_CHECK(cudaEventElapsedTime(&ms1, ge[4 * k], ge[4 * k + 1]));
        CUDA_CHECK(cudaEventElapsedTime(&ms2, ge[4 * k + 2], ge[4 * k + 3]));
        h->last_gemm_ms += ms1 + ms2;
      }
    }
    if (flag != 0) break;
    logreg_post_kernel<<<grid, wpb * 32, 0, h->stream>>>(P);
    logreg_decide_kernel<<<grid, wpb * 32, 0, h->stream>>>(P);
    h->last_launches += 2;
    CUDA_CHECK(cudaMemcpyAsync(&flag, h->error_flag.p, sizeof(int), cudaMemcpyDeviceToHost, h->stream));
    CUDA_CHECK(cudaStreamSynchronize(h->stream));
  }
  if (P.rec_table != nullptr && flag == 0 && n_scans > 0) {
    logreg_flush_recorders_kernel<<<grid, wpb * 32, 0, h->stream>>>(P);
    h->last_launches += 1;
  }
  CUDA_CHECK(cudaEventRecord(h->ev1, h->stream));
  CUDA_CHECK(cudaStreamSynchronize(h->stream));
  CUDA_CHECK(cudaEventElapsedTime(&total_ms, h->ev0, h->ev1));
  for (auto& e : ge) cudaEventDestroy(e);
  h->lr_ctl.download(&ctl, 1, h->stream);
  h->last_batch_steps = ctl.steps;
  h->last_active_cols = ctl.sum_active;
  h->last_gemm_cols = ctl.sum_gemm_cols;
  // replica arrays + statistics back
  h->lr_st.download(st.data(), nl, h->stream);
  st_out.assign(nl, ChainStatsDev{});
  for (int i = 0; i < nl; ++i) {
    const LrChainState& s = st[i];
    ri[i] = s.replica_index; ctr[i] = s.ctr; rt[i] = s.rt_state;
    ChainStatsDev& o = st_out[i];
    o.swap_n = s.swap_acc.n; o.swap_mean = s.swap_acc.mu; o.ls_fwd = s.ls_fwd.value; o.ls_bwd = s.ls_bwd.value;
    o.expl_acc_n = s.expl_acc.n; o.expl_acc_mean = s.expl_acc.mu; o.n_steps = s.n_steps;
    o.am_n = s.am.n; o.am_mean = s.am.mu; o.rev_n = s.rev.n; o.rev_mean = s.rev.mu;
    o.n_restarts = s.n_restarts; o.n_round_trips = s.n_trips; o.n_points = s.n_points; o.n_ref_evals = s.n_ref;
  }
  h->replica_index.upload(ri.data(), nl, h->stream);
  h->rng_ctr.upload(ctr.data(), nl, h->stream);
  h->rt_state.upload(rt.data(), nl, h->stream);
}

// parity entry points for LOGREG: batches of r_pad points through the same GEMM path
void logreg_points(pgn_handle* h, const double* x, int n_points, const double* beta, double* lp, double* ld, double* grad) {
  const int d = h->cfg.dim, dp = h->d_pad, rp = h->lr_r_pad;
  DevBuf<double> db, dlp, dld, dg;
  db.alloc(rp); dlp.alloc(rp); dld.alloc(rp); dg.alloc((size_t)rp * d);
  std::vector<double> stage((size_t)rp * dp);
  for (int base = 0; base < n_points; base += rp) {
    const int m = std::min(rp, n_points - base);
    std::fill(stage.begin(), stage.end(), 0.0);
    for (int i = 0; i < m; ++i) std::memcpy(&stage[(size_t)i * dp], x + (size_t)(base + i) * d, sizeof(double) * d);
    h->lr_Theta.upload(stage.data(), stage.size());
    db.upload(beta + base, m);
    {   // the points are the columns 0..m-1
      std::vector<int> ident(rp);
      for (int i = 0; i < rp; ++i) ident[i] = i;
      h->lr_cols.upload(ident.data(), rp);
      LrControl c;
      std::memset(&c, 0, sizeof(c));
      c.n_cols = m;
      CUDA_CHECK(cudaMemcpy(h->lr_ctl.p, &c, sizeof(c), cudaMemcpyHostToDevice));
    }
    logreg_eval_batch(h, h->lr_Theta.p, nullptr);
    logreg_points_finish_kernel<<<(m + 3) / 4, 128, 0, h->stream>>>(h->lr_Theta.p, d, dp, m, db.p, h->lr_lik.p, h->lr_G.p,
                                                                   h->cfg.p[5], h->cfg.p[4], lp ? dlp.p : nullptr,
                                                                   ld ? dld.p : nullptr, dg.p);
    CUDA_CHECK(cudaGetLastError());
    CUDA_CHECK(cudaStreamSynchronize(h->stream));
    if (lp) dlp.download(lp + base, m);
    if (ld) { dld.download(ld + base, m); dg.download(grad + (size_t)base * d, (size_t)m * d); }
  }
}

double logreg_measure_fp64_peak(int device) {
  CUDA_CHECK(cudaSetDevice(device));
  cudaDeviceProp prop;
  CUDA_CHECK(cudaGetDeviceProperties(&prop, device));
  DevBuf<double> sink;
  sink.alloc(1);
  const int iters = 4096, blocks = prop.multiProcessorCount * 4, threads = 256;
  cudaEvent_t a, b;
  CUDA_CHECK(cudaEventCreate(&a));
  CUDA_CHECK(cudaEventCreate(&b));
  fp64_peak_kernel<<<blocks, threads>>>(sink.p, iters, 1.0000001);
  CUDA_CHECK(cudaDeviceSynchronize());
  float best = 1e30f;
  for (int rep = 0; rep < 5; ++rep) {
    CUDA_CHECK(cudaEventRecord(a));
    fp64_peak_kernel<<<blocks, threads>>>(sink.p, iters, 1.0000001);
    CUDA_CHECK(cudaEventRecord(b));
    CUDA_CHECK(cudaEventSynchronize(b));
    float ms = 0.f;
    CUDA_CHECK(cudaEventElapsedTime(&ms, a, b));
    best = ms < best ? ms : best;
  }
  cudaEventDestroy(a);
  cudaEventDestroy(b);
  const double flops = 2.0 * 16.0 * (double)iters * (double)blocks * (double)threads;
  return flops / (best * 1e-3) / 1e12;
}

void logreg_test_dmma(const double* a, const double* b, const double* c, double* d_out, int n_trials) {
  DevBuf<double> da, db, dc, dd;
  da.alloc((size_t)n_trials * 32, false); db.alloc((size_t)n_trials * 32, false);
  dc.alloc((size_t)n_trials * 64, false); dd.alloc((size_t)n_trials * 64, false);
  da.upload(a, (size_t)n_trials * 32); db.upload(b, (size_t)n_trials * 32); dc.upload(c, (size_t)n_trials * 64);
  dmma_probe_kernel<<<(n_trials + 3) / 4, 128>>>(da.p, db.p, dc.p, dd.p, n_trials);
  CUDA_CHECK(cudaGetLastError());
  CUDA_CHECK(cudaDeviceSynchronize());
  dd.download(d_out, (size_t)n_trials * 64);
}

}  // namespace pgn
