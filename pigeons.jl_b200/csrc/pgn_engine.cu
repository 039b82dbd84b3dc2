// pgn_engine.cu — the C ABI declared in include/pigeons_b200.h: handle lifecycle, kernel
// selection and launch geometry.
//
// One handle = one shard of the chain ladder on one GPU.  `pgn_run_round`
// is the single call per round that replaces the reference's host scan loop
// (src/pt/pigeons.jl:46-55).  The kernels live in the other translation units
// (pgn_scan_vec.cu, pgn_scan_misc.cu, pgn_scan_mem.cu, pgn_logreg_host.cu; see pgn_host.hpp).
#include "pgn_host.hpp"

using namespace pgn;

namespace {

int fail(char** err, int code, const std::string& msg) {
  if (err) {
    *err = (char*)std::malloc(msg.size() + 1);
    std::memcpy(*err, msg.c_str(), msg.size() + 1);
  }
  return code;
}

// LoadBalance (src/mpi_utils/LoadBalance.jl:70-73,119-128), 0-based rank
void balanced_range(int n_chains, int world, int rank, int& first_chain, int& n_local) {
  const int basic = n_chains / world, extras = n_chains % world;
  n_local = basic + (rank < extras ? 1 : 0);
  const int with_extra = std::min(rank, extras);
  first_chain = 1 + (rank - with_extra) * basic + with_extra * (basic + 1);
}
// Two legs (0 < n_var < n_chains): the two target chains n_var and n_var + 1 stay on one shard — when a boundary of the
// balanced split falls between them, chain n_var + 1 moves to the lower shard (same rule in distributed.py:shard_layout).
// Returns false when that would empty the upper shard.
bool shard_range(int n_chains, int world, int rank, int n_var, int& first_chain, int& n_local) {
  balanced_range(n_chains, world, rank, first_chain, n_local);
  if (!(n_var > 0 && n_var < n_chains) || world < 2) return true;
  for (int r = 0; r + 1 < world; ++r) {
    int f, n;
    balanced_range(n_chains, world, r, f, n);
    if (f + n - 1 != n_var) continue;
    int f2, n2;
    balanced_range(n_chains, world, r + 1, f2, n2);
    if (n2 < 2) return false;
    if (rank == r) n_local += 1;
    if (rank == r + 1) { first_chain += 1; n_local -= 1; }
  }
  return true;
}

void fill_params(pgn_handle* h, Params& P) {
  std::memset(&P, 0, sizeof(P));
  P.target_kind = h->cfg.target_kind;
  P.d = h->cfg.dim; P.d_pad = h->d_pad; P.n_chains = h->cfg.n_chains;
  P.first_chain = h->first_chain; P.n_local = h->n_local;
  P.seed_lo = (unsigned int)(unsigned long long)h->cfg.seed;
  P.seed_hi = (unsigned int)((unsigned long long)h->cfg.seed >> 32);
  P.epoch = h->epoch;
  P.tag_base = (unsigned int)h->scan_seq;
  for (int i = 0; i < 8; ++i) P.p[i] = h->cfg.p[i];
  P.n_modes = h->cfg.n_modes;
  P.means = h->means.p; P.log_w = h->log_w.p; P.beta = h->beta.p;
  P.slice_w = h->ep.slice_w; P.slice_p = h->ep.slice_p; P.slice_n_passes = h->ep.slice_n_passes;
  P.slice_max_iter = h->ep.slice_max_iter;
  P.n_refresh = h->ep.n_refresh; P.step_size = h->ep.step_size; P.precond_kind = h->ep.precond_kind;
  P.mix_p0 = h->ep.mix_p0; P.mix_p01 = h->ep.mix_p01;
  P.n_steps = h->ep.n_steps; P.program_is_mix = h->ep.kind == PGN_EXPLORER_MIX ? 1 : 0;
  for (int v = 0; v < PGN_MAX_MIX; ++v) P.step_kind[v] = h->ep.step_kind[v];
  P.n_mix = h->ep.n_mix;
  for (int v = 0; v < PGN_MAX_MIX; ++v) {
    P.mix_n_refresh[v] = h->ep.mix_n_refresh[v]; P.mix_precond_kind[v] = h->ep.mix_precond_kind[v];
    P.mix_step_size[v] = h->ep.mix_step_size[v]; P.mix_variant_p0[v] = h->ep.mix_variant_p0[v];
    P.mix_variant_p01[v] = h->ep.mix_variant_p01[v];
  }
  P.std_devs = h->have_std ? h->std_devs.p : nullptr;
  P.ising_n_steps = h->ep.ising_n_steps;
  P.x = h->x.p; P.replica_index = h->replica_index.p; P.rng_ctr = h->rng_ctr.p; P.rt_state = h->rt_state.p;
  P.mail = h->mail.p; P.mail_left = h->mail_left; P.mail_right = h->mail_right;
  P.slot_bytes = h->slot_bytes;
  P.stats = h->stats.p;
  P.online_mean = h->online_mean.p; P.online_s2 = h->online_s2.p; P.online_n = h->online_n.p;
  P.error_flag = h->error_flag.p;
  P.timeout_ns = h->timeout_ns;
  P.progress = h->progress.p;
  const bool per_replica = h->recorder_order == PGN_RECORDERS_PER_REPLICA;
  P.rec_table = per_replica ? h->rec_table.p : nullptr;
  P.on_table = per_replica ? h->on_table.p : nullptr;
  P.n_var = h->cfg.n_chains_variational;
  P.var_tab = h->var_active ? h->var_tab.p : nullptr;
}

// does this shard own the target chain(s)?  One leg: chain N; two legs: chains n_var and n_var + 1 (one shard, pgn_create)
bool shard_owns_target(const pgn_handle* h) {
  const int nv = h->cfg.n_chains_variational, N = h->cfg.n_chains;
  const int t = (nv > 0 && nv < N) ? nv : N;
  return t >= h->first_chain && t < h->first_chain + h->n_local;
}

constexpr bool kMixedTeamsDefault = true;    // "mixed teams" launch of single-warp autoMALA ladders (pgn_run_round); PGN_MIXED_TEAMS overrides

void* select_scan_kernel(const pgn_handle* h) {
  const int ex = h->ep.kind;
  switch (h->cfg.target_kind) {
    case PGN_TARGET_TOY_MVN: return vec_scan_kernel_toy(h->cpl, ex);
    case PGN_TARGET_FUNNEL: return h->var_active ? vec_scan_kernel_funnel_var(h->cpl, ex) : vec_scan_kernel_funnel(h->cpl, ex);
    case PGN_TARGET_GMM: return h->var_active ? vec_scan_kernel_gmm_var(h->cpl, ex) : vec_scan_kernel_gmm(h->cpl, ex);
    case PGN_TARGET_MIXED: return vec_scan_kernel_mixed(h->cpl, ex);
    case PGN_TARGET_UNID: return h->var_active ? vec_scan_kernel_unid_var(h->cpl, ex) : vec_scan_kernel_unid(h->cpl, ex);
    case PGN_TARGET_ISING: return ex == PGN_EXPLORER_ISING_METROPOLIS ? ising_scan_kernel() : nullptr;
    case PGN_TARGET_TEST_SWAPPER: return ex == PGN_EXPLORER_NONE ? test_swapper_scan_kernel() : nullptr;
    default: return nullptr;
  }
}
void* select_mem_kernel(const pgn_handle* h) { return mem_scan_kernel(h->cfg.target_kind, h->ep.kind); }
void mem_allocate(pgn_handle* h) {
  if (h->mem_allocated) return;
  h->mem_rec.alloc_on(h->n_local, h->stream);
  for (int i = 0; i < 10; ++i) h->mem_vec[i].alloc_on((size_t)h->n_local * h->d_pad, h->stream);
  h->mem_allocated = true;
}
void mem_fill_params(pgn_handle* h, const Params& P, MemParams& MP) {
  MP.base = P;
  MP.rec = h->mem_rec.p;
  MP.VP = h->mem_vec[0].p; MP.VG0 = h->mem_vec[1].p; MP.VSX = h->mem_vec[2].p; MP.VSP = h->mem_vec[3].p;
  MP.VSG = h->mem_vec[4].p; MP.VTX = h->mem_vec[5].p; MP.VTP = h->mem_vec[6].p; MP.VTG = h->mem_vec[7].p;
  MP.VFX = h->mem_vec[8].p; MP.VFG = h->mem_vec[9].p;
  MP.nslots = h->d_pad / 32;
}

bool is_team_kernel(const pgn_handle* h) {   // VecChain<.., AUTOMALA>::kTeam
  const int tk = h->cfg.target_kind;
  return (h->ep.kind == PGN_EXPLORER_AUTOMALA || h->ep.kind == PGN_EXPLORER_COMPOSE || h->ep.kind == PGN_EXPLORER_MIX) && h->cpl > 0 &&
         (tk == PGN_TARGET_TOY_MVN || tk == PGN_TARGET_FUNNEL || tk == PGN_TARGET_GMM);
}
// dynamic shared memory of the scan kernel: staged target constants, then (team kernels) the
// control words and 3 rotating buffers of W trial slots
size_t scan_smem_bytes(const pgn_handle* h, int team_w = 0, int warps_per_block = 1, int pool_refresh = 0) {
  size_t n = 0;
  if (h->cfg.target_kind == PGN_TARGET_GMM) n += (size_t)KMAX_MODES * h->d_pad + KMAX_MODES;
  if (h->cfg.target_kind == PGN_TARGET_ISING)   // one Metropolis-ratio table per chain (IsingChain::build_table)
    n += (size_t)warps_per_block * IsingChain::table_doubles((int)h->cfg.p[1]);
  if (team_w > 0) n += (24 + h->cpl * 32) + (size_t)3 * team_w * (3 * h->cpl * 32 + 8);   // VecChain::TEAM_CTL_DOUBLES + slots
  if (team_w > 1) n += (size_t)pool_refresh * (h->cpl * 32 + 8);                          // VecChain::POOL_DOUBLES per refreshment
  return n * sizeof(double);
}

void launch_eval_points(pgn_handle* h, const Params& P, const double* xs, const double* betas, int n, double* lp,
                        double* ld, double* grad) {
  const int wpb = 4;
  const int grid = (n + wpb - 1) / wpb;
  const size_t smem = scan_smem_bytes(h);
  const int tk = h->cfg.target_kind;
  if (h->cpl == 0) {   // d > 128: memory-resident evaluation; xs / grad are padded [n][d_pad] here
    MemParams MP;
    std::memset(&MP, 0, sizeof(MP));
    MP.base = P;
    MP.nslots = h->d_pad / 32;
    launch_eval_points_mem(tk, grid, wpb * 32, h->stream, MP, xs, betas, n, lp, ld, grad);
    return;
  }
  switch (tk) {
    case PGN_TARGET_TOY_MVN: launch_eval_points_toy(h->cpl, grid, wpb * 32, smem, h->stream, P, xs, betas, n, lp, ld, grad); break;
    case PGN_TARGET_FUNNEL:
      if (h->var_active) launch_eval_points_funnel_var(h->cpl, grid, wpb * 32, smem, h->stream, P, xs, betas, n, lp, ld, grad);
      else launch_eval_points_funnel(h->cpl, grid, wpb * 32, smem, h->stream, P, xs, betas, n, lp, ld, grad);
      break;
    case PGN_TARGET_GMM:
      if (h->var_active) launch_eval_points_gmm_var(h->cpl, grid, wpb * 32, smem, h->stream, P, xs, betas, n, lp, ld, grad);
      else launch_eval_points_gmm(h->cpl, grid, wpb * 32, smem, h->stream, P, xs, betas, n, lp, ld, grad);
      break;
    case PGN_TARGET_MIXED: launch_eval_points_mixed(h->cpl, grid, wpb * 32, smem, h->stream, P, xs, betas, n, lp, ld, grad); break;
    case PGN_TARGET_UNID:
      if (h->var_active) launch_eval_points_unid_var(h->cpl, grid, wpb * 32, smem, h->stream, P, xs, betas, n, lp, ld, grad);
      else launch_eval_points_unid(h->cpl, grid, wpb * 32, smem, h->stream, P, xs, betas, n, lp, ld, grad);
      break;
    default: throw CudaError{PGN_ERR_INVALID, "unsupported target"};
  }
}

void use_device(const pgn_handle* h) { CUDA_CHECK(cudaSetDevice(h->cfg.device)); }

}  // namespace

extern "C" {

int pgn_abi_version(void) { return PGN_ABI_VERSION; }

void pgn_free_string(char* s) { std::free(s); }

int pgn_device_info(int device, pgn_device_info_t* out, char** err) {
  int count = 0;
  if (cudaGetDeviceCount(&count) != cudaSuccess || count == 0 || device >= count)
    return fail(err, PGN_ERR_NO_DEVICE, "no usable CUDA device (this library has no CPU fallback)");
  cudaDeviceProp prop;
  if (cudaGetDeviceProperties(&prop, device) != cudaSuccess) return fail(err, PGN_ERR_CUDA, "cudaGetDeviceProperties failed");
  std::memset(out, 0, sizeof(*out));
  out->sm_major = prop.major; out->sm_minor = prop.minor; out->n_sms = prop.multiProcessorCount;
  out->global_mem_bytes = (int64_t)prop.totalGlobalMem;
  out->max_resident_chains = prop.multiProcessorCount * (prop.maxThreadsPerMultiProcessor / 32);
  std::strncpy(out->name, prop.name, sizeof(out->name) - 1);
  return PGN_OK;
}

int pgn_create(const pgn_config* cfg, pgn_handle** out, char** err) {
  if (!cfg || !out) return fail(err, PGN_ERR_INVALID, "null argument");
  if (cfg->abi_version != PGN_ABI_VERSION) return fail(err, PGN_ERR_INVALID, "ABI version mismatch");
  if (cfg->n_chains < 1) return fail(err, PGN_ERR_INVALID, "n_chains must be >= 1");
  if (cfg->world_size < 1 || cfg->rank < 0 || cfg->rank >= cfg->world_size || cfg->world_size > cfg->n_chains)
    return fail(err, PGN_ERR_INVALID, "need 0 <= rank < world_size <= n_chains");
  switch (cfg->target_kind) {
    case PGN_TARGET_TOY_MVN: case PGN_TARGET_FUNNEL: case PGN_TARGET_GMM:
      if (cfg->dim < 1) return fail(err, PGN_ERR_INVALID, "dim must be >= 1");
      if (cfg->dim > (1 << 20)) return fail(err, PGN_ERR_INVALID, "dim too large");
      break;
    case PGN_TARGET_ISING: {
      const int L = (int)cfg->p[1];
      if (L < 2 || L > 32 || L * L != cfg->dim) return fail(err, PGN_ERR_INVALID, "ISING: 2 <= L <= 32 and dim == L*L");
      break;
    }
    case PGN_TARGET_TEST_SWAPPER: break;
    case PGN_TARGET_UNID:
      if (cfg->dim != 2 || !(cfg->p[0] >= cfg->p[1]) || cfg->p[1] < 0)
        return fail(err, PGN_ERR_INVALID, "UNID: dim == 2 and 0 <= n_successes <= n_trials");
      break;
    case PGN_TARGET_MIXED:
      if (cfg->dim < 1 || cfg->dim > 128) return fail(err, PGN_ERR_INVALID, "MIXED: 1 <= dim <= 128 (register-resident kernels only)");
      if (!cfg->means || cfg->p[2] < 1 || cfg->n_modes != 10 + (int)cfg->p[2] + 1 || cfg->p[0] < 0 || cfg->p[1] < 0 ||
          cfg->p[0] + cfg->p[1] > cfg->dim)
        return fail(err, PGN_ERR_INVALID, "MIXED: parameter table / coordinate counts inconsistent");
      break;
    case PGN_TARGET_LOGREG:
      if (cfg->dim < 1 || cfg->p[0] < 1 || !cfg->data_x || !cfg->data_y)
        return fail(err, PGN_ERR_INVALID, "LOGREG: dim, n_data, data_x, data_y required");
      break;
    default: return fail(err, PGN_ERR_INVALID, "unknown target_kind (device targets are a closed family)");
  }
  if (cfg->target_kind == PGN_TARGET_GMM) {
    if (cfg->n_modes < 1 || cfg->n_modes > KMAX_MODES) return fail(err, PGN_ERR_INVALID, "GMM: 1 <= n_modes <= 8");
    if (!cfg->means || !cfg->log_weights) return fail(err, PGN_ERR_INVALID, "GMM: means / log_weights missing");
  }
  int count = 0;
  if (cudaGetDeviceCount(&count) != cudaSuccess || count == 0)
    return fail(err, PGN_ERR_NO_DEVICE, "no usable CUDA device (this library has no CPU fallback)");
  if (cfg->device < 0 || cfg->device >= count) return fail(err, PGN_ERR_NO_DEVICE, "device ordinal out of range");

  pgn_handle* h = new pgn_handle();
  try {
    h->cfg = *cfg;
    h->cfg.means = h->cfg.log_weights = h->cfg.data_x = h->cfg.data_y = nullptr;
    use_device(h);
    cudaDeviceProp prop;
    CUDA_CHECK(cudaGetDeviceProperties(&prop, cfg->device));
    if (!prop.cooperativeLaunch) throw CudaError{PGN_ERR_NO_DEVICE, "device lacks cooperative launch"};
    h->n_sms = prop.multiProcessorCount;
    if (!shard_range(cfg->n_chains, cfg->world_size, cfg->rank, cfg->n_chains_variational, h->first_chain, h->n_local))
      throw CudaError{PGN_ERR_INVALID, "two legs: too few chains per shard to keep both target chains on one shard"};
    const int d = cfg->dim;
    if (cfg->target_kind == PGN_TARGET_ISING) { h->cpl = 1; h->d_pad = 64; h->pay_doubles = 32; }
    else if (cfg->target_kind == PGN_TARGET_TEST_SWAPPER) { h->cpl = 1; h->d_pad = 1; h->pay_doubles = 0; }
    else if (cfg->target_kind == PGN_TARGET_LOGREG) {
      h->cpl = 0;
      h->d_pad = (d + 127) / 128 * 128;
      h->pay_doubles = h->d_pad + 2;
    }
    else if (d <= 128) {
      h->cpl = d <= 32 ? 1 : (d <= 64 ? 2 : 4);
      h->d_pad = h->cpl * 32;
      h->pay_doubles = h->d_pad;
    } else {   // memory-resident chains only
      h->cpl = 0;
      h->d_pad = (d + 31) / 32 * 32;
      h->pay_doubles = h->d_pad;
    }
    {
      const char* fm = std::getenv("PGN_FORCE_MEM");
      h->force_mem = fm != nullptr && std::string(fm) == "1";
      const char* to = std::getenv("PGN_TIMEOUT_S");      // hand-shake spin limit (default 20 s; the LOGREG path uses 30x)
      if (to && std::atof(to) > 0) h->timeout_ns = (unsigned long long)(std::atof(to) * 1e9);
    }
    {
      const int nv = cfg->n_chains_variational, N = cfg->n_chains;
      if (nv < 0 || nv > N) throw CudaError{PGN_ERR_INVALID, "0 <= n_chains_variational <= n_chains"};
      if (nv > 0 && nv < N) {   // two legs
        if (cfg->recorder_order != PGN_RECORDERS_PER_REPLICA)
          throw CudaError{PGN_ERR_INVALID, "two legs need recorder_order = PGN_RECORDERS_PER_REPLICA"};
        if (cfg->target_kind == PGN_TARGET_LOGREG || (h->cpl == 0 && cfg->target_kind != PGN_TARGET_ISING &&
                                                     cfg->target_kind != PGN_TARGET_TEST_SWAPPER) || h->force_mem)
          throw CudaError{PGN_ERR_INVALID, "two legs run on the register-resident scan kernels only (d <= 128)"};
      }
    }
    if (cfg->recorder_order != PGN_RECORDERS_PER_REPLICA && cfg->recorder_order != PGN_RECORDERS_PER_CHAIN)
      throw CudaError{PGN_ERR_INVALID, "recorder_order must be PGN_RECORDERS_PER_REPLICA (0) or PGN_RECORDERS_PER_CHAIN (1)"};
    h->recorder_order = cfg->recorder_order;
    // the scan kernel's flag-in-data words need 16 bytes per payload double (8 header words + 2 words per double);
    // the logistic-regression and memory-resident kernels use the first 64 + 8 * pay_doubles bytes of a slot
    h->slot_bytes = (size_t)MAIL_HDR_BYTES + (size_t)h->pay_doubles * 2 * sizeof(double);
    h->slot_bytes = (h->slot_bytes + 127) / 128 * 128;
    h->mail_bytes = (size_t)(h->n_local + 2) * MAIL_RINGS * h->slot_bytes;
    const int nl = h->n_local;
    h->beta.alloc(cfg->n_chains);
    h->x.alloc((size_t)nl * h->d_pad);
    h->replica_index.alloc(nl); h->rt_state.alloc(nl); h->rng_ctr.alloc(nl);
    h->stats.alloc(nl);
    h->error_flag.alloc(1);
    h->progress.alloc(1);
    h->online_mean.alloc(h->d_pad); h->online_s2.alloc(h->d_pad); h->online_n.alloc(1);
    h->mail.alloc(h->mail_bytes);
    h->std_devs.alloc(std::max(d, 1));
    // GaussianReference tables: allocated here, not in pgn_set_variational — a device allocation between rounds could wait
    // for another handle's running scan kernel on the same device, which in turn waits for this handle
    if (cfg->n_chains_variational > 0) h->var_tab.alloc((size_t)5 * h->d_pad);
    if (h->recorder_order == PGN_RECORDERS_PER_REPLICA) {
      // every replica's recorder entry for every local chain, and (for vector states) its target-chain online statistics
      const size_t rec_bytes = (size_t)cfg->n_chains * (size_t)std::max(nl, 1) * sizeof(RecEntry);
      const size_t on_bytes = (size_t)cfg->n_chains * (size_t)h->d_pad * sizeof(OnEntry);
      if (rec_bytes + on_bytes > (size_t)48 << 30)
        throw CudaError{PGN_ERR_INVALID, "per-replica recorders need n_chains x n_local entries (> 48 GiB here): use "
                                         "recorder_order = PGN_RECORDERS_PER_CHAIN"};
      h->rec_table.alloc((size_t)cfg->n_chains * std::max(nl, 1), false);
      h->on_table.alloc((size_t)cfg->n_chains * h->d_pad, false);
    }
    if (cfg->target_kind == PGN_TARGET_GMM) {
      // staged layout: [KMAX_MODES][d_pad] means (zero rows beyond K) + KMAX_MODES log weights (-inf beyond K)
      std::vector<double> padded((size_t)KMAX_MODES * h->d_pad + KMAX_MODES, 0.0);
      for (int k = 0; k < cfg->n_modes; ++k)
        for (int c = 0; c < d; ++c) padded[(size_t)k * h->d_pad + c] = cfg->means[(size_t)k * d + c];
      for (int k = 0; k < KMAX_MODES; ++k)
        padded[(size_t)KMAX_MODES * h->d_pad + k] = k < cfg->n_modes ? cfg->log_weights[k] : -INFINITY;
      h->means.alloc(padded.size());
      h->means.upload(padded.data(), padded.size());
      h->log_w.alloc(cfg->n_modes);
      h->log_w.upload(cfg->log_weights, cfg->n_modes);
    }
    if (cfg->target_kind == PGN_TARGET_MIXED) {   // [log p0, log(1-p0), ..., p0, q0, log C(n, 0..n)]
      h->means.alloc(cfg->n_modes);
      h->means.upload(cfg->means, cfg->n_modes);
    }
    if (cfg->target_kind == PGN_TARGET_LOGREG) logreg_allocate(h, cfg);
    // default schedule: equally spaced (src/schedules/Schedule.jl:36-44)
    std::vector<double> b(cfg->n_chains);
    for (int i = 0; i < cfg->n_chains; ++i) b[i] = cfg->n_chains == 1 ? 1.0 : (double)i / (cfg->n_chains - 1);
    h->beta.upload(b.data(), b.size());
    h->ep = pgn_explorer_params{};
    h->ep.kind = PGN_EXPLORER_NONE;
    CUDA_CHECK(cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking));
    {   // per-round scratch comes from the stream-ordered pool; keep what it has freed instead of returning it to the OS
      cudaMemPool_t pool;
      if (cudaDeviceGetDefaultMemPool(&pool, cfg->device) == cudaSuccess) {
        unsigned long long keep = ~0ull;
        cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep);
      }
      cudaGetLastError();
    }
    CUDA_CHECK(cudaEventCreate(&h->ev0));
    CUDA_CHECK(cudaEventCreate(&h->ev1));
  } catch (CudaError& e) {
    delete h;
    return fail(err, e.code, e.msg);
  }
  *out = h;
  return PGN_OK;
}

int pgn_destroy(pgn_handle* h) {
  if (!h) return PGN_OK;
  cudaSetDevice(h->cfg.device);
  if (h->left_is_ipc && h->mail_left) cudaIpcCloseMemHandle(h->mail_left);
  if (h->right_is_ipc && h->mail_right) cudaIpcCloseMemHandle(h->mail_right);
  if (h->ev0) cudaEventDestroy(h->ev0);
  if (h->ev1) cudaEventDestroy(h->ev1);
  h->mem_rec.release();                       // stream-ordered allocations go back while their stream still exists
  for (auto& v : h->mem_vec) v.release();
  if (h->stream) { cudaStreamSynchronize(h->stream); cudaStreamDestroy(h->stream); }
  delete h;
  return PGN_OK;
}

int pgn_local_range(const pgn_handle* h, int32_t* first_chain, int32_t* n_local) {
  *first_chain = h->first_chain;
  *n_local = h->n_local;
  return PGN_OK;
}

int pgn_set_schedule(pgn_handle* h, const double* beta, int32_t n, char** err) {
  if (n != h->cfg.n_chains) return fail(err, PGN_ERR_INVALID, "schedule length != n_chains");
  try {
    use_device(h);
    h->beta.upload(beta, n, h->stream);
  } catch (CudaError& e) { return fail(err, e.code, e.msg); }
  return PGN_OK;
}

int pgn_set_explorer(pgn_handle* h, const pgn_explorer_params* ep, char** err) {
  if (!h || !ep) return fail(err, PGN_ERR_INVALID, "null argument");
  if (ep->n_mix < 0 || ep->n_mix > PGN_MAX_MIX) return fail(err, PGN_ERR_INVALID, "n_mix out of range");
  if (ep->n_mix > 1 && ep->kind != PGN_EXPLORER_AUTOMALA)
    return fail(err, PGN_ERR_INVALID, "the device mixes autoMALA kernels only (no CPU fallback for other mixtures)");
  if (h->cfg.target_kind == PGN_TARGET_UNID && ep->kind != PGN_EXPLORER_SLICE)
    return fail(err, PGN_ERR_INVALID, "UNID: SliceSampler only");
  const bool program = ep->kind == PGN_EXPLORER_COMPOSE || ep->kind == PGN_EXPLORER_MIX;
  if ((ep->n_mix > 1 || program) && (h->cfg.target_kind == PGN_TARGET_LOGREG || h->cpl == 0 || h->force_mem))
    return fail(err, PGN_ERR_INVALID, "Mix / Compose explorers run on the register-resident scan kernels only (d <= 128)");
  if (program) {
    const int tk = h->cfg.target_kind;
    if (tk != PGN_TARGET_TOY_MVN && tk != PGN_TARGET_FUNNEL && tk != PGN_TARGET_GMM)
      return fail(err, PGN_ERR_INVALID, "Mix / Compose explorers need a vector target with a gradient (toy MVN, funnel, mixture)");
    if (ep->n_steps < 1 || ep->n_steps > PGN_MAX_MIX) return fail(err, PGN_ERR_INVALID, "Compose / Mix: n_steps out of range");
    for (int s = 0; s < ep->n_steps; ++s) {
      const int k = ep->step_kind[s];
      if (k != PGN_EXPLORER_SLICE && k != PGN_EXPLORER_AUTOMALA && k != PGN_EXPLORER_MALA &&
          !(k == PGN_EXPLORER_TOY && tk == PGN_TARGET_TOY_MVN))
        return fail(err, PGN_ERR_INVALID, "Compose / Mix: explorers on the device are ToyExplorer (toy MVN), SliceSampler, MALA, AutoMALA "
                                          "(no CPU fallback for others)");
      if ((k == PGN_EXPLORER_AUTOMALA || k == PGN_EXPLORER_MALA) && !(ep->mix_step_size[s] > 0))
        return fail(err, PGN_ERR_INVALID, "Compose / Mix: step size of a gradient-based step must be positive");
    }
  }
  try {
    use_device(h);
    h->ep = *ep;
    h->have_std = ep->std_devs != nullptr;
    if (h->have_std) h->std_devs.upload(ep->std_devs, h->cfg.dim, h->stream);
    h->ep.std_devs = nullptr;
  } catch (CudaError& e) { return fail(err, e.code, e.msg); }
  return PGN_OK;
}

int pgn_init_replicas(pgn_handle* h, char** err) {
  try {
    use_device(h);
    const int nl = h->n_local;
    std::vector<int> ri(nl), rt(nl, 0);
    std::vector<unsigned long long> ctr(nl, 0ull);
    for (int i = 0; i < nl; ++i) ri[i] = h->first_chain + i;
    h->replica_index.upload(ri.data(), nl, h->stream);
    h->rt_state.upload(rt.data(), nl, h->stream);
    h->rng_ctr.upload(ctr.data(), nl, h->stream);
    CUDA_CHECK(cudaMemsetAsync(h->x.p, 0, std::max<size_t>(1, (size_t)nl * h->d_pad) * sizeof(double), h->stream));
    if (h->cfg.target_kind == PGN_TARGET_UNID) {   // initialization = [0.5, 0.5] (test/test_DistributionLogPotential.jl:14)
      std::vector<double> x0((size_t)nl * h->d_pad, 0.0);
      for (int i = 0; i < nl; ++i) x0[(size_t)i * h->d_pad] = x0[(size_t)i * h->d_pad + 1] = 0.5;
      h->x.upload(x0.data(), x0.size(), h->stream);
    }
    if (h->cfg.target_kind == PGN_TARGET_TOY_MVN) {
      Params P;
      fill_params(h, P);
      const int wpb = 4;
      launch_init_toy((nl + wpb - 1) / wpb, wpb * 32, h->stream, P);
      CUDA_CHECK(cudaGetLastError());
      CUDA_CHECK(cudaStreamSynchronize(h->stream));
    }
    h->initialised = true;
  } catch (CudaError& e) { return fail(err, e.code, e.msg); }
  return PGN_OK;
}

int pgn_get_state(pgn_handle* h, pgn_replica_state* out, char** err) {
  try {
    use_device(h);
    const int nl = h->n_local, d = h->cfg.dim;
    if (out->x && d > 0) {
      std::vector<double> raw((size_t)nl * h->d_pad);
      h->x.download(raw.data(), raw.size(), h->stream);
      if (h->cfg.target_kind == PGN_TARGET_ISING) {
        const int L = (int)h->cfg.p[1];
        for (int i = 0; i < nl; ++i) {
          const unsigned int* rows = reinterpret_cast<const unsigned int*>(&raw[(size_t)i * h->d_pad]);
          for (int a = 0; a < L; ++a)
            for (int b = 0; b < L; ++b) out->x[(size_t)i * d + a * L + b] = ((rows[a] >> b) & 1u) ? 1.0 : 0.0;
        }
      } else {
        for (int i = 0; i < nl; ++i) std::memcpy(out->x + (size_t)i * d, &raw[(size_t)i * h->d_pad], sizeof(double) * d);
      }
    }
    if (out->replica_index) h->replica_index.download(out->replica_index, nl, h->stream);
    if (out->rng_counter) h->rng_ctr.download(reinterpret_cast<unsigned long long*>(out->rng_counter), nl, h->stream);
    if (out->round_trip_state) h->rt_state.download(out->round_trip_state, nl, h->stream);
  } catch (CudaError& e) { return fail(err, e.code, e.msg); }
  return PGN_OK;
}

int pgn_set_state(pgn_handle* h, const pgn_replica_state* in, char** err) {
  if (!h->initialised) return fail(err, PGN_ERR_INVALID, "call pgn_init_replicas first");
  try {
    use_device(h);
    const int nl = h->n_local, d = h->cfg.dim;
    if (in->x && d > 0) {
      std::vector<double> raw((size_t)nl * h->d_pad, 0.0);
      if (h->cfg.target_kind == PGN_TARGET_ISING) {
        const int L = (int)h->cfg.p[1];
        for (int i = 0; i < nl; ++i) {
          unsigned int* rows = reinterpret_cast<unsigned int*>(&raw[(size_t)i * h->d_pad]);
          for (int a = 0; a < L; ++a)
            for (int b = 0; b < L; ++b)
              if (in->x[(size_t)i * d + a * L + b] != 0.0) rows[a] |= (1u << b);
        }
      } else {
        for (int i = 0; i < nl; ++i) std::memcpy(&raw[(size_t)i * h->d_pad], in->x + (size_t)i * d, sizeof(double) * d);
      }
      h->x.upload(raw.data(), raw.size(), h->stream);
    }
    if (in->replica_index) h->replica_index.upload(in->replica_index, nl, h->stream);
    if (in->rng_counter) h->rng_ctr.upload(reinterpret_cast<const unsigned long long*>(in->rng_counter), nl, h->stream);
    if (in->round_trip_state) h->rt_state.upload(in->round_trip_state, nl, h->stream);
  } catch (CudaError& e) { return fail(err, e.code, e.msg); }
  return PGN_OK;
}

int pgn_run_round(pgn_handle* h, int64_t n_scans, pgn_round_out* out, char** err) {
  if (!h || !out) return fail(err, PGN_ERR_INVALID, "null argument");
  if (!h->initialised) return fail(err, PGN_ERR_INVALID, "call pgn_init_replicas first");
  if (n_scans < 0 || n_scans >= (1LL << 32)) return fail(err, PGN_ERR_INVALID, "n_scans out of range");
  if (h->cfg.world_size > 1) {
    if (h->cfg.rank > 0 && !h->mail_left) return fail(err, PGN_ERR_INVALID, "left neighbour mailbox not attached");
    if (h->cfg.rank < h->cfg.world_size - 1 && !h->mail_right)
      return fail(err, PGN_ERR_INVALID, "right neighbour mailbox not attached");
  }
  try {
    use_device(h);
    const bool is_logreg = h->cfg.target_kind == PGN_TARGET_LOGREG;
    void* kernel = is_logreg ? nullptr : select_scan_kernel(h);
    if (!kernel && !is_logreg && !select_mem_kernel(h))
      throw CudaError{PGN_ERR_INVALID, "explorer not supported for this target on the device"};
    const int nl = h->n_local, d = h->cfg.dim;
    h->epoch += 1;
    Params P;
    fill_params(h, P);
    P.n_scans = n_scans;
    h->scan_seq += (unsigned long long)n_scans;   // P.tag_base holds the count before this round
    LrParams LP;
    if (is_logreg) logreg_fill_params(h, LP);
    // optional event logs
    StreamBuf<int> d_index;          // stream-ordered: see StreamBuf in pgn_host.hpp
    StreamBuf<double> d_lr, d_u, d_trace;
    StreamBuf<unsigned char> d_acc;
    const size_t nlog = (size_t)n_scans * nl;
    if (out->index_process) { d_index.alloc(nlog, h->stream); P.index_process = d_index.p; }
    if (out->swap_lr) { d_lr.alloc(nlog, h->stream); P.swap_lr = d_lr.p; }
    if (out->swap_u) { d_u.alloc(nlog, h->stream); P.swap_u = d_u.p; }
    if (out->swap_accept) { d_acc.alloc(nlog, h->stream); P.swap_accept = d_acc.p; }
    const bool owns_target = shard_owns_target(h);
    const int n_tgt = (h->cfg.n_chains_variational > 0 && h->cfg.n_chains_variational < h->cfg.n_chains) ? 2 : 1;
    if (out->target_trace && owns_target) { d_trace.alloc((size_t)n_scans * n_tgt * std::max(d, 1), h->stream); P.target_trace = d_trace.p; }
    CUDA_CHECK(cudaMemsetAsync(h->error_flag.p, 0, sizeof(int), h->stream));
    const bool per_replica = h->recorder_order == PGN_RECORDERS_PER_REPLICA;
    const bool vec_online = h->cfg.target_kind != PGN_TARGET_ISING && h->cfg.target_kind != PGN_TARGET_TEST_SWAPPER && d > 0;
    if (per_replica && n_scans > 0) {   // recorders are emptied every round (recorders.jl:113-118)
      launch_init_recorder_tables(h->stream, h->rec_table.p, h->rec_table.n, h->n_sms);
      if (owns_target && vec_online) CUDA_CHECK(cudaMemsetAsync(h->on_table.p, 0, h->on_table.n * sizeof(OnEntry), h->stream));
    }
    float ms = 0.f;
    std::vector<ChainStatsDev> st(nl);
    if (is_logreg) {
      LP.index_process = P.index_process; LP.swap_lr = P.swap_lr; LP.swap_u = P.swap_u;
      LP.swap_accept = P.swap_accept; LP.target_trace = P.target_trace;
      logreg_run_round(h, n_scans, LP, st, ms);
    } else {
      // launch geometry of the register-resident kernel: one warp per chain, all warps co-resident
      size_t smem = scan_smem_bytes(h);
      int wpb = 0, grid = 0;
      // every rank must pick the same kernel family (the register-resident kernels and the memory-resident one lay
      // their mailbox slots out differently), so co-residency is judged on the largest shard of the ladder
      const int nl_max = (h->cfg.n_chains + std::max(h->cfg.world_size, 1) - 1) / std::max(h->cfg.world_size, 1) +
                         ((h->cfg.n_chains_variational > 0 && h->cfg.n_chains_variational < h->cfg.n_chains) ? 1 : 0);   // shard_range
      const bool vec_target = h->cfg.target_kind == PGN_TARGET_TOY_MVN || h->cfg.target_kind == PGN_TARGET_FUNNEL ||
                              h->cfg.target_kind == PGN_TARGET_GMM;
      if (kernel && h->cfg.target_kind == PGN_TARGET_ISING && std::getenv("PGN_ISING_LITE") != nullptr &&
          std::string(std::getenv("PGN_ISING_LITE")) == "1") {   // tests: force the table-free Ising kernel
        kernel = ising_lite_scan_kernel();
        wpb = 1; grid = nl; smem = 0;
      } else
      if (kernel && !(h->force_mem && vec_target)) {
        if (is_team_kernel(h)) {
          // block = the team of W warps serving one chain.  W = the widest team (<= 8) whose blocks are all
          // co-resident with at most PGN_TEAM_WARPS_PER_SMSP (default 3) warps per scheduler; PGN_TEAM=W pins it.
          const char* e = std::getenv("PGN_TEAM");
          const char* e2 = std::getenv("PGN_TEAM_WARPS_PER_SMSP");
          const int pinned = e ? std::atoi(e) : 0;
          const int per_smsp = e2 ? std::max(1, std::atoi(e2)) : 3;
          cudaFuncAttributes fa;
          CUDA_CHECK(cudaFuncGetAttributes(&fa, kernel));
          const int w_max = std::min(8, fa.maxThreadsPerBlock / 32);   // the kernel's launch bounds
          for (int w = std::min(w_max, (pinned >= 1 && pinned <= 8) ? pinned : 8); w >= 1; --w) {
            // a team shares one scan's momentum draws through shared memory when they fit in 64 KB
            int max_refresh = h->ep.n_refresh;
            for (int v = 0; v < std::max(h->ep.n_mix, h->ep.n_steps) && v < PGN_MAX_MIX; ++v)
              max_refresh = std::max(max_refresh, h->ep.mix_n_refresh[v]);
            const int pool_max = (w > 1 && (size_t)max_refresh * (h->cpl * 32 + 8) * sizeof(double) <= 64 * 1024) ? max_refresh : 0;
            const bool wide_ok = w == 1 || pinned == w || (long long)nl * w <= (long long)per_smsp * 4 * h->n_sms;
            bool chosen = false;
            for (int pool = pool_max; pool >= 0 && !chosen; pool = pool > 0 ? 0 : -1) {   // with the momentum pool if it fits, else without
              const size_t sm = scan_smem_bytes(h, w, 1, pool);
              if (sm > 48 * 1024) CUDA_CHECK(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm));
              int per_sm = 0;
              CUDA_CHECK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kernel, w * 32, sm));
              const bool fits = (long long)per_sm * h->n_sms >= nl_max;
              if (fits && wide_ok) { wpb = w; grid = nl; smem = sm; P.pool_refresh = pool; chosen = true; }
            }
            if (chosen) break;
          }
        } else {
          for (int w = 1; w <= 8; w *= 2) {
            const size_t sm = scan_smem_bytes(h, 0, w);
            if (sm > 48 * 1024) CUDA_CHECK(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm));
            int per_sm = 0;
            CUDA_CHECK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kernel, w * 32, sm));
            const int g = (nl + w - 1) / w;
            if ((long long)per_sm * h->n_sms >= (nl_max + w - 1) / w) { wpb = w; grid = g; smem = sm; break; }
          }
          if (wpb == 0 && h->cfg.target_kind == PGN_TARGET_ISING) {
            // more Ising chains than fit with one ratio table per chain in shared memory (~1.9 K): the table-free kernel
            // (64 registers, no shared memory) holds up to 32 chains per SM; same mailbox layout, same results
            kernel = ising_lite_scan_kernel();
            for (int w = 1; w <= 2; w *= 2) {
              int per_sm = 0;
              CUDA_CHECK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kernel, w * 32, 0));
              const int g = (nl + w - 1) / w;
              if ((long long)per_sm * h->n_sms >= (nl_max + w - 1) / w) { wpb = w; grid = g; smem = 0; break; }
            }
          }
        }
      }
      // ---- mixed teams: a single-warp-per-chain autoMALA launch usually leaves warp slots free (C3: 1024 chains at 255
      // registers, 1184 slots).  Blocks of two warps then serve EITHER one chain as a team of two OR two chains with one warp
      // each, and the teams go to the chains that did most work in the previous round (explorer_n_steps: independent of the
      // team width).  The ladder runs at the local pace of its slowest region, so that is where a second warp pays.
      // Same kernels per chain as the uniform launch (every team width gives the same bits), same mailbox layout.
      StreamBuf<int> d_block_map;
      h->last_mixed_teams = 0;
      {
        const char* mx = std::getenv("PGN_MIXED_TEAMS");        // "1" / "0"; unset: kMixedTeamsDefault
        // a pinned team width (PGN_TEAM) is honoured as it is unless mixed teams are asked for explicitly
        const bool want = mx != nullptr ? std::string(mx) == "1" : (kMixedTeamsDefault && std::getenv("PGN_TEAM") == nullptr);
        const char* mcap = std::getenv("PGN_MIXED_MAX_TEAMS");   // tests: cap the number of teams (mixes team and pair blocks on small ladders)
        void* mk = nullptr;
        if (want && wpb == 1 && is_team_kernel(h) && h->ep.kind == PGN_EXPLORER_AUTOMALA && !h->var_active && nl >= 4) {
          switch (h->cfg.target_kind) {
            case PGN_TARGET_TOY_MVN: mk = vec_mixed_team_kernel_toy(h->cpl); break;
            case PGN_TARGET_FUNNEL: mk = vec_mixed_team_kernel_funnel(h->cpl); break;
            case PGN_TARGET_GMM: mk = vec_mixed_team_kernel_gmm(h->cpl); break;
            default: break;
          }
        }
        if (mk) {
          // shared memory of a block: the staged target constants + two single-warp team regions
          const size_t sm_mixed = scan_smem_bytes(h) + (size_t)2 * ((24 + h->cpl * 32) + 3 * (3 * h->cpl * 32 + 8)) * sizeof(double);
          if (sm_mixed > 48 * 1024) CUDA_CHECK(cudaFuncSetAttribute(mk, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm_mixed));
          int per_sm = 0;
          CUDA_CHECK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, mk, 64, sm_mixed));
          const long long blocks_max = (long long)per_sm * h->n_sms;
          // K teams need K + ceil((nl - K) / 2) blocks
          int K = (int)std::min<long long>(nl, 2 * blocks_max - nl);
          if (mcap && std::atoi(mcap) >= 0) K = std::min(K, std::atoi(mcap));
          while (K > 0 && K + (nl - K + 1) / 2 > blocks_max) --K;
          if (K >= 1 && K + (nl - K + 1) / 2 <= blocks_max) {
            std::vector<int> order(nl);
            for (int i = 0; i < nl; ++i) order[i] = i;
            if ((int)h->last_explore.size() == nl)
              std::stable_sort(order.begin(), order.end(), [&](int a, int b) { return h->last_explore[a] > h->last_explore[b]; });
            else
              std::reverse(order.begin(), order.end());   // no history yet: the top of the shard
            std::vector<char> team(nl, 0);
            for (int i = 0; i < K; ++i) team[order[i]] = 1;
            std::vector<int> map;
            for (int c = 0; c < nl; ++c) if (team[c]) { map.push_back(c); map.push_back(-2); }
            int pending = -1;
            for (int c = 0; c < nl; ++c) {
              if (team[c]) continue;
              if (pending < 0) pending = c;
              else { map.push_back(pending); map.push_back(c); pending = -1; }
            }
            if (pending >= 0) { map.push_back(pending); map.push_back(-1); }
            d_block_map.alloc(map.size(), h->stream);
            d_block_map.upload(map.data(), map.size());
            P.block_map = d_block_map.p;
            P.pool_refresh = 0;
            kernel = mk; wpb = 2; grid = (int)(map.size() / 2); smem = sm_mixed;
            h->last_mixed_teams = K;
          }
        }
      }
      if (wpb != 0) {
        void* args[] = {(void*)&P};
        CUDA_CHECK(cudaEventRecord(h->ev0, h->stream));
        if (n_scans > 0)
          CUDA_CHECK(cudaLaunchCooperativeKernel(kernel, dim3(grid), dim3(wpb * 32), args, smem, h->stream));
        CUDA_CHECK(cudaEventRecord(h->ev1, h->stream));
        CUDA_CHECK(cudaStreamSynchronize(h->stream));
        CUDA_CHECK(cudaEventElapsedTime(&ms, h->ev0, h->ev1));
        if (n_scans > 0) h->stats.download(st.data(), nl, h->stream);
        else std::memset(st.data(), 0, sizeof(ChainStatsDev) * nl);
        if (n_scans > 0 && is_team_kernel(h)) {   // work per chain of this round (leapfrog steps of the reference's sequential walk)
          h->last_explore.resize(nl);
          for (int i = 0; i < nl; ++i) h->last_explore[i] = st[i].n_steps;
        }
      } else {
        // memory-resident kernel: d > 128, or more chains than fit co-resident, or PGN_FORCE_MEM=1
        if (h->cfg.n_chains_variational > 0 && (h->cfg.n_chains_variational < h->cfg.n_chains || h->var_active))
          throw CudaError{PGN_ERR_INVALID, "two legs / a variational reference run on the register-resident scan kernels only "
                                           "(all chains of the shard co-resident, d <= 128)"};
        void* mk = vec_target ? select_mem_kernel(h) : nullptr;
        if (!mk) throw CudaError{PGN_ERR_INVALID, "too many chains for one GPU for this target (no memory-resident variant)"};
        mem_allocate(h);
        MemParams MP;
        mem_fill_params(h, P, MP);
        std::vector<int> ri(nl);
        std::vector<unsigned long long> ctr(nl);
        h->replica_index.download(ri.data(), nl, h->stream);
        h->rng_ctr.download(ctr.data(), nl, h->stream);
        std::vector<MemRec> rec(nl);
        std::memset(rec.data(), 0, sizeof(MemRec) * nl);
        for (int i = 0; i < nl; ++i) {
          rec[i].ctr = ctr[i]; rec[i].replica_index = ri[i];
          rec[i].ls_fwd.value = -INFINITY; rec[i].ls_bwd.value = -INFINITY;
        }
        h->mem_rec.upload(rec.data(), nl, h->stream);
        CUDA_CHECK(cudaMemsetAsync(h->online_mean.p, 0, sizeof(double) * h->d_pad, h->stream));
        CUDA_CHECK(cudaMemsetAsync(h->online_s2.p, 0, sizeof(double) * h->d_pad, h->stream));
        CUDA_CHECK(cudaMemsetAsync(h->online_n.p, 0, sizeof(long long), h->stream));
        const int mwpb = 4;
        int per_sm = 0;
        CUDA_CHECK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, mk, mwpb * 32, 0));
        if (per_sm < 1) throw CudaError{PGN_ERR_CUDA, "memory-resident scan kernel does not fit on an SM"};
        const int max_blocks = per_sm * h->n_sms;
        const int mgrid = std::min(max_blocks, (nl + mwpb - 1) / mwpb);
        void* args[] = {(void*)&MP};
        CUDA_CHECK(cudaEventRecord(h->ev0, h->stream));
        if (n_scans > 0)
          CUDA_CHECK(cudaLaunchCooperativeKernel(mk, dim3(mgrid), dim3(mwpb * 32), args, 0, h->stream));
        CUDA_CHECK(cudaEventRecord(h->ev1, h->stream));
        CUDA_CHECK(cudaStreamSynchronize(h->stream));
        CUDA_CHECK(cudaEventElapsedTime(&ms, h->ev0, h->ev1));
        h->mem_rec.download(rec.data(), nl, h->stream);
        std::vector<int> rt(nl);
        for (int i = 0; i < nl; ++i) {
          const MemRec& r = rec[i];
          ri[i] = r.replica_index; ctr[i] = r.ctr; rt[i] = r.rt_state;
          ChainStatsDev& o = st[i];
          o.swap_n = r.swap_acc.n; o.swap_mean = r.swap_acc.mu; o.ls_fwd = r.ls_fwd.value; o.ls_bwd = r.ls_bwd.value;
          o.expl_acc_n = r.expl_acc.n; o.expl_acc_mean = r.expl_acc.mu; o.n_steps = r.n_steps;
          o.am_n = r.am.n; o.am_mean = r.am.mu; o.rev_n = r.rev.n; o.rev_mean = r.rev.mu;
          o.n_restarts = r.n_restarts; o.n_round_trips = r.n_trips; o.n_points = r.n_points; o.n_ref_evals = r.n_ref;
        }
        h->replica_index.upload(ri.data(), nl, h->stream);
        h->rng_ctr.upload(ctr.data(), nl, h->stream);
        h->rt_state.upload(rt.data(), nl, h->stream);
      }
    }
    int flag = 0;
    int merge_launches = 0;
    h->error_flag.download(&flag, 1, h->stream);
    if (per_replica && n_scans > 0 && flag == 0) {
      // reduce_recorders!: the tree merge over replica indices replaces the float statistics of every chain
      CUDA_CHECK(cudaEventRecord(h->ev0, h->stream));
      launch_merge_recorders(h->stream, h->rec_table.p, h->cfg.n_chains, nl, h->stats.p);
      if (owns_target && vec_online)
        launch_merge_online(h->stream, h->on_table.p, h->cfg.n_chains, d, h->d_pad, h->online_mean.p, h->online_s2.p, h->online_n.p);
      CUDA_CHECK(cudaGetLastError());
      CUDA_CHECK(cudaEventRecord(h->ev1, h->stream));
      std::vector<ChainStatsDev> merged(nl);
      h->stats.download(merged.data(), nl, h->stream);
      float merge_ms = 0.f;
      CUDA_CHECK(cudaEventElapsedTime(&merge_ms, h->ev0, h->ev1));
      ms += merge_ms;                         // part of the round's device time
      merge_launches = owns_target && vec_online ? 2 : 1;
      for (int i = 0; i < nl; ++i) {
        ChainStatsDev& s = st[i];
        const ChainStatsDev& m = merged[i];
        s.swap_n = m.swap_n; s.swap_mean = m.swap_mean; s.ls_fwd = m.ls_fwd; s.ls_bwd = m.ls_bwd;
        s.expl_acc_n = m.expl_acc_n; s.expl_acc_mean = m.expl_acc_mean;
        s.am_n = m.am_n; s.am_mean = m.am_mean; s.rev_n = m.rev_n; s.rev_mean = m.rev_mean;
      }
    }

    if (const char* dump = std::getenv("PGN_TIMING_DUMP")) {   // diagnostics: per-chain explore / partner-wait clocks of this round
      if (FILE* f = std::fopen(dump, "a")) {
        for (int i = 0; i < nl; ++i)
          std::fprintf(f, "%u %d %lld %lld %lld %lld %lld %lld %lld\n", h->epoch, h->first_chain + i, (long long)n_scans, st[i].explore_cycles,
                       st[i].wait_cycles, st[i].n_points, st[i].trial_cycles, st[i].barrier_cycles, st[i].decide_cycles);
        std::fclose(f);
      }
    }
    long long restarts = 0, trips = 0, pts = 0, evals = 0;
    for (int i = 0; i < nl; ++i) {
      const ChainStatsDev& s = st[i];
      if (out->swap_n) out->swap_n[i] = s.swap_n;
      if (out->swap_mean) out->swap_mean[i] = s.swap_mean;
      if (out->logsum_fwd) out->logsum_fwd[i] = n_scans > 0 ? s.ls_fwd : -INFINITY;
      if (out->logsum_bwd) out->logsum_bwd[i] = n_scans > 0 ? s.ls_bwd : -INFINITY;
      if (out->expl_acc_n) out->expl_acc_n[i] = s.expl_acc_n;
      if (out->expl_acc_mean) out->expl_acc_mean[i] = s.expl_acc_mean;
      if (out->expl_n_steps) out->expl_n_steps[i] = s.n_steps;
      if (out->am_n) out->am_n[i] = s.am_n;
      if (out->am_mean) out->am_mean[i] = s.am_mean;
      if (out->rev_n) out->rev_n[i] = s.rev_n;
      if (out->rev_mean) out->rev_mean[i] = s.rev_mean;
      restarts += s.n_restarts; trips += s.n_round_trips; pts += s.n_points; evals += s.n_ref_evals;
    }
    out->n_tempered_restarts = restarts;
    out->n_round_trips = trips;
    out->n_density_points = pts;
    out->n_ref_equiv_evals = evals;
    out->kernel_ms = (double)ms;
    out->gemm_ms = is_logreg ? h->last_gemm_ms : 0.0;
    out->batch_steps = is_logreg ? h->last_batch_steps : 0;
    out->n_launches = (is_logreg ? h->last_launches : (n_scans > 0 ? 1 : 0)) + merge_launches;
    out->active_columns = is_logreg ? h->last_active_cols : 0;
    out->gemm_columns = is_logreg ? h->last_gemm_cols : 0;
    out->online_n = 0;
    if (owns_target && n_scans > 0) {
      long long on = 0;
      h->online_n.download(&on, 1, h->stream);
      const bool vec = h->cfg.target_kind != PGN_TARGET_ISING && h->cfg.target_kind != PGN_TARGET_TEST_SWAPPER;
      out->online_n = vec ? on : n_scans;
      if (vec && d > 0) {
        std::vector<double> mu(h->d_pad), s2(h->d_pad);
        h->online_mean.download(mu.data(), h->d_pad, h->stream);
        h->online_s2.download(s2.data(), h->d_pad, h->stream);
        for (int c = 0; c < d; ++c) {
          if (out->online_mean) out->online_mean[c] = mu[c];
          if (out->online_var) out->online_var[c] = on > 1 ? s2[c] * ((double)on / (double)(on - 1)) : 1.0;
        }
      }
    }
    if (out->index_process) d_index.download(out->index_process, nlog);
    if (out->swap_lr) d_lr.download(out->swap_lr, nlog);
    if (out->swap_u) d_u.download(out->swap_u, nlog);
    if (out->swap_accept) d_acc.download(out->swap_accept, nlog);
    if (out->target_trace && owns_target && d > 0) d_trace.download(out->target_trace, (size_t)n_scans * n_tgt * d);
    if (flag != 0) {
      static const char* names[] = {"", "invalid explorer state (autoMALA bounds / step size)", "", "",
                                    "Got NaN log-unnormalized ratio", "non-finite log density in SliceSampler",
                                    "slice_shrink: maximum number of iterations reached",
                                    "autoMALA: could not find a positive step size",
                                    "AutoMALA can only be called on a configuration of positive density",
                                    "neighbour hand-shake timed out"};
      return fail(err, flag, flag >= 1 && flag <= 9 ? names[flag] : "device error");
    }
  } catch (CudaError& e) { return fail(err, e.code, e.msg); }
  return PGN_OK;
}

int pgn_log_potential(pgn_handle* h, const double* x, int32_t n_points, const double* beta, double* out, char** err) {
  try {
    use_device(h);
    const int d = h->cfg.dim;
    if (h->cfg.target_kind == PGN_TARGET_TEST_SWAPPER) return fail(err, PGN_ERR_INVALID, "TestSwapper has no log_potential");
    if (h->cfg.target_kind == PGN_TARGET_LOGREG) {
      logreg_points(h, x, n_points, beta, out, nullptr, nullptr);
      return PGN_OK;
    }
    DevBuf<double> dx, db, dout;
    const int ldx = h->cpl == 0 && h->cfg.target_kind != PGN_TARGET_ISING ? h->d_pad : d;   // d > 128: padded rows
    dx.alloc((size_t)n_points * ldx, true); db.alloc(n_points, false); dout.alloc(n_points, false);
    if (ldx == d) dx.upload(x, (size_t)n_points * d);
    else CUDA_CHECK(cudaMemcpy2D(dx.p, sizeof(double) * ldx, x, sizeof(double) * d, sizeof(double) * d, n_points, cudaMemcpyHostToDevice));
    db.upload(beta, n_points);
    Params P;
    fill_params(h, P);
    switch (h->cfg.target_kind) {
      case PGN_TARGET_TOY_MVN: case PGN_TARGET_FUNNEL: case PGN_TARGET_GMM: case PGN_TARGET_MIXED: case PGN_TARGET_UNID:
        launch_eval_points(h, P, dx.p, db.p, n_points, dout.p, nullptr, nullptr); break;
      case PGN_TARGET_ISING: launch_ising_lp((n_points + 3) / 4, 128, h->stream, P, dx.p, db.p, n_points, dout.p); break;
      default: return fail(err, PGN_ERR_INVALID, "unsupported target");
    }
    CUDA_CHECK(cudaGetLastError());
    CUDA_CHECK(cudaStreamSynchronize(h->stream));
    dout.download(out, n_points);
  } catch (CudaError& e) { return fail(err, e.code, e.msg); }
  return PGN_OK;
}

int pgn_logdensity_and_gradient(pgn_handle* h, const double* x, int32_t n_points, const double* beta, double* logdens,
                                double* grad, char** err) {
  try {
    use_device(h);
    const int d = h->cfg.dim;
    const int tk = h->cfg.target_kind;
    if (tk == PGN_TARGET_LOGREG) {
      logreg_points(h, x, n_points, beta, nullptr, logdens, grad);
      return PGN_OK;
    }
    if (tk != PGN_TARGET_TOY_MVN && tk != PGN_TARGET_FUNNEL && tk != PGN_TARGET_GMM)
      return fail(err, PGN_ERR_INVALID, "target has no gradient");
    DevBuf<double> dx, db, dld, dg;
    const int ldx = h->cpl == 0 ? h->d_pad : d;   // d > 128: padded rows
    dx.alloc((size_t)n_points * ldx, true); db.alloc(n_points, false); dld.alloc(n_points, false);
    dg.alloc((size_t)n_points * ldx, true);
    if (ldx == d) dx.upload(x, (size_t)n_points * d);
    else CUDA_CHECK(cudaMemcpy2D(dx.p, sizeof(double) * ldx, x, sizeof(double) * d, sizeof(double) * d, n_points, cudaMemcpyHostToDevice));
    db.upload(beta, n_points);
    Params P;
    fill_params(h, P);
    launch_eval_points(h, P, dx.p, db.p, n_points, nullptr, dld.p, dg.p);
    CUDA_CHECK(cudaGetLastError());
    CUDA_CHECK(cudaStreamSynchronize(h->stream));
    dld.download(logdens, n_points);
    if (ldx == d) dg.download(grad, (size_t)n_points * d);
    else CUDA_CHECK(cudaMemcpy2D(grad, sizeof(double) * d, dg.p, sizeof(double) * ldx, sizeof(double) * d, n_points, cudaMemcpyDeviceToHost));
  } catch (CudaError& e) { return fail(err, e.code, e.msg); }
  return PGN_OK;
}

int pgn_set_variational(pgn_handle* h, const double* mean, const double* sd, char** err) {
  if (!h) return fail(err, PGN_ERR_INVALID, "null argument");
  if (!mean || !sd) { h->var_active = false; return PGN_OK; }
  const int tk = h->cfg.target_kind, d = h->cfg.dim;
  if (h->cfg.n_chains_variational < 1) return fail(err, PGN_ERR_INVALID, "set_variational: n_chains_variational is 0");
  if ((tk != PGN_TARGET_FUNNEL && tk != PGN_TARGET_GMM && tk != PGN_TARGET_UNID) || h->cpl == 0 || h->force_mem)
    return fail(err, PGN_ERR_INVALID, "set_variational: FUNNEL, GMM and UNID targets on the register-resident kernels (d <= 128)");
  for (int c = 0; c < d; ++c)
    if (!(sd[c] > 0.0) || !std::isfinite(sd[c]) || !std::isfinite(mean[c]))
      return fail(err, PGN_ERR_INVALID, "set_variational: finite means and positive finite standard deviations");
  try {
    use_device(h);
    std::vector<double> host((size_t)5 * h->d_pad, 0.0);
    for (int c = 0; c < d; ++c) { host[c] = mean[c]; host[(size_t)h->d_pad + c] = sd[c]; }
    for (int c = d; c < h->d_pad; ++c) host[(size_t)h->d_pad + c] = 1.0;
    h->var_tab.upload(host.data(), host.size(), h->stream);
    launch_var_tables(h->stream, h->var_tab.p, d, h->d_pad);
    CUDA_CHECK(cudaGetLastError());
    CUDA_CHECK(cudaStreamSynchronize(h->stream));
    h->var_active = true;
  } catch (CudaError& e) { return fail(err, e.code, e.msg); }
  return PGN_OK;
}

int pgn_hamiltonian_dynamics(pgn_handle* h, const double* x, const double* p, int32_t n_points, const double* beta,
                             const double* diag_precond, double step_size, int32_t n_steps, double* x_out, double* p_out,
                             char** err) {
  try {
    use_device(h);
    const int d = h->cfg.dim, tk = h->cfg.target_kind;
    if ((tk != PGN_TARGET_TOY_MVN && tk != PGN_TARGET_FUNNEL && tk != PGN_TARGET_GMM) || h->cpl == 0)
      return fail(err, PGN_ERR_INVALID, "hamiltonian_dynamics: vector targets with a gradient, d <= 128");
    if (n_points < 0 || n_steps < 0) return fail(err, PGN_ERR_INVALID, "negative count");
    StreamBuf<double> dx, dp, db, ox, op, dpre;
    const size_t n = (size_t)n_points * d;
    if (diag_precond) {
      for (int c = 0; c < d; ++c)
        if (!(diag_precond[c] != 0.0) || !std::isfinite(diag_precond[c])) return fail(err, PGN_ERR_INVALID, "diag_precond: finite and non-zero");
      dpre.alloc(d, h->stream); dpre.upload(diag_precond, d);
    }
    dx.alloc(n, h->stream); dp.alloc(n, h->stream); db.alloc(n_points, h->stream); ox.alloc(n, h->stream); op.alloc(n, h->stream);
    dx.upload(x, n); dp.upload(p, n); db.upload(beta, n_points);
    Params P;
    fill_params(h, P);
    const int wpb = 4, grid = (n_points + wpb - 1) / wpb;
    const size_t smem = scan_smem_bytes(h);
    if (tk == PGN_TARGET_TOY_MVN) launch_leapfrog_toy(h->cpl, grid, wpb * 32, smem, h->stream, P, dx.p, dp.p, db.p, dpre.p, step_size, n_steps, n_points, ox.p, op.p);
    else if (tk == PGN_TARGET_FUNNEL && h->var_active) launch_leapfrog_funnel_var(h->cpl, grid, wpb * 32, smem, h->stream, P, dx.p, dp.p, db.p, dpre.p, step_size, n_steps, n_points, ox.p, op.p);
    else if (tk == PGN_TARGET_FUNNEL) launch_leapfrog_funnel(h->cpl, grid, wpb * 32, smem, h->stream, P, dx.p, dp.p, db.p, dpre.p, step_size, n_steps, n_points, ox.p, op.p);
    else if (h->var_active) launch_leapfrog_gmm_var(h->cpl, grid, wpb * 32, smem, h->stream, P, dx.p, dp.p, db.p, dpre.p, step_size, n_steps, n_points, ox.p, op.p);
    else launch_leapfrog_gmm(h->cpl, grid, wpb * 32, smem, h->stream, P, dx.p, dp.p, db.p, dpre.p, step_size, n_steps, n_points, ox.p, op.p);
    CUDA_CHECK(cudaGetLastError());
    ox.download(x_out, n); op.download(p_out, n);
  } catch (CudaError& e) { return fail(err, e.code, e.msg); }
  return PGN_OK;
}

int pgn_ipc_export(pgn_handle* h, void* handle64, char** err) {
  try {
    use_device(h);
    cudaIpcMemHandle_t mh;
    CUDA_CHECK(cudaIpcGetMemHandle(&mh, h->mail.p));
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "cudaIpcMemHandle_t is 64 bytes");
    std::memcpy(handle64, &mh, 64);
  } catch (CudaError& e) { return fail(err, e.code, e.msg); }
  return PGN_OK;
}

int pgn_ipc_attach(pgn_handle* h, int32_t side, const void* handle64, char** err) {
  try {
    use_device(h);
    cudaIpcMemHandle_t mh;
    std::memcpy(&mh, handle64, 64);
    void* p = nullptr;
    CUDA_CHECK(cudaIpcOpenMemHandle(&p, mh, cudaIpcMemLazyEnablePeerAccess));
    if (side == 0) { h->mail_left = (char*)p; h->left_is_ipc = true; }
    else { h->mail_right = (char*)p; h->right_is_ipc = true; }
  } catch (CudaError& e) { return fail(err, e.code, e.msg); }
  return PGN_OK;
}

int pgn_peer_attach(pgn_handle* h, int32_t side, pgn_handle* neighbour, char** err) {
  try {
    use_device(h);
    if (neighbour->cfg.device != h->cfg.device) {
      int can = 0;
      CUDA_CHECK(cudaDeviceCanAccessPeer(&can, h->cfg.device, neighbour->cfg.device));
      if (!can) throw CudaError{PGN_ERR_CUDA, "devices cannot access each other's memory"};
      cudaError_t e = cudaDeviceEnablePeerAccess(neighbour->cfg.device, 0);
      if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled) CUDA_CHECK(e);
      cudaGetLastError();
    }
    if (side == 0) h->mail_left = neighbour->mail.p; else h->mail_right = neighbour->mail.p;
  } catch (CudaError& e) { return fail(err, e.code, e.msg); }
  return PGN_OK;
}

int pgn_measure_fp64_peak(int32_t device, double* tflops, char** err) {
  try {
    int count = 0;
    if (cudaGetDeviceCount(&count) != cudaSuccess || count == 0)
      return fail(err, PGN_ERR_NO_DEVICE, "no usable CUDA device (this library has no CPU fallback)");
    *tflops = logreg_measure_fp64_peak(device);
  } catch (CudaError& e) { return fail(err, e.code, e.msg); }
  return PGN_OK;
}

int pgn_test_dmma(int32_t device, const double* a, const double* b, const double* c, double* d_out, int32_t n_trials,
                  char** err) {
  try {
    int count = 0;
    if (cudaGetDeviceCount(&count) != cudaSuccess || count == 0)
      return fail(err, PGN_ERR_NO_DEVICE, "no usable CUDA device (this library has no CPU fallback)");
    CUDA_CHECK(cudaSetDevice(device));
    logreg_test_dmma(a, b, c, d_out, n_trials);
  } catch (CudaError& e) { return fail(err, e.code, e.msg); }
  return PGN_OK;
}

int pgn_test_math(int32_t device, int32_t op, const double* in, double* out, int64_t n, int64_t seed,
                  int32_t replica_index, char** err) {
  try {
    int count = 0;
    if (cudaGetDeviceCount(&count) != cudaSuccess || count == 0)
      return fail(err, PGN_ERR_NO_DEVICE, "no usable CUDA device (this library has no CPU fallback)");
    CUDA_CHECK(cudaSetDevice(device));
    const size_t nin = op == 6 ? 2 * (size_t)n : (size_t)n;
    DevBuf<double> din, dout;
    din.alloc(nin, false); dout.alloc(n, false);
    din.upload(in, nin);
    launch_test_math((int)((n + 127) / 128), 128, op, din.p, dout.p, n, (unsigned int)(unsigned long long)seed,
                     (unsigned int)((unsigned long long)seed >> 32), (unsigned int)replica_index);
    CUDA_CHECK(cudaGetLastError());
    CUDA_CHECK(cudaDeviceSynchronize());
    dout.download(out, n);
  } catch (CudaError& e) { return fail(err, e.code, e.msg); }
  return PGN_OK;
}

}  // extern "C"
