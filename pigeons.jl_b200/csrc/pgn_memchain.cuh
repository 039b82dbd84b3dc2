// pgn_memchain.cuh — memory-resident variant of the scan kernel.
//
// The register-resident kernel (pgn_kernels.cuh) needs d <= 128 and every chain of the
// shard co-resident (one warp per chain).  This variant lifts both limits with the same
// algorithms, the same arithmetic order and therefore the same bits:
//   * all per-chain vectors (state, momentum, gradients, search scratch) live in HBM/L2
//     as [chain][d_pad] rows, lane l owning the coordinates l, l+32, ... (one "slot" per 32);
//   * a warp serves the chains w, w+W, w+2W, ... of the shard: per scan it first explores and
//     posts all of them (phase A), then completes all their swaps (phase B).  Posts of a scan
//     never wait on anything, so the pairwise hand-shakes cannot deadlock although one warp
//     carries many chains; the mailbox ring argument of the fast kernel is unchanged.
// Selected automatically when d > 128 or the shard has more chains than fit co-resident
// (PGN_FORCE_MEM=1 forces it, which is how the parity tests exercise it on every case).
#pragma once
#include "pgn_kernels.cuh"
#include "pgn_memchain_types.cuh"

namespace pgn {

template <int TK, int EX>
struct MemChain {
  const Params* P;
  const MemParams* M;
  const double* means;   // global memory (GMM), [KMAX][d_pad] then log weights
  int lane, d, nslots, cl;
  double beta;
  double e0, e1;
  Rng rng;
  MeanAcc expl_acc, am, rev;
  long long n_steps, n_points, n_ref;
  int err;
  double pool;
  unsigned long long pool_base;
  bool pool_valid;
  // preconditioner of the current step (Preconditioner.jl:57-77): 0 identity, 1 1/sd, 2 mix + rmix/sd
  int pre_mode;
  double mix, rmix;

  __device__ __forceinline__ double* row(double* base) const { return base + (size_t)cl * P->d_pad; }
  __device__ __forceinline__ bool valid(int c) const { return c < d; }
  __device__ __forceinline__ double draw_uniform() {
    unsigned long long idx = rng.ctr - pool_base;
    if (!pool_valid || idx >= 32ull) {
      pool_base = rng.ctr;
      pool = uniform_at(rng, pool_base + (unsigned long long)lane);
      pool_valid = true;
      idx = 0;
    }
    rng.ctr += 1;
    return __shfl_sync(PGN_FULL_MASK, pool, (int)idx);
  }
  __device__ __forceinline__ double draw_exponential() { return -log_(1.0 - draw_uniform()); }
  __device__ __forceinline__ double toy_precision(double b) const { return (1.0 - b) * P->p[0] + b * P->p[1]; }
  __device__ __forceinline__ double lp_call(double b, double a0, double a1) const {
    if (TK == PGN_TARGET_TOY_MVN) return -0.5 * toy_precision(b) * a0;
    if (b == 0.0) return a0;
    if (b == 1.0) return a1;
    return (1.0 - b) * a0 + b * a1;
  }
  __device__ __forceinline__ double lp_ad(double b, double a0, double a1) const {
    if (TK == PGN_TARGET_TOY_MVN) return -0.5 * toy_precision(b) * a0;
    return (1.0 - b) * a0 + b * a1;
  }
  __device__ __forceinline__ double pre_at(int c) const {
    if (pre_mode == 0) return 1.0;
    const double sd = P->std_devs[c];
    if (sd == 0.0) return 1.0;
    return pre_mode == 1 ? 1.0 / sd : mix + rmix / sd;
  }

  // densities at xv with coordinate c_over replaced by v_over (c_over < 0: no override)
  __device__ void eval(const double* xv, int c_over, double v_over, double& a0, double& a1) {
    n_points += 1;
    auto X = [&](int c) { return c == c_over ? v_over : xv[c]; };
    if (TK == PGN_TARGET_TOY_MVN) {
      double acc = 0.0;
      for (int k = 0; k < nslots; ++k) { const int c = k * 32 + lane; if (valid(c)) { const double t = X(c); acc = acc + t * t; } }
      a0 = warp_sum(acc); a1 = 0.0;
    } else if (TK == PGN_TARGET_FUNNEL) {
      const double y = X(0);
      const double e = exp_(-y);
      const double sy = P->p[0], lsy = P->p[1], ivr = P->p[5], lsr = P->p[4];
      double v[2] = {0.0, 0.0};
      for (int k = 0; k < nslots; ++k) {
        const int c = k * 32 + lane;
        if (!valid(c)) continue;
        const double xc = X(c);
        v[0] = v[0] + (-(xc * xc * ivr + PGN_LOG2PI) * 0.5 - lsr);
        if (c == 0) { double zy = y / sy; v[1] = v[1] + (-(zy * zy + PGN_LOG2PI) * 0.5 - lsy); }
        else { double t = xc * xc * e; v[1] = v[1] + (-(t + PGN_LOG2PI) * 0.5 - 0.5 * y); }
      }
      warp_sum_n<2>(v);
      a0 = v[0]; a1 = v[1];
    } else {
      const double ivr = P->p[5], lsr = P->p[4], ivm = P->p[2], cst = P->p[1];
      double v[KMAX_MODES + 1];
#pragma unroll
      for (int m = 0; m <= KMAX_MODES; ++m) v[m] = 0.0;
      for (int k = 0; k < nslots; ++k) {
        const int c = k * 32 + lane;
        if (!valid(c)) continue;
        const double xc = X(c);
        v[KMAX_MODES] = v[KMAX_MODES] + (-(xc * xc * ivr + PGN_LOG2PI) * 0.5 - lsr);
#pragma unroll
        for (int m = 0; m < KMAX_MODES; ++m) { const double t = xc - means[(size_t)m * P->d_pad + c]; v[m] = v[m] + t * t; }
      }
      warp_sum_n<KMAX_MODES + 1>(v);
      a0 = v[KMAX_MODES];
      const double* lw = means + (size_t)KMAX_MODES * P->d_pad;
      double a[KMAX_MODES];
      double Mx = -PGN_INF;
#pragma unroll
      for (int m = 0; m < KMAX_MODES; ++m) { a[m] = lw[m] - 0.5 * v[m] * ivm - cst; if (a[m] > Mx) Mx = a[m]; }
      double s = 0.0;
#pragma unroll
      for (int m = 0; m < KMAX_MODES; ++m) s = s + exp_(a[m] - Mx);
      a1 = Mx + log_(s);
    }
  }
  // densities + beta-combined raw gradient written to graw[]; returns via a0, a1
  __device__ void eval_grad(const double* xv, double b, double& a0, double& a1, double* graw) {
    n_points += 1;
    if (TK == PGN_TARGET_TOY_MVN) {
      double acc = 0.0;
      for (int k = 0; k < nslots; ++k) { const int c = k * 32 + lane; if (valid(c)) acc = acc + xv[c] * xv[c]; }
      a0 = warp_sum(acc); a1 = 0.0;
      const double prec = toy_precision(b);
      for (int k = 0; k < nslots; ++k) { const int c = k * 32 + lane; graw[c] = valid(c) ? -prec * xv[c] : 0.0; }
    } else if (TK == PGN_TARGET_FUNNEL) {
      const double y = xv[0];
      const double e = exp_(-y);
      const double sy = P->p[0], lsy = P->p[1], ivy = P->p[2], ivr = P->p[5], lsr = P->p[4];
      double v[3] = {0.0, 0.0, 0.0};
      for (int k = 0; k < nslots; ++k) {
        const int c = k * 32 + lane;
        if (!valid(c)) continue;
        const double xc = xv[c];
        v[0] = v[0] + (-(xc * xc * ivr + PGN_LOG2PI) * 0.5 - lsr);
        if (c == 0) { double zy = y / sy; v[1] = v[1] + (-(zy * zy + PGN_LOG2PI) * 0.5 - lsy); v[2] = v[2] + 0.0; }
        else { double t = xc * xc * e; v[1] = v[1] + (-(t + PGN_LOG2PI) * 0.5 - 0.5 * y); v[2] = v[2] + (0.5 * (xc * xc * e) - 0.5); }
      }
      warp_sum_n<3>(v);
      a0 = v[0]; a1 = v[1];
      const double T = v[2];
      for (int k = 0; k < nslots; ++k) {
        const int c = k * 32 + lane;
        if (!valid(c)) { graw[c] = 0.0; continue; }
        const double xc = xv[c];
        const double gr = -xc * ivr;
        const double gt = c == 0 ? (-y * ivy + T) : (-xc * e);
        const double t = gr * (1.0 - b);
        graw[c] = t + gt * b;
      }
    } else {
      const double ivr = P->p[5], lsr = P->p[4], ivm = P->p[2], cst = P->p[1];
      double v[KMAX_MODES + 1];
#pragma unroll
      for (int m = 0; m <= KMAX_MODES; ++m) v[m] = 0.0;
      for (int k = 0; k < nslots; ++k) {
        const int c = k * 32 + lane;
        if (!valid(c)) continue;
        const double xc = xv[c];
        v[KMAX_MODES] = v[KMAX_MODES] + (-(xc * xc * ivr + PGN_LOG2PI) * 0.5 - lsr);
#pragma unroll
        for (int m = 0; m < KMAX_MODES; ++m) { const double t = xc - means[(size_t)m * P->d_pad + c]; v[m] = v[m] + t * t; }
      }
      warp_sum_n<KMAX_MODES + 1>(v);
      a0 = v[KMAX_MODES];
      const double* lw = means + (size_t)KMAX_MODES * P->d_pad;
      double w[KMAX_MODES];
      double Mx = -PGN_INF;
#pragma unroll
      for (int m = 0; m < KMAX_MODES; ++m) { w[m] = lw[m] - 0.5 * v[m] * ivm - cst; if (w[m] > Mx) Mx = w[m]; }
      double s = 0.0;
#pragma unroll
      for (int m = 0; m < KMAX_MODES; ++m) { w[m] = exp_(w[m] - Mx); s = s + w[m]; }
      a1 = Mx + log_(s);
      for (int k = 0; k < nslots; ++k) {
        const int c = k * 32 + lane;
        if (!valid(c)) { graw[c] = 0.0; continue; }
        const double xc = xv[c];
        double acc = 0.0;
#pragma unroll
        for (int m = 0; m < KMAX_MODES; ++m) acc = acc + w[m] * (means[(size_t)m * P->d_pad + c] - xc);
        const double gt = (acc / s) * ivm;
        const double gr = -xc * ivr;
        const double t = gr * (1.0 - b);
        graw[c] = t + gt * b;
      }
    }
  }

  __device__ void sample_iid(double b) {
    double* x = row(P->x);
    for (int k = 0; k < nslots; ++k) {
      const int c = k * 32 + lane;
      if (!valid(c)) continue;
      const double z = normal_at(rng, rng.ctr + (unsigned long long)c);
      x[c] = TK == PGN_TARGET_TOY_MVN ? z / sqrt(toy_precision(b)) : P->p[3] * z;
    }
    rng.ctr += (unsigned long long)d;
    __syncwarp();
  }

  // ---- SliceSampler (src/explorers/SliceSampler.jl:24-237) ----
  __device__ __forceinline__ double lp_at(int c, double v) {
    double a0, a1;
    eval(row(P->x), c, v, a0, a1);
    n_ref += 1;
    return lp_call(beta, a0, a1);
  }
  static __device__ __forceinline__ bool isapprox(double a, double b) {
    const double rtol = bits_to_double(0x3e50000000000000ULL);
    if (a == b) return true;
    if (!(is_finite(a) && is_finite(b))) return false;
    double aa = fabs(a), ab = fabs(b);
    return fabs(a - b) <= rtol * (aa > ab ? aa : ab);
  }
  __device__ bool slice_accept(int c, double old_position, double new_position, double z, double L, double R,
                               double lp_L, double lp_R) {
    const double w = P->slice_w;
    double Lhat = L, Rhat = R;
    bool Rstale = false, Lstale = false, D = false;
    while (Rhat - Lhat > 1.1 * w) {
      double Mid = (Lhat + Rhat) / 2.0;
      if (((old_position < Mid) && (new_position >= Mid)) || ((old_position >= Mid) && (new_position < Mid))) D = true;
      if (new_position < Mid) { Rhat = Mid; Rstale = true; } else { Lhat = Mid; Lstale = true; }
      if (D) {
        if (Lstale) { lp_L = lp_at(c, Lhat); Lstale = false; }
        if (Rstale) { lp_R = lp_at(c, Rhat); Rstale = false; }
        if ((z >= lp_L) && (z >= lp_R)) { expl_acc.fit(0.0); return false; }
      }
    }
    expl_acc.fit(1.0);
    return true;
  }
  __device__ double slice_coord(int c, double cached_lp) {
    const double w = P->slice_w;
    double* x = row(P->x);
    const double cur = x[c];
    const double z = cached_lp - draw_exponential();
    double L = cur - w * draw_uniform();
    double R = L + w;
    int K = P->slice_p;
    double lp_L = lp_at(c, L);
    double lp_R = lp_at(c, R);
    while (K > 0 && ((z < lp_L) || (z < lp_R))) {
      double V = draw_uniform();
      if (V <= 0.5) { L = L - (R - L); lp_L = lp_at(c, L); }
      else { R = R + (R - L); lp_R = lp_at(c, R); }
      K -= 1;
    }
    n_steps += P->slice_p - K;
    double Lbar = L, Rbar = R;
    int n = 1;
    while (n <= P->slice_max_iter) {
      double new_position = Lbar + draw_uniform() * (Rbar - Lbar);
      double new_lp = lp_at(c, new_position);
      bool consider = z < new_lp;
      if (consider && slice_accept(c, cur, new_position, z, L, R, lp_L, lp_R)) {
        if (lane == (c & 31)) x[c] = new_position;
        __syncwarp();
        n_steps += n;
        return new_lp;
      }
      if (new_position < cur) Lbar = new_position; else Rbar = new_position;
      if (isapprox(Lbar, Rbar)) { n_steps += n; return lp_at(c, cur); }
      n += 1;
    }
    err = PGN_ERR_SLICE_MAX_ITER;
    return 0.0;
  }
  __device__ void slice_step() {
    double cached_lp = -PGN_INF;
    for (int pass = 0; pass < P->slice_n_passes; ++pass) {
      if (cached_lp == -PGN_INF) {
        double a0, a1;
        eval(row(P->x), -1, 0.0, a0, a1);
        n_ref += 1;
        double result = lp_call(beta, a0, a1);
        if (result == -PGN_INF) { err = PGN_ERR_BAD_DENSITY; return; }
        cached_lp = result;
      }
      for (int c = 0; c < d; ++c) {
        cached_lp = slice_coord(c, cached_lp);
        if (err) return;
        if (!is_finite(cached_lp)) { err = PGN_ERR_BAD_DENSITY; return; }
      }
    }
    eval(row(P->x), -1, 0.0, e0, e1);
  }

  // ---- autoMALA / MALA with memory-resident vectors ----
  struct TrialS { double a0, a1, lp1, h_after, eps; };
  __device__ double run_trial(const double* sx, const double* sp, const double* sg, double eps, double h_before, TrialS& T) {
    double* tx = row(M->VTX); double* tp = row(M->VTP); double* tg = row(M->VTG);
    const double half_eps = eps / 2;
    double acc = 0.0;
    for (int k = 0; k < nslots; ++k) {
      const int c = k * 32 + lane;
      if (!valid(c)) { tx[c] = 0.0; continue; }
      const double ph = sp[c] + half_eps * sg[c];
      tx[c] = sx[c] + eps * (pre_mode == 0 ? ph : ph / pre_at(c));
      acc = acc + ph * ph;
    }
    const double pp = warp_sum(acc);
    __syncwarp();
    eval_grad(tx, beta, T.a0, T.a1, tg);
    T.lp1 = lp_ad(beta, T.a0, T.a1);
    const double cur = T.lp1 - 0.5 * pp;
    double s2;
    if (!is_finite(cur)) {
      for (int k = 0; k < nslots; ++k) {
        const int c = k * 32 + lane;
        if (valid(c)) { tg[c] = pre_mode == 0 ? tg[c] : tg[c] / pre_at(c); tp[c] = sp[c] + half_eps * sg[c]; } else tp[c] = 0.0;
      }
      s2 = pp;
    } else {
      double q = 0.0;
      for (int k = 0; k < nslots; ++k) {
        const int c = k * 32 + lane;
        if (!valid(c)) { tp[c] = 0.0; continue; }
        const double g1c = pre_mode == 0 ? tg[c] : tg[c] / pre_at(c);
        tg[c] = g1c;
        const double ph = sp[c] + half_eps * sg[c];
        const double p1 = ph + half_eps * g1c;
        tp[c] = p1;
        q = q + p1 * p1;
      }
      s2 = warp_sum(q);
    }
    __syncwarp();
    T.h_after = T.lp1 - 0.5 * s2;
    T.eps = eps;
    return T.h_after - h_before;
  }
  __device__ void copy(double* dst, const double* src, bool negate = false) const {
    for (int k = 0; k < nslots; ++k) { const int c = k * 32 + lane; dst[c] = negate ? src[c] * -1.0 : src[c]; }
    __syncwarp();
  }
  __device__ void build_preconditioner() {
    if (P->std_devs == nullptr || P->precond_kind == PGN_PRECOND_IDENTITY) { pre_mode = 0; return; }
    if (P->precond_kind == PGN_PRECOND_DIAGONAL) { pre_mode = 1; return; }
    const double u = draw_uniform();
    if (u <= P->mix_p0) pre_mode = 1;
    else if (u <= P->mix_p01) pre_mode = 0;
    else { pre_mode = 2; mix = draw_uniform(); rmix = 1.0 - mix; }
  }
  __device__ void gradient_sampler(bool use_mh, int n_refresh_eff, bool mala) {
    double* x = row(P->x);
    pre_mode = 0;
    if (n_refresh_eff > 0) build_preconditioner();
    if (!(P->step_size > 0)) { err = PGN_ERR_INVALID; return; }
    double* g0 = row(M->VG0);
    double lp0;
    {
      double* tg = row(M->VTG);
      eval_grad(x, beta, e0, e1, tg);
      for (int k = 0; k < nslots; ++k) { const int c = k * 32 + lane; g0[c] = (pre_mode == 0 || !valid(c)) ? tg[c] : tg[c] / pre_at(c); }
      __syncwarp();
      lp0 = lp_ad(beta, e0, e1);
    }
    TrialS T;
    double* p = row(M->VP);
    for (int i = 0; i < n_refresh_eff; ++i) {
      double acc = 0.0;
      for (int k = 0; k < nslots; ++k) {
        const int c = k * 32 + lane;
        const double z = valid(c) ? normal_at(rng, rng.ctr + (unsigned long long)c) : 0.0;
        p[c] = z;
        if (valid(c)) acc = acc + z * z;
      }
      rng.ctr += (unsigned long long)d;
      __syncwarp();
      const double init_joint = lp0 - 0.5 * warp_sum(acc);
      if (!is_finite(init_joint)) { err = PGN_ERR_NOT_POSITIVE; return; }
      if (mala) {   // MALA.jl:74-97
        run_trial(x, p, g0, P->step_size, init_joint, T);
        const double e = exp_(T.h_after - init_joint);
        const double prob = 1.0 < e ? 1.0 : e;
        expl_acc.fit(prob);
        n_ref += 4;
        if (draw_uniform() < prob) {
          copy(x, row(M->VTX)); copy(g0, row(M->VTG));
          e0 = T.a0; e1 = T.a1; lp0 = T.lp1;
        }
        n_steps += 1;
        continue;
      }
      double mine = uniform_at(rng, rng.ctr + (unsigned long long)(lane < 3 ? lane : 0));
      double lmine = log_(mine);
      rng.ctr += use_mh ? 3ull : 2ull;
      const double a = __shfl_sync(PGN_FULL_MASK, mine, 0), b = __shfl_sync(PGN_FULL_MASK, mine, 1);
      const double la = __shfl_sync(PGN_FULL_MASK, lmine, 0), lb = __shfl_sync(PGN_FULL_MASK, lmine, 1);
      const double u_mh = __shfl_sync(PGN_FULL_MASK, mine, 2);
      const double lower = a < b ? la : lb, upper = a < b ? lb : la;
      if (!(lower < upper)) { err = PGN_ERR_INVALID; return; }
      const double *sx = x, *sp = p, *sg = g0;
      double f_a0 = 0.0, f_a1 = 0.0, f_lp = 0.0, h_rev = 0.0;
      int expo[2] = {0, 0};
      double h_before = init_joint;
      const int n_dir = use_mh ? 2 : 1;
      for (int dir = 0; dir < n_dir; ++dir) {
        int mode = 0, n = 0, exponent = 0, nst = 0;
        double eps = P->step_size;
        while (true) {
          const double diff = run_trial(sx, sp, sg, eps, h_before, T);
          bool decided = false;
          if (mode == 0) {
            if (!is_finite(diff) || diff < lower) { mode = 1; n = 1; eps = eps / 2.0; }
            else if (diff > upper) { mode = 2; n = 1; eps = eps * 2.0; }
            else decided = true;
          } else if (mode == 1) {
            if (eps == 0.0) { err = PGN_ERR_STEP_UNDERFLOW; return; }
            if (diff > lower) { nst = n; exponent = -n; decided = true; }
            else { n += 1; eps = eps / 2.0; }
          } else if (mode == 2) {
            if (!is_finite(diff) || diff < upper) { nst = n; exponent = n - 1; decided = true; }
            else { n += 1; eps = eps * 2.0; }
          } else {
            break;
          }
          if (decided) {
            const double eps_final = P->step_size * pow2(exponent);
            if (T.eps == eps_final) break;
            mode = 3; eps = eps_final;
          }
        }
        n_steps += 1 + nst;
        am.fit(pow2(exponent));
        expo[dir] = exponent;
        if (dir == 0) {
          n_ref += 1 + 1 + 3 * (1 + nst) + 2;
          h_rev = T.h_after; h_before = h_rev;
          f_a0 = T.a0; f_a1 = T.a1; f_lp = T.lp1;
          copy(row(M->VFX), row(M->VTX));
          copy(row(M->VFG), row(M->VTG));
          if (use_mh) {
            copy(row(M->VSX), row(M->VTX));
            copy(row(M->VSP), row(M->VTP), true);
            copy(row(M->VSG), row(M->VTG));
            sx = row(M->VSX); sp = row(M->VSP); sg = row(M->VSG);
          }
        } else {
          n_ref += 1 + 3 * (1 + nst);
        }
      }
      bool accept = true;
      if (use_mh) {
        const bool passed = (expo[1] == expo[0]);
        rev.fit(passed ? 1.0 : 0.0);
        double prob = 0.0;
        if (passed) { double e = exp_(h_rev - init_joint); prob = 1.0 < e ? 1.0 : e; n_ref += 1; }
        expl_acc.fit(prob);
        accept = u_mh < prob;
      }
      if (accept) {
        copy(x, row(M->VFX)); copy(g0, row(M->VFG));
        e0 = f_a0; e1 = f_a1; lp0 = f_lp;
      }
    }
  }

  __device__ void explore(long long scan, bool is_reference) {
    if (is_reference) { sample_iid(beta); eval(row(P->x), -1, 0.0, e0, e1); return; }
    if (EX == PGN_EXPLORER_TOY) { sample_iid(beta); eval(row(P->x), -1, 0.0, e0, e1); }
    else if (EX == PGN_EXPLORER_SLICE) slice_step();
    else if (EX == PGN_EXPLORER_MALA) gradient_sampler(true, P->n_refresh, true);
    else gradient_sampler(scan != 1, P->n_refresh, false);
  }
  __device__ double log_ratio(double beta_partner) const { return lp_call(beta_partner, e0, e1) - lp_call(beta, e0, e1); }
};

template <int TK, int EX>
__global__ void __launch_bounds__(256) scan_kernel_mem(const __grid_constant__ MemParams MP) {
  const Params& P = MP.base;
  const int lane = threadIdx.x & 31;
  const int w = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int W = gridDim.x * (blockDim.x >> 5);
  const int N = P.n_chains;
  const int last_local = P.first_chain + P.n_local - 1;
  MemChain<TK, EX> ch;
  ch.P = &P; ch.M = &MP; ch.means = P.means; ch.lane = lane; ch.d = P.d; ch.nslots = MP.nslots;
  int err = 0;

  for (long long scan = 1; scan <= P.n_scans && err == 0; ++scan) {
    const bool even = (scan & 1LL) == 0;
    const int ring = (int)((P.epoch & 1u) * 4u + (unsigned int)(scan & 3LL));
    const unsigned int tag = P.tag_base + (unsigned int)scan;   // flag-in-data mailbox words, as in scan_kernel
    // ---------------- phase A: explore + post, for every chain this warp serves ----------------
    for (int cl = w; cl < P.n_local && err == 0; cl += W) {
      const int chain = P.first_chain + cl;
      const bool is_ref = (chain == 1 && N > 1), is_tgt = (chain == N);
      MemRec r = MP.rec[cl];
      ch.cl = cl; ch.beta = P.beta[chain - 1];
      ch.rng.key0 = P.seed_lo; ch.rng.key1 = (unsigned int)r.replica_index; ch.rng.c2 = P.seed_hi; ch.rng.c3 = 0u;
      ch.rng.ctr = r.ctr;
      ch.expl_acc = r.expl_acc; ch.am = r.am; ch.rev = r.rev;
      ch.n_steps = r.n_steps; ch.n_points = r.n_points; ch.n_ref = r.n_ref;
      ch.err = 0; ch.pool_valid = false; ch.pool = 0.0; ch.pool_base = 0; ch.e0 = r.e0; ch.e1 = r.e1;
      ch.pre_mode = 0; ch.mix = 0.0; ch.rmix = 0.0;
      ch.explore(scan, is_ref);
      if (ch.err) { err = ch.err; break; }
      double* x = P.x + (size_t)cl * P.d_pad;
      if (is_tgt) {   // target-chain recording (pigeons.jl:110-131)
        const long long n = r.on_n + 1;
        const double g = 1.0 / (double)n;
        for (int c = lane; c < P.d; c += 32) {
          const double mu_old = P.online_mean[c];
          const double mu = mu_old + g * (x[c] - mu_old);
          P.online_s2[c] = P.online_s2[c] + g * ((x[c] - mu) * (x[c] - mu_old) - P.online_s2[c]);
          P.online_mean[c] = mu;
        }
        r.on_n = n;
        if (lane == 0) *P.online_n = n;
        if (P.target_trace)
          for (int c = lane; c < P.d; c += 32) P.target_trace[(size_t)(scan - 1) * P.d + c] = x[c];
      }
      int partner = chain + ((((chain & 1) == 0) == even) ? 1 : -1);
      if (partner == 0) partner = 1;
      if (partner == N + 1) partner = N;
      const double lr = ch.log_ratio(P.beta[partner - 1]);
      ch.n_ref += 2;
      if (lr != lr) { err = PGN_ERR_NAN_RATIO; break; }
      const double u = ch.draw_uniform();
      const size_t log_at = (size_t)(scan - 1) * P.n_local + cl;
      if (lane == 0) {
        if (P.index_process) P.index_process[log_at] = r.replica_index;
        if (P.swap_lr) P.swap_lr[log_at] = lr;
        if (P.swap_u) P.swap_u[log_at] = u;
      }
      if (r.rt_state == 0 && is_ref) r.rt_state = 1;
      else if (r.rt_state == 1 && is_tgt) { r.rt_state = 2; r.n_restarts += 1; }
      else if (r.rt_state == 2 && is_ref) { r.rt_state = 1; r.n_trips += 1; }
      r.ctr = ch.rng.ctr; r.e0 = ch.e0; r.e1 = ch.e1;
      r.expl_acc = ch.expl_acc; r.am = ch.am; r.rev = ch.rev;
      r.n_steps = ch.n_steps; r.n_points = ch.n_points; r.n_ref = ch.n_ref;
      r.lr = lr; r.u = u;
      if (partner != chain) {
        const bool remote = partner < P.first_chain || partner > last_local;
        char* dst;
        if (!remote) dst = P.mail + ((size_t)(2 + cl) * MAIL_RINGS + ring) * P.slot_bytes;
        else if (partner > chain) dst = P.mail_right + ((size_t)0 * MAIL_RINGS + ring) * P.slot_bytes;
        else dst = P.mail_left + ((size_t)1 * MAIL_RINGS + ring) * P.slot_bytes;
        unsigned long long* dstw = reinterpret_cast<unsigned long long*>(dst);
        {
          const unsigned long long lb = double_to_bits(lr), ub = double_to_bits(u), cb = r.ctr;
          unsigned int hv = (unsigned int)lb;
          hv = lane == 1 ? (unsigned int)(lb >> 32) : hv;
          hv = lane == 2 ? (unsigned int)ub : hv;
          hv = lane == 3 ? (unsigned int)(ub >> 32) : hv;
          hv = lane == 4 ? (unsigned int)cb : hv;
          hv = lane == 5 ? (unsigned int)(cb >> 32) : hv;
          hv = lane == 6 ? (unsigned int)r.replica_index : hv;
          hv = lane == 7 ? (unsigned int)r.rt_state : hv;
          if (lane < LL_HDR_WORDS) ll_store(dstw + lane, hv, tag);
        }
        for (int c = lane; c < P.d_pad; c += 32) {
          const unsigned long long b = double_to_bits(x[c]);
          ll_store(dstw + LL_HDR_WORDS + 2 * c, (unsigned int)b, tag);
          ll_store(dstw + LL_HDR_WORDS + 2 * c + 1, (unsigned int)(b >> 32), tag);
        }
      }
      if (lane == 0) MP.rec[cl] = r;
      __syncwarp();
    }
    if (err) break;
    // ---------------- phase B: complete the swaps ----------------
    for (int cl = w; cl < P.n_local && err == 0; cl += W) {
      const int chain = P.first_chain + cl;
      int partner = chain + ((((chain & 1) == 0) == even) ? 1 : -1);
      if (partner == 0) partner = 1;
      if (partner == N + 1) partner = N;
      const size_t log_at = (size_t)(scan - 1) * P.n_local + cl;
      bool accepted = false;
      if (partner != chain) {
        MemRec r = MP.rec[cl];
        const bool remote = partner < P.first_chain || partner > last_local;
        const char* src;
        if (!remote) src = P.mail + ((size_t)(2 + (partner - P.first_chain)) * MAIL_RINGS + ring) * P.slot_bytes;
        else if (partner > chain) src = P.mail + ((size_t)1 * MAIL_RINGS + ring) * P.slot_bytes;
        else src = P.mail + ((size_t)0 * MAIL_RINGS + ring) * P.slot_bytes;
        const unsigned long long* srcw = reinterpret_cast<const unsigned long long*>(src);
        int status = 0;
        unsigned long long hw = 0ull;
        {
          unsigned long long t0 = 0;
          unsigned int it = 0;
          while (true) {
            if (lane < LL_HDR_WORDS) hw = ld_relaxed_sys(srcw + lane);
            const bool ok = lane >= LL_HDR_WORDS || (unsigned int)(hw >> 32) == tag;
            if (__all_sync(PGN_FULL_MASK, ok)) break;
            ++it;
            if (it > 4u) __nanosleep(it < 64u ? 32 : 256);
            if ((it & 255u) == 0u) {
              if (lane == 0) {
                if (*reinterpret_cast<volatile int*>(P.error_flag) != 0) status = 1;
                const unsigned long long now = globaltimer_ns();
                if (t0 == 0) t0 = now;
                else if (now - t0 > P.timeout_ns) status = 2;
              }
              status = __shfl_sync(PGN_FULL_MASK, status, 0);
              if (status != 0) break;
            }
          }
        }
        if (status != 0) { err = status == 2 ? PGN_ERR_TIMEOUT : -1; break; }
        const unsigned int lo = (unsigned int)hw;
        const unsigned int h0 = __shfl_sync(PGN_FULL_MASK, lo, 0), h1 = __shfl_sync(PGN_FULL_MASK, lo, 1);
        const unsigned int h2 = __shfl_sync(PGN_FULL_MASK, lo, 2), h3 = __shfl_sync(PGN_FULL_MASK, lo, 3);
        const unsigned int h4 = __shfl_sync(PGN_FULL_MASK, lo, 4), h5 = __shfl_sync(PGN_FULL_MASK, lo, 5);
        const int ri_p = (int)__shfl_sync(PGN_FULL_MASK, lo, 6), rt_p = (int)__shfl_sync(PGN_FULL_MASK, lo, 7);
        const double lr_p = bits_to_double(((unsigned long long)h1 << 32) | h0);
        const double u_p = bits_to_double(((unsigned long long)h3 << 32) | h2);
        const bool lower = chain < partner;
        const double e = lower ? exp_(r.lr + lr_p) : exp_(lr_p + r.lr);
        const double acceptance_pr = 1.0 < e ? 1.0 : e;
        if (lower) { r.swap_acc.fit(acceptance_pr); r.ls_fwd.fit(r.lr); r.ls_bwd.fit(lr_p); }
        accepted = (lower ? r.u : u_p) < acceptance_pr;
        if (accepted) {
          if (P.rec_table != nullptr) {   // per-replica recorders: see scan_kernel
            RecEntry* eo = P.rec_table + (size_t)(r.replica_index - 1) * P.n_local + cl;
            const RecEntry* en = P.rec_table + (size_t)(ri_p - 1) * P.n_local + cl;
            if (lane == 0) {
              eo->expl_acc = r.expl_acc; eo->am = r.am; eo->rev = r.rev; eo->swap_acc = r.swap_acc;
              eo->ls_fwd = r.ls_fwd; eo->ls_bwd = r.ls_bwd;
            }
            r.expl_acc = en->expl_acc; r.am = en->am; r.rev = en->rev; r.swap_acc = en->swap_acc;
            r.ls_fwd = en->ls_fwd; r.ls_bwd = en->ls_bwd;
            if (chain == N && P.d > 0) {   // the target-chain online statistics travel with their replica too
              OnEntry* oo = P.on_table + (size_t)(r.replica_index - 1) * P.d_pad;
              const OnEntry* on = P.on_table + (size_t)(ri_p - 1) * P.d_pad;
              for (int c = lane; c < P.d; c += 32) {
                oo[c] = OnEntry{r.on_n, P.online_mean[c], P.online_s2[c]};
                const OnEntry e = on[c];
                P.online_mean[c] = e.mu; P.online_s2[c] = e.s2;
              }
              r.on_n = on[0].n;
              if (lane == 0) *P.online_n = r.on_n;
            }
          }
          r.replica_index = ri_p;
          r.rt_state = rt_p;
          r.ctr = ((unsigned long long)h5 << 32) | h4;
          double* x = P.x + (size_t)cl * P.d_pad;
          bool got = true;
          for (int c = lane; c < P.d_pad; c += 32) {
            unsigned int xlo, xhi;
            got = ll_load(srcw + LL_HDR_WORDS + 2 * c, tag, xlo) && got;
            got = ll_load(srcw + LL_HDR_WORDS + 2 * c + 1, tag, xhi) && got;
            x[c] = bits_to_double(((unsigned long long)xhi << 32) | xlo);
          }
          if (!__all_sync(PGN_FULL_MASK, got)) { err = PGN_ERR_TIMEOUT; break; }
        }
        if (lane == 0) MP.rec[cl] = r;
        __syncwarp();
      }
      if (lane == 0 && P.swap_accept) P.swap_accept[log_at] = accepted ? 1 : 0;
    }
  }
  if (err > 0 && lane == 0) atomicCAS(P.error_flag, 0, err);
  if (P.rec_table != nullptr && err == 0) {   // per-replica recorders: the entries of the replicas held when the round ends
    for (int cl = w; cl < P.n_local; cl += W) {
      const MemRec r = MP.rec[cl];
      RecEntry* eo = P.rec_table + (size_t)(r.replica_index - 1) * P.n_local + cl;
      if (lane == 0) {
        eo->expl_acc = r.expl_acc; eo->am = r.am; eo->rev = r.rev; eo->swap_acc = r.swap_acc;
        eo->ls_fwd = r.ls_fwd; eo->ls_bwd = r.ls_bwd;
      }
      if (P.first_chain + cl == N && P.d > 0) {
        OnEntry* oo = P.on_table + (size_t)(r.replica_index - 1) * P.d_pad;
        for (int c = lane; c < P.d; c += 32) oo[c] = OnEntry{r.on_n, P.online_mean[c], P.online_s2[c]};
      }
    }
  }
}

// parity entry points for d > 128: one warp per point
template <int TK>
__global__ void eval_points_mem_kernel(const __grid_constant__ MemParams MP, const double* xs, const double* betas,
                                       int n_points, double* lp_out, double* ld_out, double* grad_out) {
  const Params& P = MP.base;
  const int lane = threadIdx.x & 31;
  const int w = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (w >= n_points) return;
  MemChain<TK, PGN_EXPLORER_SLICE> ch;
  ch.P = &P; ch.M = &MP; ch.means = P.means; ch.lane = lane; ch.d = P.d; ch.nslots = MP.nslots; ch.cl = 0;
  ch.beta = betas[w]; ch.n_points = 0; ch.n_ref = 0; ch.err = 0; ch.pre_mode = 0;
  const double* x = xs + (size_t)w * P.d_pad;   // padded rows
  if (lp_out) {
    double a0, a1;
    ch.eval(x, -1, 0.0, a0, a1);
    if (lane == 0) lp_out[w] = ch.lp_call(ch.beta, a0, a1);
  }
  if (ld_out) {
    double a0, a1;
    double* g = grad_out + (size_t)w * P.d_pad;
    ch.eval_grad(x, ch.beta, a0, a1, g);
    if (lane == 0) ld_out[w] = (TK == PGN_TARGET_TOY_MVN) ? ch.lp_ad(ch.beta, a0, a1)
                                                          : ((0.0 + a0 * (1.0 - ch.beta)) + a1 * ch.beta);
  }
}

}  // namespace pgn
