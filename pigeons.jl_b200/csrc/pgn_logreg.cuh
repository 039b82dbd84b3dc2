// pgn_logreg.cuh — device path for the logistic-regression target (BASELINE config 5:
// d = 4096, n_data = 65536, autoMALA, 256 chains per GPU).
//
// This is the one target whose density is a dense contraction: every density /
// gradient evaluation is  z = X theta  followed by  g = X^T (y - sigmoid(z)).  All
// chains of a shard advance in lock-step "batch steps": each chain is a small state
// machine (the same autoMALA search as VecChain::automala, one density evaluation per
// step) that emits the next point it needs evaluated; the evaluations of all chains
// are then done together as two FP64 GEMMs
//     Z  = X  Theta      (n_data x d) (d x R)      + fused Bernoulli epilogue
//     G  = X^T Resid     (d x n_data) (n_data x R)  split-K over 4096-row chunks
// so X (2 GiB) is streamed once per GEMM for all R chains instead of once per chain.
//
// The GEMMs are hand-written with a FIXED summation order (sequential fma over k for every
// output element, split-K chunk partials added in chunk order), which is part of the arithmetic
// spec mirrored by the CPU oracle.  Two implementations produce identical bits:
//   dgemm_km_dmma_kernel  FP64 tensor cores (mma.sync.m8n8k4.f64 = DMMA.8x8x4), default (128 x 64 tiles, two blocks per SM): the
//                         hardware accumulates each instruction as an fma chain in ascending k
//                         (probed and asserted), 76 % of the measured FP64 peak;
//   dgemm_km_kernel       SIMT DFMA, 8x8 register tiles; limited by shared-memory bandwidth
//                         (16 operand doubles per 64 fma = 128 B/clk/SM), 56 % of peak.
// tcgen05 has no FP64 path, so DMMA/DFMA (64 fma/clk/SM) is the roofline here.
#pragma once
#include "pgn_kernels.cuh"
#include "pgn_logreg_types.cuh"

namespace pgn {


// ---------------------------------------------------------------------------
// C[m][n] = sum_k A[k][m] * B[k][n], k ascending, one fma per k, acc starts at 0.0.
// A: [K][lda] (M contiguous), B: [K][ldb] (N contiguous).  128x128 block tile,
// 8x8 register tile per thread, BK = 16, register-prefetch double buffering.
// EPI == 0: store C (row-major [m][ldc]) for split z into Cout + z*split_stride
// EPI == 1: Bernoulli epilogue: z -> LL[m][n], RES[m][n] (rows m >= n_valid_rows -> 0)
// ---------------------------------------------------------------------------
__device__ __forceinline__ void logreg_terms(double z, double y, double& ll, double& resid) {
  const double az = z < 0.0 ? -z : z;
  const double t = exp_(-az);
  const double sp = (z > 0.0 ? z : 0.0) + log1p_(t);
  const double sig = z >= 0.0 ? 1.0 / (1.0 + t) : t / (1.0 + t);
  ll = y * z - sp;
  resid = y - sig;
}

template <int EPI>
__global__ void __launch_bounds__(GEMM_THREADS)
dgemm_km_kernel(const double* __restrict__ A, int lda, const double* __restrict__ B, int ldb, int k_total,
                int k_chunk, double* __restrict__ C0, double* __restrict__ C1, int ldc, size_t split_stride,
                const double* __restrict__ yvec, int n_valid_rows, const int* __restrict__ n_cols_ptr) {
  // column tiles beyond the compacted list of this batch step have nothing to multiply (the count lives on the device,
  // the host launched the full grid without knowing it)
  if ((int)blockIdx.y * GEMM_BN >= *n_cols_ptr) return;
  extern __shared__ double gemm_smem[];
  double (*As)[GEMM_BK][GEMM_BM] = reinterpret_cast<double (*)[GEMM_BK][GEMM_BM]>(gemm_smem);
  double (*Bs)[GEMM_BK][GEMM_BN] = reinterpret_cast<double (*)[GEMM_BK][GEMM_BN]>(gemm_smem + 2 * GEMM_BK * GEMM_BM);
  const int tid = threadIdx.x;
  const int m0 = blockIdx.x * GEMM_BM, n0 = blockIdx.y * GEMM_BN;
  const int k_begin = blockIdx.z * k_chunk;
  const int k_end = min(k_total, k_begin + k_chunk);
  // loader mapping: 16 threads cover one 128-double row, consecutive threads load consecutive
  // 16-byte words (coalesced global loads, conflict-free shared stores)
  const int lk = tid >> 4, lo = (tid & 15) * 2;
  // compute mapping: 16 x 16 logical threads; a warp covers 4 (ty) x 8 (tx) of them so that its
  // shared-memory reads touch 4 x 64 B of As (broadcast) and 8 x 16 B of Bs per load.
  // thread (ty, tx) owns the row pairs {32 i + 2 ty, 32 i + 2 ty + 1}, i = 0..3, and the column pairs
  // {16 j + 2 tx, 16 j + 2 tx + 1}, j = 0..3, within the 64-column half selected by the warp parity
  // (both shared-memory reads are then 64 / 128 contiguous bytes per warp: one wavefront each).
  const int warp = tid >> 5, lane = tid & 31;
  const int ty = 4 * (warp >> 1) + (lane >> 3);
  const int tx = lane & 7;
  const int cbase = 64 * (warp & 1);
  double acc[8][8];
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[i][j] = 0.0;

  double2 ra[4], rb[4];
  auto gload = [&](int kt) {
    const double* ap = A + (size_t)(kt + lk) * lda + m0 + lo;
    const double* bp = B + (size_t)(kt + lk) * ldb + n0 + lo;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      ra[i] = *reinterpret_cast<const double2*>(ap + 32 * i);
      rb[i] = *reinterpret_cast<const double2*>(bp + 32 * i);
    }
  };
  auto sstore = [&](int buf) {
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      *reinterpret_cast<double2*>(&As[buf][lk][lo + 32 * i]) = ra[i];
      *reinterpret_cast<double2*>(&Bs[buf][lk][lo + 32 * i]) = rb[i];
    }
  };
  int buf = 0;
  if (k_begin < k_end) {
    gload(k_begin);
    sstore(0);
  }
  __syncthreads();
  for (int kt = k_begin; kt < k_end; kt += GEMM_BK) {
    const bool has_next = kt + GEMM_BK < k_end;
    if (has_next) gload(kt + GEMM_BK);
#pragma unroll
    for (int kk = 0; kk < GEMM_BK; ++kk) {
      double a[8], b[8];
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const double2 t = *reinterpret_cast<const double2*>(&As[buf][kk][32 * i + 2 * ty]);
        a[2 * i] = t.x; a[2 * i + 1] = t.y;
      }
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const double2 t = *reinterpret_cast<const double2*>(&Bs[buf][kk][cbase + 16 * j + 2 * tx]);
        b[2 * j] = t.x; b[2 * j + 1] = t.y;
      }
#pragma unroll
      for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[i][j] = fma(a[i], b[j], acc[i][j]);
    }
    if (has_next) sstore(buf ^ 1);
    __syncthreads();
    buf ^= 1;
  }
  // epilogue
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int m = m0 + 32 * (i >> 1) + 2 * ty + (i & 1);
    if (EPI == 0) {
      double* crow = C0 + (size_t)blockIdx.z * split_stride + (size_t)m * ldc + n0;
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        double2 v; v.x = acc[i][2 * j]; v.y = acc[i][2 * j + 1];
        *reinterpret_cast<double2*>(crow + cbase + 16 * j + 2 * tx) = v;
      }
    } else {
      const bool live = m < n_valid_rows;
      const double y = live ? yvec[m] : 0.0;
      double* llrow = C0 + (size_t)m * ldc + n0;
      double* rsrow = C1 + (size_t)m * ldc + n0;
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        double2 l, r;
        logreg_terms(acc[i][2 * j], y, l.x, r.x);
        logreg_terms(acc[i][2 * j + 1], y, l.y, r.y);
        if (!live) { l.x = l.y = 0.0; r.x = r.y = 0.0; }
        *reinterpret_cast<double2*>(llrow + cbase + 16 * j + 2 * tx) = l;
        *reinterpret_cast<double2*>(rsrow + cbase + 16 * j + 2 * tx) = r;
      }
    }
  }
}

// ---------------------------------------------------------------------------
// Same contraction on the FP64 tensor cores (DMMA, mma.sync.m8n8k4.f64).  Valid as a
// drop-in ONLY because the hardware accumulates the four k-products of one instruction as
// a sequential fma chain in ascending k starting from C (established by
// tools/probe_dmma_order.py and asserted by tests/test_gpu_parity.py::test_dmma_is_a_sequential_fma_chain),
// so chaining the instructions over k reproduces the spec's summation order bit for bit.
// Operands come from registers (1 double per lane per 8x4 / 4x8 fragment), which removes the
// shared-memory bandwidth limit of the SIMT kernel (16 operand doubles per 64 fma).
// Block tile 128x128, BK = 16, 8 warps; warp tile 32 (M) x 64 (N) = 4 x 8 DMMA tiles.
// ---------------------------------------------------------------------------
__device__ __forceinline__ void dmma_m8n8k4(double& d0, double& d1, double a, double b, double c0, double c1);
__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gmem_src) {
  const unsigned int d = (unsigned int)__cvta_generic_to_shared(smem_dst);
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(d), "l"(gmem_src) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_group 0;" ::: "memory"); }

// BN = 128: 8 warps (4 x 2), one block per SM.  BN = 64 (default): 4 warps (4 x 1), THREE blocks per SM — with a single
// block the tensor pipe idles whenever its 8 warps meet at the k-tile barrier (ncu, round 1: DMMA pipe 74 % active, 16 % of
// the stall samples on the barrier); independent blocks drift apart and fill each other's gaps (measured: 29.3 -> 30.7
// TFLOP/s with two blocks of the register-staged loader).  The narrower tile also
// halves the granularity of the column compaction.  The summation order per output element does not depend on the tile.
template <int EPI, int BN>
__global__ void __launch_bounds__(2 * BN, BN == 64 ? 3 : 1)
dgemm_km_dmma_kernel(const double* __restrict__ A, int lda, const double* __restrict__ B, int ldb, int k_total,
                     int k_chunk, double* __restrict__ C0, double* __restrict__ C1, int ldc, size_t split_stride,
                     const double* __restrict__ yvec, int n_valid_rows, const int* __restrict__ n_cols_ptr) {
  // column tiles beyond the compacted list of this batch step have nothing to multiply (the count lives on the device,
  // the host launched the full grid without knowing it)
  if ((int)blockIdx.y * BN >= *n_cols_ptr) return;
  constexpr int NT = 2 * BN;            // threads
  constexpr int WN = BN / 64;           // warp columns (each warp owns 32 rows x 64 columns)
  constexpr int LDB = BN + 4;           // padded row stride of the B tile, = 8 words mod 32 like DMMA_LD
  constexpr int TPR = NT / GEMM_BK;     // loader threads per k-row
  constexpr int STEP = 2 * TPR;         // doubles covered by one pass of a k-row's loader threads
  extern __shared__ double gemm_smem[];
  double (*As)[GEMM_BK][DMMA_LD] = reinterpret_cast<double (*)[GEMM_BK][DMMA_LD]>(gemm_smem);
  double (*Bs)[GEMM_BK][LDB] = reinterpret_cast<double (*)[GEMM_BK][LDB]>(gemm_smem + 2 * GEMM_BK * DMMA_LD);
  const int tid = threadIdx.x;
  const int m0 = blockIdx.x * GEMM_BM, n0 = blockIdx.y * BN;
  const int k_begin = blockIdx.z * k_chunk;
  const int k_end = min(k_total, k_begin + k_chunk);
  const int lk = tid / TPR, lo = (tid % TPR) * 2;
  const int warp = tid >> 5, lane = tid & 31;
  const int wm = (warp / WN) * 32;      // warp row offset inside the block tile (4 warp rows)
  const int wn = (warp % WN) * 64;      // warp column offset
  const int fr = lane >> 2, fk = lane & 3;
  double acc[4][8][2];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 8; ++j) { acc[i][j][0] = 0.0; acc[i][j][1] = 0.0; }
  constexpr int NA = GEMM_BM / STEP, NB = BN / STEP;
  // global -> shared with cp.async (16-byte LDGSTS): no staging registers and no STS instructions, which is what lets
  // three 128-thread blocks share an SM at BN = 64.  Two buffers: tile kt+1 is in flight while tile kt is multiplied.
  auto issue = [&](int kt, int buf) {
    const double* ap = A + (size_t)(kt + lk) * lda + m0 + lo;
    const double* bp = B + (size_t)(kt + lk) * ldb + n0 + lo;
#pragma unroll
    for (int i = 0; i < NA; ++i) cp_async16(&As[buf][lk][lo + STEP * i], ap + STEP * i);
#pragma unroll
    for (int i = 0; i < NB; ++i) cp_async16(&Bs[buf][lk][lo + STEP * i], bp + STEP * i);
    cp_async_commit();
  };
  int buf = 0;
  if (k_begin < k_end) issue(k_begin, 0);
  for (int kt = k_begin; kt < k_end; kt += GEMM_BK) {
    cp_async_wait_all();      // this thread's part of tile kt has landed ...
    __syncthreads();          // ... and everybody's; everybody is also done reading the other buffer (tile kt - 1)
    if (kt + GEMM_BK < k_end) issue(kt + GEMM_BK, buf ^ 1);
#pragma unroll
    for (int k4 = 0; k4 < GEMM_BK; k4 += 4) {
      double a[4], b[8];
#pragma unroll
      for (int i = 0; i < 4; ++i) a[i] = As[buf][k4 + fk][wm + 8 * i + fr];
#pragma unroll
      for (int j = 0; j < 8; ++j) b[j] = Bs[buf][k4 + fk][wn + 8 * j + fr];
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) dmma_m8n8k4(acc[i][j][0], acc[i][j][1], a[i], b[j], acc[i][j][0], acc[i][j][1]);
    }
    buf ^= 1;
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int m = m0 + wm + 8 * i + fr;
    const bool live = EPI == 0 || m < n_valid_rows;
    const double y = (EPI == 1 && live) ? yvec[m] : 0.0;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int n = n0 + wn + 8 * j + 2 * fk;
      if (EPI == 0) {
        double2 v; v.x = acc[i][j][0]; v.y = acc[i][j][1];
        *reinterpret_cast<double2*>(C0 + (size_t)blockIdx.z * split_stride + (size_t)m * ldc + n) = v;
      } else {
        double2 l, r;
        logreg_terms(acc[i][j][0], y, l.x, r.x);
        logreg_terms(acc[i][j][1], y, l.y, r.y);
        if (!live) { l.x = l.y = 0.0; r.x = r.y = 0.0; }
        *reinterpret_cast<double2*>(C0 + (size_t)m * ldc + n) = l;
        *reinterpret_cast<double2*>(C1 + (size_t)m * ldc + n) = r;
      }
    }
  }
}

// Bernoulli terms, elementwise at full occupancy: Z[m][n] (in LL) -> LL[m][n], RES[m][n];
// rows m >= n_valid_rows are zeroed.  (Fusing this into the GEMM epilogue costs more: the
// GEMM runs 8 warps per SM, so ~150 dependent FP64 instructions per element there stall the
// tensor pipe; measured 0.7 ms fused vs 0.2 ms as a separate pass over 128 MB.)
__global__ void logreg_bernoulli_kernel(double* __restrict__ LL, double* __restrict__ RES, const double* __restrict__ yvec,
                                        int ld, int n_rows_pad, int n_valid_rows, const int* __restrict__ n_cols_ptr, int col_tile) {
  const int ncp = (*n_cols_ptr + col_tile - 1) / col_tile * col_tile;   // the column tiles the GEMM filled
  const int half = ncp / 2;
  const size_t total2 = (size_t)n_rows_pad * half;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total2; i += (size_t)gridDim.x * blockDim.x) {
    const int m = (int)(i / half);
    const size_t e = (size_t)m * ld + 2 * (i - (size_t)m * half);
    const bool live = m < n_valid_rows;
    const double y = live ? yvec[m] : 0.0;
    const double2 z = *reinterpret_cast<const double2*>(LL + e);
    double2 l, r;
    logreg_terms(z.x, y, l.x, r.x);
    logreg_terms(z.y, y, l.y, r.y);
    if (!live) { l.x = l.y = 0.0; r.x = r.y = 0.0; }
    *reinterpret_cast<double2*>(LL + e) = l;
    *reinterpret_cast<double2*>(RES + e) = r;
  }
}

// lik[r] = sum over row tiles (in order) of the canonical 32-lane tree over the tile's rows.
// One warp per chain column; LL is [n_pad][ld].
__global__ void logreg_reduce_ll_kernel(const double* __restrict__ LL, int ld, int n_data, const int* __restrict__ n_cols_ptr,
                                        const int* __restrict__ cols, double* __restrict__ lik) {
  const int lane = threadIdx.x & 31;
  const int j = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (j >= *n_cols_ptr) return;
  double total = 0.0;
  for (int t0 = 0; t0 < n_data; t0 += LR_TILE) {
    const int len = min(LR_TILE, n_data - t0);
    double acc = 0.0;
    for (int i = lane; i < len; i += 32) acc = acc + LL[(size_t)(t0 + i) * ld + j];
    total = total + warp_sum(acc);
  }
  if (lane == 0) lik[cols[j]] = total;
}

// G[r][c] = sum over split-K chunks (in order) of Gp[s][c][r]  (also transposes)
__global__ void logreg_finalize_grad_kernel(const double* __restrict__ Gp, int n_splits, size_t split_stride, int ldp,
                                            int d_pad, const int* __restrict__ n_cols_ptr, const int* __restrict__ cols,
                                            double* __restrict__ G) {
  __shared__ double tile[32][33];
  const int n_cols = *n_cols_ptr;
  const int c0 = blockIdx.x * 32, r0 = blockIdx.y * 32;
  if (r0 >= n_cols) return;
  const int tx = threadIdx.x, ty = threadIdx.y;   // 32 x 8
  for (int i = ty; i < 32; i += 8) {
    const int c = c0 + i, r = r0 + tx;
    double g = 0.0;
    if (c < d_pad && r < n_cols)
      for (int s = 0; s < n_splits; ++s) g = g + Gp[(size_t)s * split_stride + (size_t)c * ldp + r];
    tile[i][tx] = g;
  }
  __syncthreads();
  for (int i = ty; i < 32; i += 8) {
    const int r = r0 + i, c = c0 + tx;
    if (r < n_cols && c < d_pad) G[(size_t)cols[r] * d_pad + c] = tile[tx][i];
  }
}

// Thetat[c][j] = Theta[cols[j]][c] for the compacted columns j < n_cols; the rest of the last column tile is zero-filled
__global__ void logreg_gather_transpose_kernel(const double* __restrict__ src, int cols_dim, int ld_src, const int* __restrict__ n_cols_ptr,
                                               const int* __restrict__ cols, double* __restrict__ dst, int ld_dst, int col_tile) {
  __shared__ double tile[32][33];
  const int n_cols = *n_cols_ptr;
  const int ncp = (n_cols + col_tile - 1) / col_tile * col_tile;
  const int r0 = blockIdx.y * 32, c0 = blockIdx.x * 32;
  if (r0 >= ncp) return;
  const int tx = threadIdx.x, ty = threadIdx.y;
  for (int i = ty; i < 32; i += 8) {
    const int j = r0 + i, c = c0 + tx;
    tile[i][tx] = (j < n_cols && c < cols_dim) ? src[(size_t)cols[j] * ld_src + c] : 0.0;
  }
  __syncthreads();
  for (int i = ty; i < 32; i += 8) {
    const int c = c0 + i, j = r0 + tx;
    if (c < cols_dim) dst[(size_t)c * ld_dst + j] = tile[tx][i];
  }
}

// The chains whose state machine emitted a point in this batch step, in chain order: cols[0..n_cols).  One block.
// Also keeps the round's counters (batch steps that evaluated something, columns requested, columns multiplied)
// and the per-step history the host reads once per chunk of steps.
__global__ void logreg_compact_kernel(const LrChainState* __restrict__ st, int n_local, int* __restrict__ cols, LrControl* ctl,
                                      int step_in_chunk, int col_tile) {
  __shared__ int warp_count[32];
  __shared__ int base;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, n_warps = blockDim.x >> 5;
  if (tid == 0) base = 0;
  __syncthreads();
  for (int r0 = 0; r0 < n_local; r0 += blockDim.x) {
    const int r = r0 + tid;
    const bool want = r < n_local && st[r].phase != LR_DONE && st[r].phase != LR_SCAN_START && st[r].err == 0;
    const unsigned int m = __ballot_sync(PGN_FULL_MASK, want);
    if (lane == 0) warp_count[warp] = __popc(m);
    __syncthreads();
    int off = base;
    for (int w = 0; w < warp; ++w) off += warp_count[w];
    if (want) cols[off + __popc(m & ((1u << lane) - 1u))] = r;
    __syncthreads();
    if (tid == 0) { int t = 0; for (int w = 0; w < n_warps; ++w) t += warp_count[w]; base += t; }
    __syncthreads();
  }
  if (tid == 0) {
    const int n = base;
    ctl->n_cols = n;
    ctl->hist[step_in_chunk] = n;
    if (n > 0) {
      ctl->steps += 1;
      ctl->sum_active += n;
      ctl->sum_gemm_cols += (n + col_tile - 1) / col_tile * col_tile;
    }
  }
}

// Thetat[c][r] = Theta[r][c]
__global__ void logreg_transpose_kernel(const double* __restrict__ src, int rows, int cols, int ld_src,
                                        double* __restrict__ dst, int ld_dst) {
  __shared__ double tile[32][33];
  const int r0 = blockIdx.y * 32, c0 = blockIdx.x * 32;
  const int tx = threadIdx.x, ty = threadIdx.y;
  for (int i = ty; i < 32; i += 8) {
    const int r = r0 + i, c = c0 + tx;
    tile[i][tx] = (r < rows && c < cols) ? src[(size_t)r * ld_src + c] : 0.0;
  }
  __syncthreads();
  for (int i = ty; i < 32; i += 8) {
    const int c = c0 + i, r = r0 + tx;
    if (c < cols && r < rows) dst[(size_t)c * ld_dst + r] = tile[tx][i];
  }
}

// ---------------------------------------------------------------------------
// Per-chain controller: the autoMALA / MALA state machine of VecChain, one density
// evaluation per batch step, vectors in HBM ([chain][d_pad], lane-strided access).
// ---------------------------------------------------------------------------
struct LrCtx {
  const LrParams& P;
  LrChainState& s;
  int lane, r, nslots;
  double beta;
  Rng rng;
  __device__ LrCtx(const LrParams& P_, LrChainState& s_, int lane_, int r_)
      : P(P_), s(s_), lane(lane_), r(r_), nslots(P_.d_pad / 32) {
    beta = P.beta[P.first_chain + r - 1];
    rng.key0 = P.seed_lo; rng.key1 = (unsigned int)s.replica_index; rng.c2 = P.seed_hi; rng.c3 = 0u;
    rng.ctr = s.ctr;
  }
  __device__ __forceinline__ double* row(double* base) const { return base + (size_t)r * P.d_pad; }
  __device__ __forceinline__ const double* row(const double* base) const { return base + (size_t)r * P.d_pad; }
  __device__ __forceinline__ bool valid(int c) const { return c < P.d; }
  __device__ __forceinline__ double pre_at(int c) const {
    if (s.pre_mode == 0) return 1.0;
    const double sd = P.std_devs[c];
    if (sd == 0.0) return 1.0;
    return s.pre_mode == 1 ? 1.0 / sd : s.mix + s.rmix / sd;
  }
  __device__ double ref_density(const double* v) const {
    double acc = 0.0;
    for (int k = 0; k < nslots; ++k) {
      const int c = k * 32 + lane;
      if (valid(c)) acc = acc + (-(v[c] * v[c] * P.iv_ref + PGN_LOG2PI) * 0.5 - P.ls_ref);
    }
    return warp_sum(acc);
  }
  __device__ double sqr_norm(const double* v) const {
    double acc = 0.0;
    for (int k = 0; k < nslots; ++k) {
      const int c = k * 32 + lane;
      if (valid(c)) acc = acc + v[c] * v[c];
    }
    return warp_sum(acc);
  }
  __device__ __forceinline__ double lp_ad(double a0, double a1) const { return (1.0 - beta) * a0 + beta * a1; }
  // conditioned gradient at point v with likelihood gradient gl: out = (g_ref (1-b) + g_tgt b) / pre
  __device__ void cond_grad(const double* v, const double* gl, double* out) const {
    for (int k = 0; k < nslots; ++k) {
      const int c = k * 32 + lane;
      if (!valid(c)) { out[c] = 0.0; continue; }
      const double gr = -v[c] * P.iv_ref;
      const double gt = -v[c] * P.iv_ref + gl[c];
      const double t = gr * (1.0 - beta);
      const double g = t + gt * beta;
      out[c] = s.pre_mode == 0 ? g : g / pre_at(c);
    }
  }
  // emit the next trial point TX = SX + eps * (ph / pre), ph = SP + (eps/2) SG; s.pp = |ph|^2
  __device__ void emit_trial() {
    const double eps = s.eps, half_eps = eps / 2;
    const double *sx = row(P.SX), *sp = row(P.SP), *sg = row(P.SG);
    double* tx = row(P.TX);
    double acc = 0.0;
    for (int k = 0; k < nslots; ++k) {
      const int c = k * 32 + lane;
      if (!valid(c)) { tx[c] = 0.0; continue; }
      const double ph = sp[c] + half_eps * sg[c];
      tx[c] = sx[c] + eps * (s.pre_mode == 0 ? ph : ph / pre_at(c));
      acc = acc + ph * ph;
    }
    s.pp = warp_sum(acc);
    s.phase = LR_WAIT_TRIAL;
    s.n_points += 1;
  }
  __device__ void copy(double* dst, const double* src, double scale_sign = 1.0) const {
    for (int k = 0; k < nslots; ++k) {
      const int c = k * 32 + lane;
      dst[c] = scale_sign == 1.0 ? src[c] : src[c] * scale_sign;
    }
  }
  __device__ void build_preconditioner() {   // Preconditioner.jl:57-77
    if (P.std_devs == nullptr || P.precond_kind == PGN_PRECOND_IDENTITY) { s.pre_mode = 0; return; }
    if (P.precond_kind == PGN_PRECOND_DIAGONAL) { s.pre_mode = 1; return; }
    const double u = next_uniform(rng);
    if (u <= P.mix_p0) s.pre_mode = 1;
    else if (u <= P.mix_p01) s.pre_mode = 0;
    else { s.pre_mode = 2; s.mix = next_uniform(rng); s.rmix = 1.0 - s.mix; }
  }
  // begin refresh: momentum, bounds, first trial of the forward search
  __device__ bool begin_refresh(bool use_mh) {
    if (s.refresh_i >= P.n_refresh) { s.phase = LR_DONE; return false; }
    double* p = row(P.P);
    double acc = 0.0;
    for (int k = 0; k < nslots; ++k) {
      const int c = k * 32 + lane;
      const double z = valid(c) ? normal_at(rng, rng.ctr + (unsigned long long)c) : 0.0;
      p[c] = z;
      if (valid(c)) acc = acc + z * z;
    }
    rng.ctr += (unsigned long long)P.d;
    s.init_joint = s.lp0 - 0.5 * warp_sum(acc);
    if (!is_finite(s.init_joint)) { s.err = PGN_ERR_NOT_POSITIVE; return false; }
    if (P.explorer_kind == PGN_EXPLORER_AUTOMALA) {
      double mine = uniform_at(rng, rng.ctr + (unsigned long long)(lane < 3 ? lane : 0));
      double lmine = log_(mine);
      rng.ctr += use_mh ? 3ull : 2ull;
      const double a = __shfl_sync(PGN_FULL_MASK, mine, 0), b = __shfl_sync(PGN_FULL_MASK, mine, 1);
      const double la = __shfl_sync(PGN_FULL_MASK, lmine, 0), lb = __shfl_sync(PGN_FULL_MASK, lmine, 1);
      s.u_mh = __shfl_sync(PGN_FULL_MASK, mine, 2);
      s.lower = a < b ? la : lb;
      s.upper = a < b ? lb : la;
      if (!(s.lower < s.upper)) { s.err = PGN_ERR_INVALID; return false; }
    }
    copy(row(P.SX), row(P.X));
    copy(row(P.SP), row(P.P));
    copy(row(P.SG), row(P.G0));
    s.dir = 0; s.mode = 0; s.n = 0; s.exponent = 0; s.nst = 0;
    s.eps = P.step_size;
    s.h_before = s.init_joint;
    emit_trial();
    return true;
  }
  // consume the evaluation of TX; returns with the next request emitted or the scan finished
  __device__ void on_trial(bool use_mh) {
    const double* tx = row(P.TX);
    s.t_a0 = ref_density(tx);
    s.t_a1 = s.t_a0 + P.lik[r];
    s.t_lp1 = lp_ad(s.t_a0, s.t_a1);
    double* tg = row(P.TG);
    cond_grad(tx, row(P.G), tg);
    const double eps = s.eps, half_eps = eps / 2;
    const double *sp = row(P.SP), *sg = row(P.SG);
    double* tp = row(P.TP);
    const double cur = s.t_lp1 - 0.5 * s.pp;
    double s2;
    if (!is_finite(cur)) {
      for (int k = 0; k < nslots; ++k) { const int c = k * 32 + lane; tp[c] = valid(c) ? sp[c] + half_eps * sg[c] : 0.0; }
      s2 = s.pp;
    } else {
      double acc = 0.0;
      for (int k = 0; k < nslots; ++k) {
        const int c = k * 32 + lane;
        if (!valid(c)) { tp[c] = 0.0; continue; }
        const double ph = sp[c] + half_eps * sg[c];
        const double p1 = ph + half_eps * tg[c];
        tp[c] = p1;
        acc = acc + p1 * p1;
      }
      s2 = warp_sum(acc);
    }
    s.t_h_after = s.t_lp1 - 0.5 * s2;
    s.t_eps = eps;
    const double diff = s.t_h_after - s.h_before;

    if (P.explorer_kind == PGN_EXPLORER_MALA) {   // MALA.jl:74-97
      const double e = exp_(s.t_h_after - s.init_joint);
      const double prob = 1.0 < e ? 1.0 : e;
      s.expl_acc.fit(prob);
      s.n_ref += 4;
      if (next_uniform(rng) < prob) {
        copy(row(P.X), tx); copy(row(P.G0), tg);
        s.e0 = s.t_a0; s.e1 = s.t_a1; s.lp0 = s.t_lp1;
      }
      s.n_steps += 1;
      s.refresh_i += 1;
      begin_refresh(use_mh);
      return;
    }

    // ---- autoMALA step-size search (same transitions as VecChain::automala).  grow_step_size :216-226 ends one
    // step back from the candidate that stopped it, so a growing search keeps its previous trial and returns
    // to it instead of evaluating that step a second time; the reversed search (:160-163) only needs its
    // exponent, so it never re-evaluates.
    bool decided = false;
    bool keep = false;   // the search goes on growing: this trial becomes "the previous one"
    if (s.mode == 0) {
      if (!is_finite(diff) || diff < s.lower) { s.mode = 1; s.n = 1; s.eps = eps / 2.0; }
      else if (diff > s.upper) { s.mode = 2; s.n = 1; s.eps = eps * 2.0; keep = true; }
      else decided = true;
    } else if (s.mode == 1) {
      if (eps == 0.0) { s.err = PGN_ERR_STEP_UNDERFLOW; return; }
      if (diff > s.lower) { s.nst = s.n; s.exponent = -s.n; decided = true; }
      else { s.n += 1; s.eps = eps / 2.0; }
    } else if (s.mode == 2) {
      if (!is_finite(diff) || diff < s.upper) { s.nst = s.n; s.exponent = s.n - 1; decided = true; }
      else { s.n += 1; s.eps = eps * 2.0; keep = true; }
    } else {
      decided = true;   // mode 3: re-evaluated at the chosen step
    }
    if (!decided) {
      if (keep && s.dir == 0) {
        copy(row(P.QX), tx); copy(row(P.QP), tp); copy(row(P.QG), tg);
        s.q_a0 = s.t_a0; s.q_a1 = s.t_a1; s.q_lp1 = s.t_lp1; s.q_h_after = s.t_h_after; s.q_eps = s.t_eps;
      }
      emit_trial();
      return;
    }
    if (s.mode != 3 && s.dir == 0) {
      const double eps_final = P.step_size * pow2(s.exponent);
      if (s.t_eps != eps_final) {
        if (s.mode == 2 && s.q_eps == eps_final) {   // back to the previous trial: it is the leap_frog! at the chosen step :144-151
          double* txw = row(P.TX);
          copy(txw, row(P.QX)); copy(tp, row(P.QP)); copy(tg, row(P.QG));
          s.t_a0 = s.q_a0; s.t_a1 = s.q_a1; s.t_lp1 = s.q_lp1; s.t_h_after = s.q_h_after; s.t_eps = s.q_eps;
        } else {
          s.mode = 3; s.eps = eps_final; emit_trial(); return;
        }
      }
    }
    // search finished
    s.n_steps += 1 + s.nst;
    s.am.fit(pow2(s.exponent));
    if (s.dir == 0) {
      s.n_ref += 1 + 1 + 3 * (1 + s.nst) + 2;
      s.expo0 = s.exponent;
      s.h_rev = s.t_h_after;
      s.f_a0 = s.t_a0; s.f_a1 = s.t_a1; s.f_lp = s.t_lp1;
      copy(row(P.FX), tx);
      copy(row(P.FG), tg);
      if (use_mh) {
        copy(row(P.SX), tx);
        copy(row(P.SP), tp, -1.0);
        copy(row(P.SG), tg);
        s.dir = 1; s.mode = 0; s.n = 0; s.exponent = 0; s.nst = 0;
        s.eps = P.step_size;
        s.h_before = s.h_rev;
        emit_trial();
        return;
      }
    } else {
      s.n_ref += 1 + 3 * (1 + s.nst);
    }
    bool accept = true;
    if (use_mh) {
      const bool passed = (s.exponent == s.expo0);
      s.rev.fit(passed ? 1.0 : 0.0);
      double prob = 0.0;
      if (passed) { const double e = exp_(s.h_rev - s.init_joint); prob = 1.0 < e ? 1.0 : e; s.n_ref += 1; }
      s.expl_acc.fit(prob);
      accept = s.u_mh < prob;
    }
    if (accept) {
      copy(row(P.X), row(P.FX));
      copy(row(P.G0), row(P.FG));
      s.e0 = s.f_a0; s.e1 = s.f_a1; s.lp0 = s.f_lp;
    }
    s.refresh_i += 1;
    begin_refresh(use_mh);
  }
};

// one warp per chain
__global__ void logreg_controller_kernel(const __grid_constant__ LrParams P) {
  const int lane = threadIdx.x & 31;
  const int r = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (r >= P.n_local) return;
  LrChainState s = P.st[r];
  if (s.phase == LR_DONE || s.err != 0) return;
  const int chain = P.first_chain + r;
  const bool is_ref = (chain == 1 && P.n_chains > 1);
  const bool use_mh = (P.scan != 1);
  LrCtx cx(P, s, lane, r);
  if (s.phase == LR_SCAN_START) {
    if (is_ref) {   // sample_iid! from N(0, sigma_ref^2 I), then densities at the new state
      double* x = cx.row(P.X);
      for (int k = 0; k < cx.nslots; ++k) {
        const int c = k * 32 + lane;
        x[c] = c < P.d ? P.sigma_ref * normal_at(cx.rng, cx.rng.ctr + (unsigned long long)c) : 0.0;
      }
      cx.rng.ctr += (unsigned long long)P.d;
    } else {
      cx.build_preconditioner();
    }
    cx.copy(cx.row(P.TX), cx.row(P.X));
    s.n_points += 1;
    s.phase = LR_WAIT_X0;
  } else if (s.phase == LR_WAIT_X0) {
    const double* x = cx.row(P.X);
    s.e0 = cx.ref_density(x);
    s.e1 = s.e0 + P.lik[r];
    if (is_ref) {
      s.phase = LR_DONE;
    } else {
      cx.cond_grad(x, cx.row(P.G), cx.row(P.G0));
      s.lp0 = cx.lp_ad(s.e0, s.e1);
      if (!(P.step_size > 0)) s.err = PGN_ERR_INVALID;
      s.refresh_i = 0;
      if (s.err == 0) cx.begin_refresh(use_mh);
    }
  } else {
    cx.on_trial(use_mh);
  }
  s.ctr = cx.rng.ctr;
  if (lane == 0) {
    P.st[r] = s;
    if (s.err != 0) atomicCAS(P.error_flag, 0, s.err);
  }
}

// ---- swap phase: post, then decide (same mailbox protocol as scan_kernel) ---------------
__global__ void logreg_post_kernel(const __grid_constant__ LrParams P) {
  const int lane = threadIdx.x & 31;
  const int r = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (r >= P.n_local) return;
  LrChainState s = P.st[r];
  const int N = P.n_chains, chain = P.first_chain + r, last_local = P.first_chain + P.n_local - 1;
  const double beta = P.beta[chain - 1];
  const bool is_ref = (chain == 1 && N > 1), is_tgt = (chain == N);
  const double* x = P.X + (size_t)r * P.d_pad;
  if (is_tgt) {   // target-chain recording (pigeons.jl:110-131)
    const long long n = *P.online_n + 1;
    __syncwarp();
    const double g = 1.0 / (double)n;
    for (int c = lane; c < P.d; c += 32) {
      const double mu_old = P.online_mean[c];
      const double mu = mu_old + g * (x[c] - mu_old);
      P.online_s2[c] = P.online_s2[c] + g * ((x[c] - mu) * (x[c] - mu_old) - P.online_s2[c]);
      P.online_mean[c] = mu;
    }
    __syncwarp();
    if (lane == 0) *P.online_n = n;
    if (P.target_trace)
      for (int c = lane; c < P.d; c += 32) P.target_trace[(size_t)(P.scan - 1) * P.d + c] = x[c];
  }
  const bool even = (P.scan & 1LL) == 0;
  int partner = chain + ((((chain & 1) == 0) == even) ? 1 : -1);
  if (partner == 0) partner = 1;
  if (partner == N + 1) partner = N;
  auto lp_call = [&](double b) { return b == 0.0 ? s.e0 : (b == 1.0 ? s.e1 : (1.0 - b) * s.e0 + b * s.e1); };
  const double lr = lp_call(P.beta[partner - 1]) - lp_call(beta);
  s.n_ref += 2;
  if (lr != lr) s.err = PGN_ERR_NAN_RATIO;
  Rng rng{P.seed_lo, (unsigned int)s.replica_index, P.seed_hi, 0u, s.ctr};
  const double u = next_uniform(rng);
  s.ctr = rng.ctr;
  s.lr = lr; s.u = u; s.accepted = 0;
  const size_t log_at = (size_t)(P.scan - 1) * P.n_local + r;
  if (lane == 0) {
    if (P.index_process) P.index_process[log_at] = s.replica_index;
    if (P.swap_lr) P.swap_lr[log_at] = lr;
    if (P.swap_u) P.swap_u[log_at] = u;
  }
  if (s.rt_state == 0 && is_ref) s.rt_state = 1;
  else if (s.rt_state == 1 && is_tgt) { s.rt_state = 2; s.n_restarts += 1; }
  else if (s.rt_state == 2 && is_ref) { s.rt_state = 1; s.n_trips += 1; }
  if (partner != chain && s.err == 0) {
    const int ring = (int)((P.epoch & 1u) * 4u + (unsigned int)(P.scan & 3LL));
    const unsigned long long tag = ((unsigned long long)P.epoch << 32) | (unsigned long long)P.scan;
    const bool remote = partner < P.first_chain || partner > last_local;
    char* dst;
    if (!remote) dst = P.mail + ((size_t)(2 + r) * MAIL_RINGS + ring) * P.slot_bytes;
    else if (partner > chain) dst = P.mail_right + ((size_t)0 * MAIL_RINGS + ring) * P.slot_bytes;
    else dst = P.mail_left + ((size_t)1 * MAIL_RINGS + ring) * P.slot_bytes;
    if (lane == 0) {
      MailHdr* h = reinterpret_cast<MailHdr*>(dst + 32);
      h->lr = lr; h->u = u; h->ctr = s.ctr; h->replica_index = s.replica_index; h->rt_state = s.rt_state;
    }
    double* pay = reinterpret_cast<double*>(dst + MAIL_HDR_BYTES);
    for (int c = lane; c < P.d_pad; c += 32) pay[c] = x[c];
    if (lane == 0) { pay[P.d_pad] = s.e0; pay[P.d_pad + 1] = s.e1; }
    __syncwarp();
    if (lane == 0) {
      if (remote) st_release_sys(reinterpret_cast<unsigned long long*>(dst), tag);
      else st_release_gpu(reinterpret_cast<unsigned long long*>(dst), tag);
    }
  }
  if (lane == 0) {
    P.st[r] = s;
    if (s.err != 0) atomicCAS(P.error_flag, 0, s.err);
  }
}

__global__ void logreg_decide_kernel(const __grid_constant__ LrParams P) {
  const int lane = threadIdx.x & 31;
  const int r = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (r >= P.n_local) return;
  LrChainState s = P.st[r];
  const int N = P.n_chains, chain = P.first_chain + r, last_local = P.first_chain + P.n_local - 1;
  const bool even = (P.scan & 1LL) == 0;
  int partner = chain + ((((chain & 1) == 0) == even) ? 1 : -1);
  if (partner == 0) partner = 1;
  if (partner == N + 1) partner = N;
  const size_t log_at = (size_t)(P.scan - 1) * P.n_local + r;
  bool accepted = false;
  int err = 0;
  if (partner != chain) {
    const int ring = (int)((P.epoch & 1u) * 4u + (unsigned int)(P.scan & 3LL));
    const unsigned long long tag = ((unsigned long long)P.epoch << 32) | (unsigned long long)P.scan;
    const bool remote = partner < P.first_chain || partner > last_local;
    const char* src;
    if (!remote) src = P.mail + ((size_t)(2 + (partner - P.first_chain)) * MAIL_RINGS + ring) * P.slot_bytes;
    else if (partner > chain) src = P.mail + ((size_t)1 * MAIL_RINGS + ring) * P.slot_bytes;
    else src = P.mail + ((size_t)0 * MAIL_RINGS + ring) * P.slot_bytes;
    int status = 0;
    if (lane == 0) {
      const unsigned long long* flag = reinterpret_cast<const unsigned long long*>(src);
      unsigned long long t0 = 0;
      unsigned int it = 0;
      while (ld_relaxed_sys(flag) != tag) {
        ++it;
        if ((it & 255u) == 0u) {
          if (*reinterpret_cast<volatile int*>(P.error_flag) != 0) { status = 1; break; }
          const unsigned long long now = globaltimer_ns();
          if (t0 == 0) t0 = now;
          else if (now - t0 > P.timeout_ns) { status = 2; break; }
          __nanosleep(1000);
        }
      }
      if (remote) fence_acq_rel_sys(); else fence_acq_rel_gpu();
    }
    status = __shfl_sync(PGN_FULL_MASK, status, 0);
    if (status != 0) {
      err = status == 2 ? PGN_ERR_TIMEOUT : -1;
    } else {
      const MailHdr* hp = reinterpret_cast<const MailHdr*>(src + 32);
      const double lr_p = __ldcg(&hp->lr), u_p = __ldcg(&hp->u);
      const bool lower = chain < partner;
      const double e = lower ? exp_(s.lr + lr_p) : exp_(lr_p + s.lr);
      const double acceptance_pr = 1.0 < e ? 1.0 : e;
      if (lower) { s.swap_acc.fit(acceptance_pr); s.ls_fwd.fit(s.lr); s.ls_bwd.fit(lr_p); }
      accepted = (lower ? s.u : u_p) < acceptance_pr;
      if (accepted) {
        const int ri_new = __ldcg(&hp->replica_index);
        if (P.rec_table != nullptr) {   // per-replica recorders: see scan_kernel
          RecEntry* eo = P.rec_table + (size_t)(s.replica_index - 1) * P.n_local + r;
          const RecEntry* en = P.rec_table + (size_t)(ri_new - 1) * P.n_local + r;
          if (lane == 0) {
            eo->expl_acc = s.expl_acc; eo->am = s.am; eo->rev = s.rev; eo->swap_acc = s.swap_acc;
            eo->ls_fwd = s.ls_fwd; eo->ls_bwd = s.ls_bwd;
          }
          s.expl_acc = en->expl_acc; s.am = en->am; s.rev = en->rev; s.swap_acc = en->swap_acc;
          s.ls_fwd = en->ls_fwd; s.ls_bwd = en->ls_bwd;
          if (chain == N) {
            OnEntry* oo = P.on_table + (size_t)(s.replica_index - 1) * P.d_pad;
            const OnEntry* on = P.on_table + (size_t)(ri_new - 1) * P.d_pad;
            const long long n_old = *P.online_n;
            __syncwarp();
            for (int c = lane; c < P.d; c += 32) {
              oo[c] = OnEntry{n_old, P.online_mean[c], P.online_s2[c]};
              const OnEntry e = on[c];
              P.online_mean[c] = e.mu; P.online_s2[c] = e.s2;
            }
            __syncwarp();
            if (lane == 0) *P.online_n = on[0].n;
          }
        }
        s.replica_index = ri_new;
        s.rt_state = __ldcg(&hp->rt_state);
        s.ctr = __ldcg(&hp->ctr);
        const double* pay = reinterpret_cast<const double*>(src + MAIL_HDR_BYTES);
        double* x = P.X + (size_t)r * P.d_pad;
        for (int c = lane; c < P.d_pad; c += 32) x[c] = __ldcg(pay + c);
        s.e0 = __ldcg(pay + P.d_pad);
        s.e1 = __ldcg(pay + P.d_pad + 1);
      }
    }
  }
  s.accepted = accepted ? 1 : 0;
  s.phase = LR_SCAN_START;   // next scan
  if (lane == 0) {
    if (P.swap_accept) P.swap_accept[log_at] = accepted ? 1 : 0;
    P.st[r] = s;
    if (err > 0) atomicCAS(P.error_flag, 0, err);
  }
}

// per-replica recorders: when the round ends every chain hands the statistics it holds to its current replica's entry
__global__ void logreg_flush_recorders_kernel(const __grid_constant__ LrParams P) {
  const int lane = threadIdx.x & 31;
  const int r = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (r >= P.n_local || P.rec_table == nullptr) return;
  const LrChainState s = P.st[r];
  RecEntry* eo = P.rec_table + (size_t)(s.replica_index - 1) * P.n_local + r;
  if (lane == 0) {
    eo->expl_acc = s.expl_acc; eo->am = s.am; eo->rev = s.rev; eo->swap_acc = s.swap_acc;
    eo->ls_fwd = s.ls_fwd; eo->ls_bwd = s.ls_bwd;
  }
  if (P.first_chain + r == P.n_chains) {
    OnEntry* oo = P.on_table + (size_t)(s.replica_index - 1) * P.d_pad;
    const long long n = *P.online_n;
    for (int c = lane; c < P.d; c += 32) oo[c] = OnEntry{n, P.online_mean[c], P.online_s2[c]};
  }
}

// one warp per point: reference density and the combination used by the parity entry points
__global__ void logreg_points_finish_kernel(const double* __restrict__ xs, int d, int d_pad, int n_points,
                                            const double* __restrict__ betas, const double* __restrict__ lik,
                                            const double* __restrict__ G, double iv_ref, double ls_ref,
                                            double* lp_out, double* ld_out, double* grad_out) {
  const int lane = threadIdx.x & 31;
  const int w = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (w >= n_points) return;
  const double* x = xs + (size_t)w * d_pad;
  double acc = 0.0;
  for (int c = lane; c < d; c += 32) acc = acc + (-(x[c] * x[c] * iv_ref + PGN_LOG2PI) * 0.5 - ls_ref);
  const double a0 = warp_sum(acc);
  const double a1 = a0 + lik[w];
  const double b = betas[w];
  if (lp_out && lane == 0) lp_out[w] = b == 0.0 ? a0 : (b == 1.0 ? a1 : (1.0 - b) * a0 + b * a1);
  if (ld_out) {
    if (lane == 0) ld_out[w] = (0.0 + a0 * (1.0 - b)) + a1 * b;
    for (int c = lane; c < d; c += 32) {
      const double gr = -x[c] * iv_ref;
      const double gt = -x[c] * iv_ref + G[(size_t)w * d_pad + c];
      const double t = gr * (1.0 - b);
      grad_out[(size_t)w * d + c] = t + gt * b;
    }
  }
}

// D = A B + C with one mma.sync.m8n8k4.f64 per warp (fragment layout of the PTX ISA:
// a: row lane/4, col lane%4; b: row lane%4, col lane/4; c/d: row lane/4, cols 2*(lane%4)+{0,1})
__device__ __forceinline__ void dmma_m8n8k4(double& d0, double& d1, double a, double b, double c0, double c1) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%4,%5};"
               : "=d"(d0), "=d"(d1) : "d"(a), "d"(b), "d"(c0), "d"(c1));
}
__global__ void dmma_probe_kernel(const double* A, const double* B, const double* Cm, double* D, int n_trials) {
  const int lane = threadIdx.x & 31;
  const int t = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (t >= n_trials) return;
  const double a = A[(size_t)t * 32 + (lane >> 2) * 4 + (lane & 3)];
  const double b = B[(size_t)t * 32 + (lane & 3) * 8 + (lane >> 2)];
  const double c0 = Cm[(size_t)t * 64 + (lane >> 2) * 8 + 2 * (lane & 3)];
  const double c1 = Cm[(size_t)t * 64 + (lane >> 2) * 8 + 2 * (lane & 3) + 1];
  double d0, d1;
  dmma_m8n8k4(d0, d1, a, b, c0, c1);
  D[(size_t)t * 64 + (lane >> 2) * 8 + 2 * (lane & 3)] = d0;
  D[(size_t)t * 64 + (lane >> 2) * 8 + 2 * (lane & 3) + 1] = d1;
}

// 16 independent DFMA chains per thread, register resident: FP64 FMA peak probe
__global__ void fp64_peak_kernel(double* sink, int iters, double m) {
  double a[16];
#pragma unroll
  for (int i = 0; i < 16; ++i) a[i] = 1.0 + 1e-9 * (threadIdx.x + i);
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 16; ++i) a[i] = fma(a[i], m, 1e-12);
  }
  double s = 0.0;
#pragma unroll
  for (int i = 0; i < 16; ++i) s += a[i];
  if (s == 123.456) *sink = s;
}

}  // namespace pgn
