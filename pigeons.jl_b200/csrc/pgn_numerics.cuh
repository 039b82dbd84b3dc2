// pgn_numerics.cuh — device arithmetic of the PT scan engine (sm_100a).
//
// Elementary functions built from IEEE-754 basic operations only (+ - * /
// sqrt fma are correctly rounded in fp64 on the GPU), a Philox4x32-10 counter
// RNG with one block per draw, and the canonical warp summation tree.  The
// library is compiled with -fmad=false: the only fused operations are the
// explicit fma() calls below, so results are reproducible bit for bit on any
// IEEE machine (the test-suite checks this against an independent CPU
// restatement).
//
// Role in the reference: stands in for Julia's libm (exp/log), the `Random`
// stdlib (rand/randn/randexp: src/explorers/SliceSampler.jl:91,110,130,188,
// src/explorers/AutoMALA.jl:125,132-133,173, src/swap/pair_swapper.jl:46) and
// SplittableRandoms.jl (one independent stream per replica,
// src/replicas/replicas.jl:87-99).
#pragma once
#include <cstdint>

namespace pgn {

#define PGN_FULL_MASK 0xffffffffu

__device__ __forceinline__ double bits_to_double(unsigned long long b) { return __longlong_as_double((long long)b); }
__device__ __forceinline__ unsigned long long double_to_bits(double d) { return (unsigned long long)__double_as_longlong(d); }

#define PGN_INF bits_to_double(0x7ff0000000000000ULL)
#define PGN_NAN bits_to_double(0x7ff8000000000000ULL)
#define PGN_LN2_HI bits_to_double(0x3fe62e42fee00000ULL)
#define PGN_LN2_LO bits_to_double(0x3dea39ef35793c76ULL)
#define PGN_INV_LN2 bits_to_double(0x3ff71547652b82feULL)
#define PGN_EXP_OVERFLOW bits_to_double(0x40862e42fefa39efULL)
#define PGN_EXP_UNDERFLOW bits_to_double(0xc0874910d52d3051ULL)
#define PGN_PI bits_to_double(0x400921fb54442d18ULL)
#define PGN_LOG2PI bits_to_double(0x3ffd67f1c864beb5ULL)
#define PGN_TWO_M52 bits_to_double(0x3cb0000000000000ULL)

__device__ __forceinline__ bool is_finite(double x) {
  return (double_to_bits(x) & 0x7ff0000000000000ULL) != 0x7ff0000000000000ULL;
}

__device__ __forceinline__ double pow2i(int e) { return bits_to_double((unsigned long long)(e + 1023) << 52); }

__device__ __forceinline__ double scale2(double x, int k) {
  int k1 = k >> 1;
  int k2 = k - k1;
  return (x * pow2i(k1)) * pow2i(k2);
}

// 2.0^e for any int e (exact; 0 / inf outside the double range)
__device__ __forceinline__ double pow2(int e) {
  if (e > 1023) return PGN_INF;
  if (e < -1074) return 0.0;
  if (e >= -1022) return pow2i(e);
  return scale2(1.0, e);
}

// exp(x): Cody-Waite reduction, degree-13 Taylor polynomial, Horner with fma.
// Special cases are patched in with selects after the main path (no branches on
// the critical path); the main path is harmless on NaN/inf inputs.
__device__ __forceinline__ double exp_(double x) {
  double kf = rint(x * PGN_INV_LN2);
  double r = fma(-kf, PGN_LN2_HI, x);
  r = fma(-kf, PGN_LN2_LO, r);
  double p = 1.0 / 6227020800.0;
  p = fma(p, r, 1.0 / 479001600.0);
  p = fma(p, r, 1.0 / 39916800.0);
  p = fma(p, r, 1.0 / 3628800.0);
  p = fma(p, r, 1.0 / 362880.0);
  p = fma(p, r, 1.0 / 40320.0);
  p = fma(p, r, 1.0 / 5040.0);
  p = fma(p, r, 1.0 / 720.0);
  p = fma(p, r, 1.0 / 120.0);
  p = fma(p, r, 1.0 / 24.0);
  p = fma(p, r, 1.0 / 6.0);
  p = fma(p, r, 0.5);
  p = fma(p, r, 1.0);
  p = fma(p, r, 1.0);
  int k = (int)kf;
  k = k > 1100 ? 1100 : (k < -1100 ? -1100 : k);
  double res = scale2(p, k);
  res = x > PGN_EXP_OVERFLOW ? PGN_INF : res;
  res = x < PGN_EXP_UNDERFLOW ? 0.0 : res;
  res = x != x ? x : res;
  return res;
}

// log(x): x = 2^k (1+f), s = f/(2+f), log(1+f) = f - hfsq + s (hfsq + R(s^2)).
// Branch-free main path; special cases selected at the end.
__device__ __forceinline__ double log_(double x) {
  const double x_in = x;
  unsigned long long ix = double_to_bits(x);
  const bool sub = ix < 0x0010000000000000ULL;   // positive subnormal (or +0)
  const double xs = x * bits_to_double(0x4350000000000000ULL);
  ix = sub ? double_to_bits(xs) : ix;
  int k = sub ? -54 : 0;
  unsigned int hx = (unsigned int)(ix >> 32);
  k += (int)(hx >> 20) - 1023;
  hx &= 0x000fffffu;
  unsigned int i = (hx + 0x95f64u) & 0x100000u;
  unsigned long long hi = (unsigned long long)(hx | (i ^ 0x3ff00000u));
  x = bits_to_double((hi << 32) | (ix & 0xffffffffULL));
  k += (int)(i >> 20);
  const double LG1 = bits_to_double(0x3fe5555555555593ULL);
  const double LG2 = bits_to_double(0x3fd999999997fa04ULL);
  const double LG3 = bits_to_double(0x3fd2492494229359ULL);
  const double LG4 = bits_to_double(0x3fcc71c51d8e78afULL);
  const double LG5 = bits_to_double(0x3fc7466496cb03deULL);
  const double LG6 = bits_to_double(0x3fc39a09d078c69fULL);
  const double LG7 = bits_to_double(0x3fc2f112df3e5244ULL);
  double f = x - 1.0;
  double hfsq = 0.5 * f * f;
  double s = f / (2.0 + f);
  double z = s * s;
  double w = z * z;
  double t1 = w * fma(w, fma(w, LG6, LG4), LG2);
  double t2 = z * fma(w, fma(w, fma(w, LG7, LG5), LG3), LG1);
  double R = t2 + t1;
  double dk = (double)k;
  double res = s * (hfsq + R) + dk * PGN_LN2_LO - hfsq + f + dk * PGN_LN2_HI;
  res = x_in == PGN_INF ? PGN_INF : res;
  res = x_in == 0.0 ? -PGN_INF : res;
  res = x_in < 0.0 ? PGN_NAN : res;
  res = x_in != x_in ? x_in : res;
  return res;
}

__device__ __forceinline__ double sin_kernel(double y) {
  double z = y * y;
  double p = -1.0 / 1307674368000.0;
  p = fma(p, z, 1.0 / 6227020800.0);
  p = fma(p, z, -1.0 / 39916800.0);
  p = fma(p, z, 1.0 / 362880.0);
  p = fma(p, z, -1.0 / 5040.0);
  p = fma(p, z, 1.0 / 120.0);
  p = fma(p, z, -1.0 / 6.0);
  return fma(y * z, p, y);
}
__device__ __forceinline__ double cos_kernel(double y) {
  double z = y * y;
  double p = 1.0 / 20922789888000.0;
  p = fma(p, z, -1.0 / 87178291200.0);
  p = fma(p, z, 1.0 / 479001600.0);
  p = fma(p, z, -1.0 / 3628800.0);
  p = fma(p, z, 1.0 / 40320.0);
  p = fma(p, z, -1.0 / 720.0);
  p = fma(p, z, 1.0 / 24.0);
  p = fma(p, z, -0.5);
  return fma(z, p, 1.0);
}
// cos(pi t), t in [0, 2]
__device__ __forceinline__ double cospi_(double t) {
  double q = rint(2.0 * t);
  double r = fma(-0.5, q, t);
  double y = r * PGN_PI;
  int qi = ((int)q) & 3;
  double c = cos_kernel(y);
  double s = sin_kernel(y);
  return qi == 0 ? c : (qi == 1 ? -s : (qi == 2 ? -c : s));
}

// Out-of-line copies for code that runs once per scan or per refreshment rather than once per density
// evaluation: an inlined fp64 division is ~50 instructions and log_/exp_ ~80, and the scan kernels are
// large enough for their instruction-cache footprint to matter (same operations, same results).
static __device__ __noinline__ double ddiv_(double a, double b) { return a / b; }
static __device__ __noinline__ double log_ni(double x) { return log_(x); }
static __device__ __noinline__ double exp_ni(double x) { return exp_(x); }

// COMPACT selects the out-of-line copies (kernels whose code size matters more than a call)
template <bool COMPACT>
__device__ __forceinline__ double div_(double a, double b) {
  if constexpr (COMPACT) return ddiv_(a, b); else return a / b;
}
template <bool COMPACT = false>
__device__ __forceinline__ double log1p_(double t) {
  double w = 1.0 + t;
  if (w == 1.0) return t;
  if constexpr (COMPACT) return log_ni(w) * ddiv_(t, w - 1.0);
  else return log_(w) * (t / (w - 1.0));
}
// LogExpFunctions.logaddexp as used by LogSum (src/recorders/LogSum.jl:10-18)
template <bool COMPACT = false>
__device__ __forceinline__ double logaddexp_(double a, double b) {
  if (a == -PGN_INF) return b;
  if (b == -PGN_INF) return a;
  double m = a > b ? a : b;
  double dlt = a > b ? b - a : a - b;
  if (a == b) dlt = 0.0;
  if constexpr (COMPACT) return m + log1p_<true>(exp_ni(dlt));
  else return m + log1p_<false>(exp_(dlt));
}

// ---- Philox4x32-10 ----------------------------------------------------------
struct Rng {
  unsigned int key0, key1;   // (seed low word, replica_index)
  unsigned int c2, c3;       // (seed high word, stream tag)
  unsigned long long ctr;    // draws consumed so far
};

__device__ __forceinline__ void philox4x32_10(unsigned int c0, unsigned int c1, unsigned int c2, unsigned int c3,
                                              unsigned int k0, unsigned int k1, unsigned int out[4]) {
  const unsigned int M0 = 0xD2511F53u, M1 = 0xCD9E8D57u;
  const unsigned int W0 = 0x9E3779B9u, W1 = 0xBB67AE85u;
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    unsigned int hi0 = __umulhi(M0, c0), lo0 = M0 * c0;
    unsigned int hi1 = __umulhi(M1, c2), lo1 = M1 * c2;
    unsigned int n0 = hi1 ^ c1 ^ k0;
    unsigned int n2 = hi0 ^ c3 ^ k1;
    c0 = n0; c1 = lo1; c2 = n2; c3 = lo0;
    k0 += W0; k1 += W1;
  }
  out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}
__device__ __forceinline__ void philox_tick(const Rng& g, unsigned long long ctr, unsigned int out[4]) {
  philox4x32_10((unsigned int)ctr, (unsigned int)(ctr >> 32), g.c2, g.c3, g.key0, g.key1, out);
}
__device__ __forceinline__ double u52(unsigned int lo, unsigned int hi) {
  unsigned long long b = ((unsigned long long)hi << 32) | lo;
  return (double)(b >> 12) * PGN_TWO_M52;
}
__device__ __forceinline__ double uniform_at(const Rng& g, unsigned long long ctr) {
  unsigned int o[4]; philox_tick(g, ctr, o);
  return u52(o[0], o[1]);
}
__device__ __forceinline__ double exponential_at(const Rng& g, unsigned long long ctr) {
  unsigned int o[4]; philox_tick(g, ctr, o);
  return -log_(1.0 - u52(o[0], o[1]));
}
__device__ __forceinline__ double normal_at(const Rng& g, unsigned long long ctr) {
  unsigned int o[4]; philox_tick(g, ctr, o);
  double u1 = 1.0 - u52(o[0], o[1]);
  double t = 2.0 * u52(o[2], o[3]);
  double rad = sqrt(-2.0 * log_(u1));
  return rad * cospi_(t);
}
static __device__ __noinline__ double normal_at_ni(unsigned int key0, unsigned int key1, unsigned int c2, unsigned int c3,
                                            unsigned long long ctr) {
  Rng g{key0, key1, c2, c3, 0ull};
  return normal_at(g, ctr);
}
__device__ __forceinline__ unsigned int bits32_at(const Rng& g, unsigned long long ctr) {
  unsigned int o[4]; philox_tick(g, ctr, o);
  return o[0];
}
// warp-uniform sequential draws (every lane computes the same value)
__device__ __forceinline__ double next_uniform(Rng& g) { return uniform_at(g, g.ctr++); }
__device__ __forceinline__ double next_exponential(Rng& g) { return exponential_at(g, g.ctr++); }

// ---- canonical summation tree: xor butterfly over the 32 lanes --------------
__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int off = 16; off >= 1; off >>= 1) v = v + __shfl_xor_sync(PGN_FULL_MASK, v, off);
  return v;
}
// Sum NV values over the warp.  Same binary tree per value as warp_sum (pairs
// (l, l^16), then (l, l^8), ...), but evaluated as a reduce-scatter: at each of
// the first log2(NV) levels a lane keeps half of its values and ships the other
// half, so the shuffle count is NV/2 + NV/4 + ... instead of 5*NV; the totals are
// then broadcast.  fp addition is commutative, so every total is bit-identical to
// the plain butterfly's.
template <int NV, int OFF>
struct ReduceScatter {
  static __device__ __forceinline__ void run(double* v, int lane) {
    if constexpr (OFF >= 1) {
      if constexpr (NV > 1) {
        constexpr int H = NV / 2;
        const bool hi = (lane & OFF) != 0;
#pragma unroll
        for (int i = 0; i < H; ++i) {
          const double keep = hi ? v[H + i] : v[i];
          const double send = hi ? v[i] : v[H + i];
          v[i] = keep + __shfl_xor_sync(PGN_FULL_MASK, send, OFF);
        }
        ReduceScatter<H, OFF / 2>::run(v, lane);
      } else {
        v[0] = v[0] + __shfl_xor_sync(PGN_FULL_MASK, v[0], OFF);
        ReduceScatter<1, OFF / 2>::run(v, lane);
      }
    }
  }
};
template <int NV> struct Log2 { static constexpr int value = 1 + Log2<NV / 2>::value; };
template <> struct Log2<1> { static constexpr int value = 0; };

// NP = NV rounded up to a power of two (2, 4, 8 or 16)
template <int NV>
__device__ __forceinline__ void warp_sum_n(double (&v)[NV]) {
  constexpr int NP = NV <= 2 ? 2 : (NV <= 4 ? 4 : (NV <= 8 ? 8 : 16));
  static_assert(NV <= 16, "warp_sum_n supports at most 16 values");
  const int lane = threadIdx.x & 31;
  double w[NP];
#pragma unroll
  for (int i = 0; i < NP; ++i) w[i] = i < NV ? v[i] : 0.0;
  ReduceScatter<NP, 16>::run(w, lane);
  constexpr int SH = 5 - Log2<NP>::value;
#pragma unroll
  for (int i = 0; i < NV; ++i) v[i] = __shfl_sync(PGN_FULL_MASK, w[0], i << SH);
}

}  // namespace pgn
