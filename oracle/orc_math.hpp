// ORACLE (test infrastructure only — never linked into the product library).
//
// Scalar arithmetic specification shared by every oracle routine: elementary
// functions built from IEEE-754 basic operations (+ - * / sqrt fma, all
// correctly rounded on x86-64 and on sm_100a), a Philox4x32-10 counter RNG,
// and the canonical 32-leaf summation tree.  The CUDA engine carries its own,
// independently typed copy of these definitions
// (pigeons.jl_b200/csrc/pgn_numerics.cuh); tests/test_gpu_parity.py::test_device_numerics_bit_identical pins the
// two bit-for-bit against each other and tests/test_oracle_math.py pins this
// file against libm/mpmath.
//
// Why not libm: the reference (Julia) uses openlibm-style exp/log and the
// Random stdlib's ziggurat; neither is available offline and neither exists on
// the device.  North star prescribes Philox; bit-exact swap permutations
// between the CPU restatement and the GPU engine then require *identical*
// transcendental functions on both sides, hence functions defined from basic
// operations only.  Compile with -ffp-contract=off: every fused operation in
// the spec is written as an explicit fma().
//
// parity unpinned (vs Julia Pigeons): the RNG stream (SplittableRandoms.jl +
// Random ziggurat) and last-ulp libm behaviour cannot be reproduced offline;
// see DESIGN.md "Oracle".
#pragma once
#include <cmath>
#include <cstdint>
#include <cstring>
#include <limits>

namespace orc {

inline double from_bits(uint64_t b) { double d; std::memcpy(&d, &b, 8); return d; }
inline uint64_t to_bits(double d) { uint64_t b; std::memcpy(&b, &d, 8); return b; }

static const double INF = std::numeric_limits<double>::infinity();
static const double QNAN = std::numeric_limits<double>::quiet_NaN();

// ---- constants (bit patterns, so that no decimal parsing is involved) ------
static const double LN2_HI = from_bits(0x3fe62e42fee00000ULL);  // 6.93147180369123816490e-01
static const double LN2_LO = from_bits(0x3dea39ef35793c76ULL);  // 1.90821492927058770002e-10
static const double INV_LN2 = from_bits(0x3ff71547652b82feULL); // 1.44269504088896338700e+00
static const double EXP_OVERFLOW = from_bits(0x40862e42fefa39efULL);   // 709.782712893384
static const double EXP_UNDERFLOW = from_bits(0xc0874910d52d3051ULL);  // -745.1332191019411
static const double PI_D = from_bits(0x400921fb54442d18ULL);
static const double LOG2PI = from_bits(0x3ffd67f1c864beb5ULL);  // log(2*pi) = 1.8378770664093453

// 2^e for e in [-1022, 1023]
inline double pow2i(int e) { return from_bits((uint64_t)(e + 1023) << 52); }

// x * 2^k with a single rounding (k in [-1080, 1030])
inline double scale2(double x, int k) {
  int k1 = k >> 1;
  int k2 = k - k1;
  return (x * pow2i(k1)) * pow2i(k2);
}

// exp: k = rint(x/ln2), r = x - k ln2 (Cody-Waite, fma), degree-13 Taylor
// polynomial in r (|r| <= 0.3466: truncation 4e-18), Horner with fma.
inline double exp_(double x) {
  if (x != x) return x;
  if (x > EXP_OVERFLOW) return INF;
  if (x < EXP_UNDERFLOW) return 0.0;
  double kf = std::nearbyint(x * INV_LN2);
  double r = std::fma(-kf, LN2_HI, x);
  r = std::fma(-kf, LN2_LO, r);
  double p = 1.0 / 6227020800.0;         // 1/13!
  p = std::fma(p, r, 1.0 / 479001600.0);  // 1/12!
  p = std::fma(p, r, 1.0 / 39916800.0);   // 1/11!
  p = std::fma(p, r, 1.0 / 3628800.0);    // 1/10!
  p = std::fma(p, r, 1.0 / 362880.0);     // 1/9!
  p = std::fma(p, r, 1.0 / 40320.0);      // 1/8!
  p = std::fma(p, r, 1.0 / 5040.0);       // 1/7!
  p = std::fma(p, r, 1.0 / 720.0);        // 1/6!
  p = std::fma(p, r, 1.0 / 120.0);        // 1/5!
  p = std::fma(p, r, 1.0 / 24.0);         // 1/4!
  p = std::fma(p, r, 1.0 / 6.0);          // 1/3!
  p = std::fma(p, r, 0.5);                // 1/2!
  p = std::fma(p, r, 1.0);
  p = std::fma(p, r, 1.0);
  return scale2(p, (int)kf);
}

// log: fdlibm/musl argument reduction x = 2^k (1+f), sqrt(2)/2 <= 1+f < sqrt(2);
// s = f/(2+f); log(1+f) = f - hfsq + s (hfsq + R(s^2)).
static const double LG1 = from_bits(0x3fe5555555555593ULL);
static const double LG2 = from_bits(0x3fd999999997fa04ULL);
static const double LG3 = from_bits(0x3fd2492494229359ULL);
static const double LG4 = from_bits(0x3fcc71c51d8e78afULL);
static const double LG5 = from_bits(0x3fc7466496cb03deULL);
static const double LG6 = from_bits(0x3fc39a09d078c69fULL);
static const double LG7 = from_bits(0x3fc2f112df3e5244ULL);

inline double log_(double x) {
  if (x != x) return x;
  if (x < 0.0) return QNAN;
  if (x == 0.0) return -INF;
  if (x == INF) return INF;
  int k = 0;
  uint64_t ix = to_bits(x);
  if (ix < 0x0010000000000000ULL) {  // subnormal: scale up by 2^54
    x = x * from_bits(0x4350000000000000ULL);
    k -= 54;
    ix = to_bits(x);
  }
  uint32_t hx = (uint32_t)(ix >> 32);
  k += (int)(hx >> 20) - 1023;
  hx &= 0x000fffffu;
  uint32_t i = (hx + 0x95f64u) & 0x100000u;
  uint64_t hi = (uint64_t)(hx | (i ^ 0x3ff00000u));
  x = from_bits((hi << 32) | (ix & 0xffffffffULL));
  k += (int)(i >> 20);
  double f = x - 1.0;
  double hfsq = 0.5 * f * f;
  double s = f / (2.0 + f);
  double z = s * s;
  double w = z * z;
  double t1 = w * std::fma(w, std::fma(w, LG6, LG4), LG2);
  double t2 = z * std::fma(w, std::fma(w, std::fma(w, LG7, LG5), LG3), LG1);
  double R = t2 + t1;
  double dk = (double)k;
  return s * (hfsq + R) + dk * LN2_LO - hfsq + f + dk * LN2_HI;
}

// cos(pi * t) for t in [0, 2]: exact quadrant reduction, Taylor kernels on
// |y| <= pi/4 (sin through y^15, cos through y^16).
inline double sin_kernel(double y) {
  double z = y * y;
  double p = -1.0 / 1307674368000.0;        // -1/15!
  p = std::fma(p, z, 1.0 / 6227020800.0);   //  1/13!
  p = std::fma(p, z, -1.0 / 39916800.0);    // -1/11!
  p = std::fma(p, z, 1.0 / 362880.0);       //  1/9!
  p = std::fma(p, z, -1.0 / 5040.0);        // -1/7!
  p = std::fma(p, z, 1.0 / 120.0);          //  1/5!
  p = std::fma(p, z, -1.0 / 6.0);           // -1/3!
  return std::fma(y * z, p, y);
}
inline double cos_kernel(double y) {
  double z = y * y;
  double p = 1.0 / 20922789888000.0;        //  1/16!
  p = std::fma(p, z, -1.0 / 87178291200.0); // -1/14!
  p = std::fma(p, z, 1.0 / 479001600.0);    //  1/12!
  p = std::fma(p, z, -1.0 / 3628800.0);     // -1/10!
  p = std::fma(p, z, 1.0 / 40320.0);        //  1/8!
  p = std::fma(p, z, -1.0 / 720.0);         // -1/6!
  p = std::fma(p, z, 1.0 / 24.0);           //  1/4!
  p = std::fma(p, z, -0.5);                 // -1/2!
  return std::fma(z, p, 1.0);
}
inline double cospi_(double t) {
  double q = std::nearbyint(2.0 * t);        // 0..4
  double r = std::fma(-0.5, q, t);           // exact, |r| <= 0.25
  double y = r * PI_D;
  int qi = ((int)q) & 3;
  double c = cos_kernel(y);
  double s = sin_kernel(y);
  // cos(pi t) = cos(pi r + q pi/2)
  return qi == 0 ? c : (qi == 1 ? -s : (qi == 2 ? -c : s));
}

// ---- Philox4x32-10 (Salmon et al. 2011) ------------------------------------
struct Philox {
  uint32_t key0, key1;   // (seed low word, replica_index)
  uint32_t c2, c3;       // (seed high word, stream tag)
};

inline void philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3,
                          uint32_t k0, uint32_t k1, uint32_t out[4]) {
  const uint32_t M0 = 0xD2511F53u, M1 = 0xCD9E8D57u;
  const uint32_t W0 = 0x9E3779B9u, W1 = 0xBB67AE85u;
  for (int r = 0; r < 10; ++r) {
    uint64_t p0 = (uint64_t)M0 * c0;
    uint64_t p1 = (uint64_t)M1 * c2;
    uint32_t hi0 = (uint32_t)(p0 >> 32), lo0 = (uint32_t)p0;
    uint32_t hi1 = (uint32_t)(p1 >> 32), lo1 = (uint32_t)p1;
    uint32_t n0 = hi1 ^ c1 ^ k0;
    uint32_t n1 = lo1;
    uint32_t n2 = hi0 ^ c3 ^ k1;
    uint32_t n3 = lo0;
    c0 = n0; c1 = n1; c2 = n2; c3 = n3;
    k0 += W0; k1 += W1;
  }
  out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}

// One RNG "tick" = one Philox block at counter value `ctr`.
inline void philox_tick(const Philox& g, uint64_t ctr, uint32_t out[4]) {
  philox4x32_10((uint32_t)ctr, (uint32_t)(ctr >> 32), g.c2, g.c3, g.key0, g.key1, out);
}

static const double TWO_M52 = from_bits(0x3cb0000000000000ULL);  // 2^-52

// 52 random mantissa bits -> [0,1) (same value set as Julia's rand(Float64))
inline double u52(uint32_t lo, uint32_t hi) {
  uint64_t b = ((uint64_t)hi << 32) | lo;
  return (double)(b >> 12) * TWO_M52;
}

inline double uniform_at(const Philox& g, uint64_t ctr) {
  uint32_t o[4]; philox_tick(g, ctr, o);
  return u52(o[0], o[1]);
}
inline double exponential_at(const Philox& g, uint64_t ctr) {
  uint32_t o[4]; philox_tick(g, ctr, o);
  return -log_(1.0 - u52(o[0], o[1]));
}
// Box-Muller (cosine branch): one normal per tick.
inline double normal_at(const Philox& g, uint64_t ctr) {
  uint32_t o[4]; philox_tick(g, ctr, o);
  double u1 = 1.0 - u52(o[0], o[1]);   // (0,1]
  double t = 2.0 * u52(o[2], o[3]);    // [0,2)
  double rad = std::sqrt(-2.0 * log_(u1));
  return rad * cospi_(t);
}
inline uint32_t bits32_at(const Philox& g, uint64_t ctr) {
  uint32_t o[4]; philox_tick(g, ctr, o);
  return o[0];
}

// ---- canonical summation: 32 lane partials + xor butterfly ------------------
// coordinate c belongs to lane (c % 32), slot (c / 32); a lane accumulates its
// slots in increasing order starting from 0.0; lanes are then combined with
// offsets 16, 8, 4, 2, 1 (v[l] = v[l] + v[l ^ off]).
inline double butterfly32(double v[32]) {
  for (int off = 16; off >= 1; off >>= 1) {
    double n[32];
    for (int l = 0; l < 32; ++l) n[l] = v[l] + v[l ^ off];
    for (int l = 0; l < 32; ++l) v[l] = n[l];
  }
  return v[0];
}
template <class F>
inline double tree_sum(int d, F term) {
  double v[32];
  for (int l = 0; l < 32; ++l) {
    double acc = 0.0;
    for (int c = l; c < d; c += 32) acc = acc + term(c);
    v[l] = acc;
  }
  return butterfly32(v);
}

// log(exp(a)+exp(b)) (role of LogExpFunctions.logaddexp in LogSum.jl:11,16)
inline double log1p_(double t) {   // t >= 0 small-ish; Kahan's trick
  double w = 1.0 + t;
  if (w == 1.0) return t;
  return log_(w) * (t / (w - 1.0));
}
inline double logaddexp_(double a, double b) {
  if (a == -INF) return b;
  if (b == -INF) return a;
  double m = a > b ? a : b;
  double dlt = a > b ? b - a : a - b;   // -|a-b|
  if (a == b) dlt = 0.0;
  return m + log1p_(exp_(dlt));
}

}  // namespace orc
