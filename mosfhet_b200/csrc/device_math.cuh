// device_math.cuh -- scalar device helpers shared by all kernels.
#pragma once
#include "common.cuh"

namespace mb {

__device__ __forceinline__ double2 cmul(double2 a, double2 b) {
  return make_double2(fma(a.x, b.x, -a.y * b.y), fma(a.x, b.y, a.y * b.x));
}
__device__ __forceinline__ double2 cmul_conj(double2 a, double2 b) {  // a * conj(b)
  return make_double2(fma(a.x, b.x, a.y * b.y), fma(a.y, b.x, -a.x * b.y));
}
__device__ __forceinline__ double2 cadd(double2 a, double2 b) { return make_double2(a.x + b.x, a.y + b.y); }
__device__ __forceinline__ double2 csub(double2 a, double2 b) { return make_double2(a.x - b.x, a.y - b.y); }
__device__ __forceinline__ void cfma(double2 &acc, double2 a, double2 b) {  // acc += a*b
  acc.x = fma(a.x, b.x, acc.x);
  acc.x = fma(-a.y, b.y, acc.x);
  acc.y = fma(a.x, b.y, acc.y);
  acc.y = fma(a.y, b.x, acc.y);
}

// round(x * 2^log / 2^64): reference torus2int (misc.c:18-22)
__device__ __forceinline__ u64 torus2int(u64 x, int log_scale) {
  return (x + (1ull << (64 - log_scale - 1))) >> (64 - log_scale);
}

// f64 -> u64 mod 2^64 with round-to-nearest (the reference's default AVX-512 build:
// scalef/reduce/scalef/cvtpd_epi64, fft_processor_spqlios.c:158-164).  Integer pipe only, so it
// does not compete with the FP64 pipe.  Valid for |x| < 2^116 (values on this path are < 2^90).
__device__ __forceinline__ u64 f64_to_torus(double x) {
  const u64 bits = (u64)__double_as_longlong(x);
  const int expo = (int)((bits >> 52) & 0x7FF);
  const u64 mant = (bits & 0x000FFFFFFFFFFFFFull) | 0x0010000000000000ull;
  const int sh = expo - 1075;
  u64 v;
  if (sh >= 0) {
    v = (sh < 64) ? (mant << sh) : 0ull;
  } else {
    const int r = -sh;
    if (r > 54 || expo == 0) v = 0ull;
    else {
      // nearest, ties to even
      const u64 q = mant >> r;
      const u64 rem = mant & ((1ull << r) - 1ull);
      const u64 half = 1ull << (r - 1);
      v = q + ((rem > half) || (rem == half && (q & 1ull)) ? 1ull : 0ull);
    }
  }
  return (bits >> 63) ? (0ull - v) : v;
}

// Gadget decomposition constants of polynomial_decompose_i (polynomial.c:74-89)
__device__ __forceinline__ u64 decomp_offset(int Bg_bit, int l) {
  u64 off = 1ull << (64 - l * Bg_bit - 1);
  for (int i = 0; i < l; i++) off += 1ull << (64 - i * Bg_bit - 1);
  return off;
}
// signed digit j of (v + offset) -- v already offset by the caller
__device__ __forceinline__ int decomp_digit(u64 v_off, int Bg_bit, int j) {
  const int sh = 64 - (j + 1) * Bg_bit;
  return (int)((v_off >> sh) & ((1ull << Bg_bit) - 1ull)) - (1 << (Bg_bit - 1));
}

// coefficient i of p * X^a (a in [0, 2N)), reading p from memory (polynomial.c:184-199)
__device__ __forceinline__ u64 rotated_coeff(const u64 *p, int i, int a, int N) {
  int src = i - a;                 // in (-2N, N)
  bool neg = false;
  if (src < 0) { src += N; neg = true; }
  if (src < 0) { src += N; neg = false; }
  const u64 v = p[src];
  return neg ? (0ull - v) : v;
}

// programmable_bootstrap input shaping (bootstrap.c:210-217)
__device__ __forceinline__ u64 pb_preprocess(u64 x, int kappa, int theta, int log_N2) {
  const u64 rnd_os = 1ull << (64 - log_N2 + theta - 1);
  const u64 theta_mask = ~((1ull << (64 - log_N2 + theta)) - 1ull);
  return ((x << kappa) + rnd_os) & theta_mask;
}

// splitmix64: counter-based generator for the synthetic keys
__device__ __host__ __forceinline__ u64 splitmix64(u64 x) {
  x += 0x9E3779B97F4A7C15ull;
  x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ull;
  x = (x ^ (x >> 27)) * 0x94D049BB133111EBull;
  return x ^ (x >> 31);
}

}  // namespace mb
