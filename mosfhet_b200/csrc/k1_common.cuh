// k1_common.cuh -- register-resident FFT building blocks shared by the k = 1 blind-rotation kernels.
#pragma once
#include <mutex>
#include "common.cuh"
#include "device_math.cuh"
#include "w64_constants.cuh"

namespace mb {

// W_64 powers in constant memory: FP64 instructions take c[bank][offset] operands directly, whereas
// folded 64-bit immediates cost two UMOV each time they are rematerialised (ncu: 184 UMOV per step).
static __constant__ double CW64C[64], CW64S[64];   // one copy per translation unit
static __constant__ double CW64T[9];               // tan(2 pi m / 64), m = 0..8 (folded-twiddle butterflies)
static void upload_w64() {
  static bool done_dev[MB_MAX_DEV] = {false};          // __constant__ memory is per device
  static std::mutex mu;                                // first calls may come from several host threads at once
  std::lock_guard<std::mutex> lk(mu);
  bool &done = done_dev[current_device()];
  if (done) return;
  MB_CHECK(cudaMemcpyToSymbol(CW64C, W64C_HOST, sizeof(double) * 64));
  MB_CHECK(cudaMemcpyToSymbol(CW64S, W64S_HOST, sizeof(double) * 64));
  MB_CHECK(cudaMemcpyToSymbol(CW64T, W64T_HOST, sizeof(double) * 9));
  MB_CHECK(cudaDeviceSynchronize());
  done = true;
}

// x * W_64^idx (or its conjugate).  idx is a compile-time constant after unrolling, so the trivial cases cost nothing
// and the 8th roots cost 2 mul + 2 add.  Every other power is folded onto the first octant, W = i^qd * (c + i s) with
// (c, s) or (s, c) taken from entries 1..7 of the tables: 14 distinct constants in all, which the compiler keeps in
// uniform registers (with one entry per power the N = 2048 kernels need 72 of the 63 uniform registers and spill them
// through local memory).  The tables hold correctly rounded values, so cos(t) read as sin(pi/2 - t) is the same double
// and the folded product rounds exactly like the direct one.
__device__ __forceinline__ double2 mul_w64(double2 x, int idx, bool conj) {
  idx &= 63;
  if (conj) idx = (64 - idx) & 63;
  const double h = 0.70710678118654752440;
  const int qd = idx >> 4, rem = idx & 15;
  double2 y;                                           // x * e^(2 pi i rem / 64)
  if (rem == 0) y = x;
  else if (rem == 8) y = make_double2((x.x - x.y) * h, (x.x + x.y) * h);
  else {
    const double c = rem < 8 ? CW64C[rem] : CW64S[16 - rem], s = rem < 8 ? CW64S[rem] : CW64C[16 - rem];
    y = make_double2(fma(x.x, c, -x.y * s), fma(x.x, s, x.y * c));
  }
  if (qd == 0) return y;                               // times i^qd
  if (qd == 1) return make_double2(-y.y, y.x);
  if (qd == 2) return make_double2(-y.x, -y.y);
  return make_double2(y.y, -y.x);
}

__host__ __device__ constexpr int brev(int x, int bits) {
  int r = 0;
  for (int i = 0; i < bits; ++i) r |= ((x >> i) & 1) << (bits - 1 - i);
  return r;
}
__host__ __device__ constexpr int clog2(int x) { return x <= 1 ? 0 : 1 + clog2(x >> 1); }

// in-register radix-R DIF: on return x[pos] = X[brev(pos)],  X_k = sum_m x_m W_R^(+mk)
template <int R>
__device__ __forceinline__ void reg_dif(double2 (&x)[R]) {
#pragma unroll
  for (int h = R / 2; h >= 1; h >>= 1) {
#pragma unroll
    for (int b = 0; b < R / 2; ++b) {
      const int j = b & (h - 1);
      const int i0 = ((b - j) << 1) + j, i1 = i0 + h;
      const double2 u = x[i0], v = x[i1];
      x[i0] = cadd(u, v);
      x[i1] = mul_w64(csub(u, v), j * (32 / h), false);
    }
  }
}

// in-register radix-R DIT inverse: input x[pos] = X[brev(pos)], output x[m] = sum_k X_k W_R^(-mk)
template <int R>
__device__ __forceinline__ void reg_dit_inv(double2 (&x)[R]) {
#pragma unroll
  for (int h = 1; h < R; h <<= 1) {
#pragma unroll
    for (int b = 0; b < R / 2; ++b) {
      const int j = b & (h - 1);
      const int i0 = ((b - j) << 1) + j, i1 = i0 + h;
      const double2 u = x[i0], v = mul_w64(x[i1], j * (32 / h), true);
      x[i0] = cadd(u, v);
      x[i1] = csub(u, v);
    }
  }
}

// ---- butterflies with the twiddle folded in (compile-time powers of W_64) ---------------------------------------------------
// (u, v) <- (u + w v, u - w v), w = W_64^idx (or its conjugate), in SIX FP64 instructions instead of 4 + 4: with
// w = i^qd * g * (1 + i tau) -- g = cos or sin of the first-octant angle, |tau| <= 1 its tangent (or the cotangent form
// w = i^qd * g * (tau + i)) -- the product is g * (p, q) with p, q one FMA each, and g rides in the FMAs that add it to u.
// Trivial powers cost the 4 additions, the eighth roots 2 additions + 4 FMAs.
__device__ __forceinline__ void bfly_w64(double2 &u, double2 &v, int idx, bool conj) {
  idx &= 63;
  if (conj) idx = (64 - idx) & 63;
  const int qd = idx >> 4, rem = idx & 15;
  double p, q, g;                                      // w v = i^qd * g * (p + i q)
  if (rem == 0) { p = v.x; q = v.y; g = 1.0; }
  else if (rem == 8) { p = v.x - v.y; q = v.x + v.y; g = 0.70710678118654752440; }
  else if (rem < 8) { const double t = CW64T[rem]; p = fma(-t, v.y, v.x); q = fma(t, v.x, v.y); g = CW64C[rem]; }          // g (1 + i t)
  else { const double t = CW64T[16 - rem]; p = fma(t, v.x, -v.y); q = fma(t, v.y, v.x); g = CW64C[16 - rem]; }            // g (t + i)
  // rotate by i^qd: (p, q) -> (-q, p) -> (-p, -q) -> (q, -p)
  const double a = (qd == 0) ? p : (qd == 1) ? -q : (qd == 2) ? -p : q;
  const double b = (qd == 0) ? q : (qd == 1) ? p : (qd == 2) ? -q : -p;
  const double2 u0 = u;
  if (rem == 0) { u = make_double2(u0.x + a, u0.y + b); v = make_double2(u0.x - a, u0.y - b); }
  else { u = make_double2(fma(g, a, u0.x), fma(g, b, u0.y)); v = make_double2(fma(-g, a, u0.x), fma(-g, b, u0.y)); }
}

// In-register radix-R DFT of TWISTED input, X_k = sum_m x_m (t W_R^k)^m with t = W_64^TW (TW = 0: the plain DFT of
// reg_dif), as a decimation-in-time network of folded butterflies: the twist t^m is absorbed into the butterfly twiddles
// (stage of block size 2h: t^(R/2h) W_2h^j), so pass A pays nothing for it.  Natural-order input; on return
// x[pos] = X[brev(pos)] like reg_dif.
template <int R, int TW>
__device__ __forceinline__ void reg_dft_fma(double2 (&x)[R]) {
  constexpr int LOGR = clog2(R);
  double2 a[R];
#pragma unroll
  for (int i = 0; i < R; ++i) a[i] = x[brev(i, LOGR)];
#pragma unroll
  for (int h = 1; h < R; h <<= 1) {
#pragma unroll
    for (int b = 0; b < R / 2; ++b) {
      const int j = b & (h - 1);
      const int i0 = ((b - j) << 1) + j, i1 = i0 + h;
      bfly_w64(a[i0], a[i1], TW * (R / (2 * h)) + j * (32 / h), false);
    }
  }
#pragma unroll
  for (int i = 0; i < R; ++i) x[i] = a[brev(i, LOGR)];
}

// reg_dit_inv with folded butterflies: input x[pos] = X[brev(pos)], output x[m] = sum_k X_k W_R^(-mk)
template <int R>
__device__ __forceinline__ void reg_dit_inv_fma(double2 (&x)[R]) {
#pragma unroll
  for (int h = 1; h < R; h <<= 1) {
#pragma unroll
    for (int b = 0; b < R / 2; ++b) {
      const int j = b & (h - 1);
      const int i0 = ((b - j) << 1) + j, i1 = i0 + h;
      bfly_w64(x[i0], x[i1], j * (32 / h), true);
    }
  }
}

__device__ __forceinline__ int swz(int s) { return s ^ ((s >> 3) & 7); }

struct K1Args {
  const double2 *bsk;
  const double2 *tab;     // TA[16][S] then TB[R2][8]
  const u64 *tv;
  int tv_count;
  // DIRECT instantiations (one external product / CMUX) reuse three bootstrap-only slots through anonymous unions.
  // The struct must keep its size and layout: the Level-2 kernel sits on the 255-register cliff and ptxas' allocation
  // depends on the parameter block (32 more bytes of parameters: 176 B of spills and 141 -> 200 ms per 4096).
  union { const u64 *in; const int *sel; };          // sel: [count] TRGSW index per ciphertext (when sel_const < 0)
  union { int in_stride; int sel_const; };
  int in_div;          // ciphertexts per input TLWE (>= 1): ct uses input ct / in_div, test vector ct % tv_count
  int size;
  u64 *out;
  int extract, init_rotate;
  union { u64 prec_offset; const u64 *in1; };        // in1: CMUX operand, product of (tv - in1) and result in1 + product
  int preprocess, kappa, theta;
  int Bg_bit;
  int count;              // ciphertexts in the launch (kernels with several ciphertexts per CTA)
};
static_assert(sizeof(K1Args) == 104, "K1Args layout is performance critical, see above");

// f64 -> u64 mod 2^64, round to nearest (AVX-512 path of the reference, fft_processor_spqlios.c:158-164)
// done on the FP64 pipe + one F2I instead of ~25 integer instructions: r = rint(x / 2^64) by the
// 1.5*2^52 trick (|x| < 2^115), y = x - r*2^64 is exact and |y| <= 2^63, then a rounding convert.
__device__ __forceinline__ u64 f64_to_torus_fast(double x) {
  const double C = 6755399441055744.0;
  const double t = x * 5.42101086242752217e-20;       // 2^-64
  const double r = (t + C) - C;
  const double y = fma(r, -18446744073709551616.0, x);
  return (u64)__double2ll_rn(y);
}

// ---- tensor memory as a per-thread constant store -----------------------------------------------
// The 16 pass-A / A' twiddles of a thread are thread constants that do not fit the register budget, so they were
// re-read from global memory (through the L1 data pipe, the busiest unit of this kernel) three times per step.
// Blackwell's tensor memory is private per lane and has its own datapath: each thread parks its 64 words there once
// (tcgen05.st) and fetches them 16 words at a time (tcgen05.ld.32x32b.x16) when needed.
__device__ __forceinline__ void tmem_st4(unsigned taddr, const double2 (&v)[4]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};"
      ::"r"(taddr), "r"(__double2loint(v[0].x)), "r"(__double2hiint(v[0].x)), "r"(__double2loint(v[0].y)),
      "r"(__double2hiint(v[0].y)), "r"(__double2loint(v[1].x)), "r"(__double2hiint(v[1].x)), "r"(__double2loint(v[1].y)),
      "r"(__double2hiint(v[1].y)), "r"(__double2loint(v[2].x)), "r"(__double2hiint(v[2].x)), "r"(__double2loint(v[2].y)),
      "r"(__double2hiint(v[2].y)), "r"(__double2loint(v[3].x)), "r"(__double2hiint(v[3].x)), "r"(__double2loint(v[3].y)),
      "r"(__double2hiint(v[3].y))
      : "memory");
}
// 16 raw words: 4 x (acc[j], acc[j + M]) as u64 pairs
__device__ __forceinline__ void tmem_st_u64x8(unsigned taddr, const u64 (&v)[8]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};"
      ::"r"(taddr), "r"((unsigned)v[0]), "r"((unsigned)(v[0] >> 32)), "r"((unsigned)v[1]), "r"((unsigned)(v[1] >> 32)),
      "r"((unsigned)v[2]), "r"((unsigned)(v[2] >> 32)), "r"((unsigned)v[3]), "r"((unsigned)(v[3] >> 32)),
      "r"((unsigned)v[4]), "r"((unsigned)(v[4] >> 32)), "r"((unsigned)v[5]), "r"((unsigned)(v[5] >> 32)),
      "r"((unsigned)v[6]), "r"((unsigned)(v[6] >> 32)), "r"((unsigned)v[7]), "r"((unsigned)(v[7] >> 32))
      : "memory");
}
__device__ __forceinline__ void tmem_ld_u64x8(u64 (&v)[8], unsigned taddr) {
  unsigned w[16];
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];\n\t"
      "tcgen05.wait::ld.sync.aligned;"
      : "=r"(w[0]), "=r"(w[1]), "=r"(w[2]), "=r"(w[3]), "=r"(w[4]), "=r"(w[5]), "=r"(w[6]), "=r"(w[7]), "=r"(w[8]),
        "=r"(w[9]), "=r"(w[10]), "=r"(w[11]), "=r"(w[12]), "=r"(w[13]), "=r"(w[14]), "=r"(w[15])
      : "r"(taddr)
      : "memory");
#pragma unroll
  for (int i = 0; i < 8; ++i) v[i] = ((u64)w[2 * i + 1] << 32) | (u64)w[2 * i];
}
// tcgen05.ld.16x256b.x1 (measured mapping, profiles/r1r_tmem_shapes.log): thread t receives the 64-bit word at columns
// 2(t%4), 2(t%4)+1 of lane L0 + t/4 and of lane L0 + 8 + t/4 (L0 = lane field of taddr, 0 or 16 inside the warp's window)
__device__ __forceinline__ void tmem_ld_16x256(double &lo_lane, double &hi_lane, unsigned taddr) {
  int w0, w1, w2, w3;
  asm volatile("tcgen05.ld.sync.aligned.16x256b.x1.b32 {%0, %1, %2, %3}, [%4];"
               : "=r"(w0), "=r"(w1), "=r"(w2), "=r"(w3) : "r"(taddr) : "memory");
  lo_lane = __hiloint2double(w1, w0);
  hi_lane = __hiloint2double(w3, w2);
}
// .x4: four consecutive 256-bit windows (32 columns); registers 4i..4i+3 belong to window i with the same pattern
__device__ __forceinline__ void tmem_ld_16x256_x4(double (&lo)[4], double (&hi)[4], unsigned taddr) {
  int w[16];
  asm volatile(
      "tcgen05.ld.sync.aligned.16x256b.x4.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(w[0]), "=r"(w[1]), "=r"(w[2]), "=r"(w[3]), "=r"(w[4]), "=r"(w[5]), "=r"(w[6]), "=r"(w[7]), "=r"(w[8]),
        "=r"(w[9]), "=r"(w[10]), "=r"(w[11]), "=r"(w[12]), "=r"(w[13]), "=r"(w[14]), "=r"(w[15])
      : "r"(taddr)
      : "memory");
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    lo[i] = __hiloint2double(w[4 * i + 1], w[4 * i]);
    hi[i] = __hiloint2double(w[4 * i + 3], w[4 * i + 2]);
  }
}
__device__ __forceinline__ void tmem_wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void tmem_st_f64x8(unsigned taddr, double a0, double a1, double a2, double a3, double a4,
                                              double a5, double a6, double a7) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};"
      ::"r"(taddr), "r"(__double2loint(a0)), "r"(__double2hiint(a0)), "r"(__double2loint(a1)), "r"(__double2hiint(a1)),
      "r"(__double2loint(a2)), "r"(__double2hiint(a2)), "r"(__double2loint(a3)), "r"(__double2hiint(a3)),
      "r"(__double2loint(a4)), "r"(__double2hiint(a4)), "r"(__double2loint(a5)), "r"(__double2hiint(a5)),
      "r"(__double2loint(a6)), "r"(__double2hiint(a6)), "r"(__double2loint(a7)), "r"(__double2hiint(a7))
      : "memory");
}
__device__ __forceinline__ void tmem_st_u32x16(unsigned taddr, const unsigned (&w)[16]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};"
      ::"r"(taddr), "r"(w[0]), "r"(w[1]), "r"(w[2]), "r"(w[3]), "r"(w[4]), "r"(w[5]), "r"(w[6]), "r"(w[7]), "r"(w[8]),
      "r"(w[9]), "r"(w[10]), "r"(w[11]), "r"(w[12]), "r"(w[13]), "r"(w[14]), "r"(w[15])
      : "memory");
}
__device__ __forceinline__ void tmem_ld_u32x16(unsigned (&w)[16], unsigned taddr) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];\n\t"
      "tcgen05.wait::ld.sync.aligned;"
      : "=r"(w[0]), "=r"(w[1]), "=r"(w[2]), "=r"(w[3]), "=r"(w[4]), "=r"(w[5]), "=r"(w[6]), "=r"(w[7]), "=r"(w[8]),
        "=r"(w[9]), "=r"(w[10]), "=r"(w[11]), "=r"(w[12]), "=r"(w[13]), "=r"(w[14]), "=r"(w[15])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_ld4(double2 (&v)[4], unsigned taddr) {
  int w[16];
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];\n\t"
      "tcgen05.wait::ld.sync.aligned;"
      : "=r"(w[0]), "=r"(w[1]), "=r"(w[2]), "=r"(w[3]), "=r"(w[4]), "=r"(w[5]), "=r"(w[6]), "=r"(w[7]), "=r"(w[8]),
        "=r"(w[9]), "=r"(w[10]), "=r"(w[11]), "=r"(w[12]), "=r"(w[13]), "=r"(w[14]), "=r"(w[15])
      : "r"(taddr)
      : "memory");
#pragma unroll
  for (int i = 0; i < 4; ++i)
    v[i] = make_double2(__hiloint2double(w[4 * i + 1], w[4 * i]), __hiloint2double(w[4 * i + 3], w[4 * i + 2]));
}

// Key rows are read with the plain read-only load.  The streaming form (ld.global.nc.L1::no_allocate), used while the
// twiddles still competed for the 28 KB of L1, also leaves the lines first in line for eviction from L2: a single
// bootstrap then re-reads the whole key from HBM every time once something else has displaced it (2.6 -> 3.4 ms,
// scripts/latency_after_load.py), and the full batch is 2.3 % slower (profiles/r1p_k1_tmem.log).  Round 2, T = M/4 kernel:
// no_allocate 0.8 % slower, no_allocate + an L2 evict_last policy the same as the plain load (profiles/r2r_key_load_policy.log).
__device__ __forceinline__ double2 ldg_key(const double2 *p) { return __ldg(p); }


}  // namespace mb
