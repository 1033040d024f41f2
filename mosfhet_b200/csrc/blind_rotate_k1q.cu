// blind_rotate_k1q.cu -- the k = 1 blind rotation at FOUR warps per scheduler: T = M/4 threads per ciphertext
// (128 at N = 1024, 256 at N = 2048) with a quarter of the per-thread state of blind_rotate_k1.cu, so that 16 warps
// are resident per SM at <= 128 registers instead of 8 warps at 254 (ncu on that kernel, profiles/r1u: 1.94 warps per
// scheduler, issue slots 45 % busy, stall mix dominated by fixed-latency and shared-memory waits).
//
// Every N/2-point negacyclic transform is three register-resident passes, radix RA x 16 x 4 (RA = M/64), in the same
// position order as the other k = 1 kernels (position s holds the value at root exponent 1 + 4*bitrev(s)), so the
// resident key layout of keys.cu is shared:
//
//   pass A  thread (row slot, q), q < 64: coefficients q + 64 m (+M), m < RA, of one polynomial and one gadget level:
//           (X^a - 1)*acc, digit -> f64, fold/twist, radix-RA DIF, twiddle w^q W_M^(q k1), -> smem block pos1
//   pass B  task (row, pos1, r), r < 4: the 16 elements r + 4 m2 of block pos1: radix-16 DIF in place, NO twiddle
//   pass C  thread c owns positions 4c..4c+3: twiddle W_64^(r k2) on load (3 thread-constant values), radix-4 DIF,
//           Fourier-domain multiply-accumulate against both output polynomials' key rows from registers
//   inverse C' -> B' -> A' mirrors it (DIT, conjugate twiddles); A' untwists, reduces mod 2^64 and accumulates.
//
// Shared-memory elements are 16-byte complex values at phys(i) = i ^ ((i >> 3) & 7) ^ (bit6(i) << 2): all three
// access patterns are bank-conflict free (scripts/experiments/k1q_index_model.py checks this and the index algebra on
// the CPU).  Thread constants that do not fit 128 registers live in tensor memory as in blind_rotate_k1.cu: the RA
// pass-A twiddles, the 3 pass-C twiddles and (N = 1024) the accumulator words the thread owns.
//
// Reference functions fused: bootstrap.c:107-122, 192-206, polynomial.c:74-89, 220-235, 359-375 (+ src/fft),
// trlwe.c:437, 491-505, 540-552, 629-634 -- see blind_rotate_k1.cu.
#include <map>
#include <mutex>
#include <type_traits>
#include <vector>

#include "k1_common.cuh"

namespace mb {

// y = x * 2^-64 / M reduced mod 1, as a 64-bit torus word; round to nearest like fft_processor_spqlios.c:158-164.
// s = t + 1.5*2^52 rounds t to an integer r = s - 1.5*2^52; (t - r) * 2^64 is formed exactly by two FMAs.
__device__ __forceinline__ u64 f64_to_torus_scaled(double x, double scale /* 2^-64 / M */) {
  const double C = 6755399441055744.0, P64 = 18446744073709551616.0;
  const double t = x * scale;
  const double s = t + C;
  const double q = fma(s, -P64, C * P64);          // -(r * 2^64), exact (C * 2^64 is a power-of-two multiple of C)
  const double y = fma(t, P64, q);                 // (t - r) * 2^64, exact
  return (u64)__double2ll_rn(y);
}

// (cos, sin) of W_64^idx for a compile-time idx, from the 14 first-octant constants (see mul_w64)
__device__ __forceinline__ void w64_cs(int idx, double &c, double &s) {
  idx &= 63;
  const double h = 0.70710678118654752440;
  const int qd = idx >> 4, rem = idx & 15;
  double c0, s0;
  if (rem == 0) { c0 = 1.0; s0 = 0.0; }
  else if (rem == 8) { c0 = h; s0 = h; }
  else { c0 = rem < 8 ? CW64C[rem] : CW64S[16 - rem]; s0 = rem < 8 ? CW64S[rem] : CW64C[16 - rem]; }
  if (qd == 0) { c = c0; s = s0; } else if (qd == 1) { c = -s0; s = c0; } else if (qd == 2) { c = -c0; s = -s0; } else { c = s0; s = -c0; }
}

// Tensor-memory loads issued early and waited for late: the 16 words become valid at tmem_wait, which takes them as
// read-write operands so that no use can be scheduled ahead of it.
struct TmemLd { unsigned w[16]; };
__device__ __forceinline__ void tmem_issue(TmemLd &t, unsigned taddr) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(t.w[0]), "=r"(t.w[1]), "=r"(t.w[2]), "=r"(t.w[3]), "=r"(t.w[4]), "=r"(t.w[5]), "=r"(t.w[6]), "=r"(t.w[7]), "=r"(t.w[8]),
        "=r"(t.w[9]), "=r"(t.w[10]), "=r"(t.w[11]), "=r"(t.w[12]), "=r"(t.w[13]), "=r"(t.w[14]), "=r"(t.w[15])
      : "r"(taddr)
      : "memory");
}
#define MB_TMEM_RW(t) "+r"((t).w[0]), "+r"((t).w[1]), "+r"((t).w[2]), "+r"((t).w[3]), "+r"((t).w[4]), "+r"((t).w[5]), "+r"((t).w[6]), \
                      "+r"((t).w[7]), "+r"((t).w[8]), "+r"((t).w[9]), "+r"((t).w[10]), "+r"((t).w[11]), "+r"((t).w[12]), "+r"((t).w[13]), \
                      "+r"((t).w[14]), "+r"((t).w[15])
__device__ __forceinline__ void tmem_wait(TmemLd &a) { asm volatile("tcgen05.wait::ld.sync.aligned;" : MB_TMEM_RW(a)::"memory"); }
__device__ __forceinline__ void tmem_also(TmemLd &a) { asm volatile("" : MB_TMEM_RW(a)); }   // after a tmem_wait: same dependency, no instruction
__device__ __forceinline__ double2 tmem_c(const TmemLd &t, int i) {
  return make_double2(__hiloint2double((int)t.w[4 * i + 1], (int)t.w[4 * i]), __hiloint2double((int)t.w[4 * i + 3], (int)t.w[4 * i + 2]));
}
__device__ __forceinline__ u64 tmem_u(const TmemLd &t, int i) { return ((u64)t.w[2 * i + 1] << 32) | (u64)t.w[2 * i]; }

// Kernel-uniform constants travel in the parameter block: FP64 and integer instructions take c[0][..] operands directly,
// which keeps them out of the 128-register budget (at 128 registers every long-lived scalar counts).
struct K1QArgs {
  K1Args a;
  u64 off;            // decomp_offset(Bg_bit, l)                                  (polynomial.c:80-83)
  double dbias;       // 2^52 + Bg/2: digit u in [0, Bg) -> double(u - Bg/2) = hiloint2double(0x43300000, u) - dbias, exact
  double out_scale;   // 2^-64 / M
  unsigned dmask;     // Bg - 1
  int b_index;        // index of b in the input row (init_rotate)
};

template <int LOGM, int L, int LB, bool PKALL, int VAR>
__global__ void __launch_bounds__((1 << LOGM) / 4, 512 / ((1 << LOGM) / 4)) blind_rotate_k1q_kernel(const __grid_constant__ K1QArgs Q) {
  const K1Args &A = Q.a;
  // VAR bit 0 (KPF): pass-C key values pipelined one half row ahead; bit 1 (HYB): two-row phases of pass B mapped to two of
  // the warps (full radix-16 tasks, block barriers) instead of lane pairs in every warp -- see the two-row branches below
  // bit 2 (FOLD): butterflies with the twiddle folded in (k1_common.cuh), the pass-A twist absorbed into them
  constexpr bool KPF = (VAR & 1) != 0, HYB = (VAR & 2) != 0, FOLD = (VAR & 4) != 0;
  constexpr int M = 1 << LOGM, N = 2 * M, T = M / 4, RA = M / 64, LOGRA = LOGM - 6;
  constexpr int SLOTS = T / 64;                       // pass-A rows in flight: 2 (N = 1024), 4 (N = 2048)
  constexpr int LBO = SLOTS / 2;                      // gadget levels per pass-A round
  constexpr int ROWS = 2 * L;
  constexpr int WQ = 2048 / N;                        // w^(64 m) = W_64^(WQ * m)
  static_assert(LOGM == 9 || LOGM == 10, "k1q: N = 1024 or 2048");
  static_assert(LB >= 1 && LB <= L && LB <= 2, "levels per batch");

  extern __shared__ __align__(16) unsigned char smem_raw[];
  u64 *acc = reinterpret_cast<u64 *>(smem_raw);                       // [2][N]
  double2 *buf = reinterpret_cast<double2 *>(acc + 2 * N);            // [2*LB][M]
  unsigned short *rot = reinterpret_cast<unsigned short *>(buf + 2 * LB * M);
  __shared__ unsigned tmem_base_s;

  const int tid = threadIdx.x, ct = blockIdx.x;
  const int log_N2 = LOGM + 2;
  const u64 *in = A.in + (size_t)(ct / A.in_div) * A.in_stride;
  const u64 *tv = A.tv + (size_t)(A.tv_count > 1 ? ct % A.tv_count : 0) * 2 * N;
  const int Bg_bit = A.Bg_bit;

  // ---- thread roles.  They are re-derived from the thread index inside the step loop (MB_K1Q_ROLES): as loop invariants
  // the compiler keeps the dozen derived offsets live across the whole step in registers it does not have, spills them,
  // and the reloads miss the 28 KB of L1 that is left (3 % of the kernel waiting on local loads, profiles/r2b)
#define MB_K1Q_ROLES(t_)                                                                                                       \
  const int slot = (t_) >> 6, qA = (t_) & 63;         /* pass A / A': row slot and in-block index */                           \
  const int pA = slot & 1, lboA = slot >> 1;          /* polynomial and level offset inside a round */                         \
  /* pass B / B' run PER WARP on the two blocks pos1 = 2*warp + {0, 1} whose positions the same warp's pass C / C' */        \
  /* threads own, so that B -> C and C' -> B' need a warp barrier only: lane = (row of the batch, block bit, r) */            \
  const int warp = (t_) >> 5, lane = (t_) & 31;                                                                                \
  const int rowB = lane >> 3, pos1B = 2 * warp + ((lane >> 2) & 1), rB = lane & 3;                                             \
  /* two-row phases: the 16 tasks of a warp are split over lane pairs (lane, lane + 16), see dif16_pair */                     \
  const int hB = lane >> 4, rowB2 = (lane >> 3) & 1;                                                                           \
  /* pass C / C': lanes 0..15 take the key columns c' = 16*warp + lane of the even position groups, lanes 16..31 of the */   \
  /* odd ones: the 8 lanes of a quarter warp read 8 consecutive 16-byte key words, one 128-byte line per request (with */    \
  /* c = tid a request straddles two lines and the key loads cost twice the L1 wavefronts: profiles/r2a) */                   \
  const int cpC = 16 * warp + (lane & 15), halfC = lane >> 4;                                                                  \
  const int cC = 2 * cpC + halfC, pos2C = cC & 15;                                                                             \
  /* swizzled offsets (elements): see the header */                                                                            \
  const int qs0 = qA ^ ((qA >> 3) & 7), qs1 = qs0 ^ 4;                /* pass A: block pos1 even / odd */                     \
  const int b4B = (pos1B & 1) << 2;                                    /* pass B: bit 2 flips in odd blocks */                 \
  const int cbase = 8 * cpC, cxor = (4 * halfC) ^ (cpC & 7) ^ (((cpC >> 3) & 1) << 2);   /* pass C: element r at cbase + (r ^ cxor) */ \
  /* element 4*e + rB of a block (e < 16), swizzled: bits 0-1 ^= (e >> 1) & 3, bit 2 ^= bit 2 of (e >> 1) and the block bit */ \
  auto swz_b = [&](int e) { return ((4 * e) ^ ((e >> 1) & 4) ^ b4B) + (rB ^ ((e >> 1) & 3)); };                                \
  /* the same for e = 8*hB + j, j < 8 (two-row phases) */                                                                      \
  auto swz_b2 = [&](int j) { return 32 * hB + ((4 * j) ^ (4 * hB) ^ b4B) + (rB ^ ((j >> 1) & 3)); };                          \
  (void)lboA; (void)rowB; (void)rowB2; (void)pos2C; (void)qs1; (void)cbase; (void)cxor; (void)swz_b; (void)swz_b2; (void)hB;

  // ---- tensor memory: per-thread constants and state ------------------------------------------------------------
  constexpr bool TMEM_ACC = (LOGM == 9);              // N = 2048: two warps share a lane quarter and pass-A threads are not the owners
  constexpr int CPT = 128;                            // columns per thread
  constexpr int TMEM_COLS = CPT * (T / 128);
  // N = 1024: TA 0..31, TC 32..47, FA 48..63 + 112..127, ACC 64..95, PK 96..111; N = 2048: TA 0..63, TC 64..79, FA 80..111
  constexpr int COL_TA = 0, COL_TC = 4 * RA, COL_ACC = 64, COL_PK = 96;
  // the packed digit words outlive the first batch: keep them out of passes B, C (N = 1024; at N = 2048 the columns are taken)
  constexpr bool PK_PARK = PKALL && (L > LB) && LOGM == 9;
  // The 2 x 4 Fourier accumulators (32 registers) are only touched in pass C: between batches they wait in tensor memory,
  // which is what lets passes A and B (16 complex values in flight) fit 128 registers without spilling
  constexpr bool FA_PARK = (L > LB);
  constexpr int COL_FA0 = (LOGM == 9) ? 48 : 80, COL_FA1 = (LOGM == 9) ? 112 : 96;
  static_assert(COL_FA1 + 16 <= CPT && (LOGM == 9 || COL_TC + 16 <= COL_FA0), "tensor-memory column layout");
  static_assert(!TMEM_ACC || (COL_TC + 16 <= 64 && COL_ACC == 64), "tensor-memory column layout");
  static_assert(COL_PK + 2 * RA <= CPT && COL_TC + 16 <= COL_PK, "tensor-memory column layout");
  if (tid < 32) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;\n\t"
                 "tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;"
                 ::"r"((unsigned)__cvta_generic_to_shared(&tmem_base_s)), "n"(TMEM_COLS) : "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  // lane field (bits 31:16) = 32 * (warp % 4); warps 4..7 (N = 2048) use the second half of the columns
  const unsigned taddr = tmem_base_s + ((unsigned)((tid >> 5) & 3) << 21) + (unsigned)((tid >> 7) * CPT);
  {
    MB_K1Q_ROLES(tid)
    // pass-A twiddles w^q * W_M^(q k1), k1 = brev(pos1): A.tab[k1 * 64 + q]
#pragma unroll
    for (int g4 = 0; g4 < RA / 4; ++g4) {
      double2 tw[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) tw[i] = __ldg(&A.tab[brev(4 * g4 + i, LOGRA) * 64 + qA]);
      tmem_st4(taddr + COL_TA + 16 * g4, tw);
    }
    // pass-C twiddles W_64^(r k2), r = 1..3, k2 = brev(pos2): A.tab[RA * 64 + k2 * 4 + r]
    double2 tw[4];
#pragma unroll
    for (int r = 0; r < 4; ++r) tw[r] = __ldg(&A.tab[RA * 64 + brev(pos2C, 4) * 4 + r]);
    tmem_st4(taddr + COL_TC, tw);
    asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
  }

  // ---- initial accumulator: tv * X^(2N - round((b + 1/(4*torus_base)) * 2N))  (bootstrap.c:194-195) ----------------
  int rot0 = 0;
  if (A.init_rotate) {
    u64 b = in[Q.b_index];
    if (A.preprocess) b = pb_preprocess(b, A.kappa, A.theta, log_N2);
    rot0 = (2 * N - (int)torus2int(b + A.prec_offset, log_N2)) & (2 * N - 1);
  }
  for (int c = tid; c < 2 * N; c += T) {
    const int p = c / N, i = c - p * N;
    acc[c] = rot0 ? rotated_coeff(tv + (size_t)p * N, i, rot0, N) : tv[c];
  }
  for (int i = tid; i < A.size; i += T) {             // all rotation amounts up front (bootstrap.c:113)
    u64 av = in[i];
    if (A.preprocess) av = pb_preprocess(av, A.kappa, A.theta, log_N2);
    rot[i] = (unsigned short)(torus2int(av, log_N2) & (2 * N - 1));
  }
  __syncthreads();
  if (TMEM_ACC) {                                     // park the accumulator words this thread owns: (j, j + M), j = q + 64 m
    MB_K1Q_ROLES(tid)
    const u64 *ap0 = acc + pA * N;
#pragma unroll
    for (int g4 = 0; g4 < RA / 4; ++g4) {
      u64 v[8];
#pragma unroll
      for (int i = 0; i < 4; ++i) { v[2 * i] = ap0[qA + (4 * g4 + i) * 64]; v[2 * i + 1] = ap0[qA + (4 * g4 + i) * 64 + M]; }
      tmem_st_u64x8(taddr + COL_ACC + 16 * g4, v);
    }
    asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
  }


  for (int step = 0; step < A.size; ++step) {
    const int a_i = rot[step];
    if (a_i == 0) continue;                           // bootstrap.c:114
    const double2 *__restrict__ key = A.bsk + (size_t)step * ROWS * 2 * M;
    int t_ = tid;
    asm volatile("" : "+r"(t_));                      // opaque copy of the thread index: the roles below are not loop invariants
    MB_K1Q_ROLES(t_)

    double2 fa[2][4];                                 // Fourier accumulators: positions 4c..4c+3 of both outputs
#pragma unroll
    for (int o = 0; o < 2; ++o)
#pragma unroll
      for (int i = 0; i < 4; ++i) fa[o][i] = make_double2(0.0, 0.0);

    // coefficient pair m of (X^a - 1)*acc + rounding offset (polynomial.c:74-89, 220-235): j = q + 64 m and j + M
    const u64 *ap = acc + pA * N;
    int base = (qA - a_i) & (2 * N - 1);              // index of coefficient qA in acc * X^a (sign in bit log2 N)
    u64 own[8];
    auto coef_pair = [&](int m, u64 &v0, u64 &v1) {
      const int j = qA + m * 64;
      const int s0 = (base + m * 64) & (2 * N - 1), s1 = (s0 + M) & (2 * N - 1);
      const u64 r0 = ap[s0 & (N - 1)], r1 = ap[s1 & (N - 1)];
      if (TMEM_ACC && (m & 3) == 0) tmem_ld_u64x8(own, taddr + COL_ACC + 4 * m);
      const u64 t0 = Q.off - (TMEM_ACC ? own[2 * (m & 3)] : ap[j]), t1 = Q.off - (TMEM_ACC ? own[2 * (m & 3) + 1] : ap[j + M]);
      v0 = (s0 & N) ? t0 - r0 : t0 + r0;
      v1 = (s1 & N) ? t1 - r1 : t1 + r1;
    };
    // FUSED: a thread transforms ONE level per batch (N = 2048: the four row slots are 2 polynomials x 2 levels), so the
    // digit is taken straight from the coefficient; otherwise the top lev_end*Bg_bit bits of every coefficient are packed
    // into one 32-bit word and serve the levels of the batch (PKALL: of the whole step)
    constexpr bool FUSED = !PKALL && (LBO >= LB);
    // FUSED with exactly two batches (N = 2048, l = 4): the first batch also takes the digit of the thread's level in the
    // SECOND batch from the same coefficient and parks the 16 digit pairs (one word per coefficient pair) in tensor memory;
    // the second batch then needs no accumulator reads and no 64-bit arithmetic at all (they were a third of the kernel,
    // profiles/r2c_level2: pass A at 34 % FP64 instructions)
    constexpr bool DIG_PARK = FUSED && LOGM == 10 && L == 2 * LB && LBO == LB;
    constexpr int COL_DG = 112;
    unsigned pk0[FUSED ? 1 : RA], pk1[FUSED ? 1 : RA];
    auto pack_digits = [&](int lev_end) {
      const int pk_shift = 64 - lev_end * Bg_bit;
#pragma unroll
      for (int m = 0; m < RA; ++m) {
        u64 v0, v1;
        coef_pair(m, v0, v1);
        pk0[FUSED ? 0 : m] = (unsigned)(v0 >> pk_shift);
        pk1[FUSED ? 0 : m] = (unsigned)(v1 >> pk_shift);
      }
    };
    if (PKALL) {
      pack_digits(L);
      if (PK_PARK) {
#pragma unroll
        for (int h = 0; h < RA / 8; ++h) {
          unsigned w[16];
#pragma unroll
          for (int i = 0; i < 8; ++i) { w[2 * i] = pk0[8 * h + i]; w[2 * i + 1] = pk1[8 * h + i]; }
          tmem_st_u32x16(taddr + COL_PK + 16 * h, w);
        }
        asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
      }
    }

    auto batch = [&](auto nb_tag, const int lev0) {
      constexpr int NB = decltype(nb_tag)::value, RB = 2 * NB;
      // ------------------------------- pass A -----------------------------------------------------------------
      // FUSED: every batch re-reads the coefficients; hide the common subexpressions (32 shared-memory addresses) from the
      // compiler, which otherwise keeps them across the batch in registers it does not have and spills them
      if (FUSED && !DIG_PARK) asm volatile("" : "+r"(base));
      if (!PKALL && !FUSED) pack_digits(lev0 + NB);
      if (PK_PARK && lev0 > 0) {
#pragma unroll
        for (int h = 0; h < RA / 8; ++h) {
          unsigned w[16];
          tmem_ld_u32x16(w, taddr + COL_PK + 16 * h);
#pragma unroll
          for (int i = 0; i < 8; ++i) { pk0[8 * h + i] = w[2 * i]; pk1[8 * h + i] = w[2 * i + 1]; }
        }
      }
      constexpr int ROUNDS = (NB + LBO - 1) / LBO;
#pragma unroll
      for (int rd = 0; rd < ROUNDS; ++rd) {
        const int lb = rd * LBO + lboA;
        if (LBO > 1 && lb >= NB) continue;           // warp-uniform (a slot is two whole warps)
        const int sh = FUSED ? 64 - (lev0 + lb + 1) * Bg_bit : ((PKALL ? (L - 1 - lev0) : (NB - 1)) - lb) * Bg_bit;
        TmemLd ta0, ta1;                              // the first 8 twiddles fly under the digit conversion and the butterflies
        tmem_issue(ta0, taddr + COL_TA);
        tmem_issue(ta1, taddr + COL_TA + 16);
        double2 x[RA];
        unsigned dg[16];
        if constexpr (DIG_PARK) { if (lev0 > 0) tmem_ld_u32x16(dg, taddr + COL_DG); }
#pragma unroll
        for (int m = 0; m < RA; ++m) {
          unsigned u0, u1;
          if (DIG_PARK && lev0 > 0) {
            u0 = dg[m & 15] & Q.dmask; u1 = (dg[m & 15] >> 16) & Q.dmask;
          } else if (FUSED) {
            u64 v0, v1;
            coef_pair(m, v0, v1);
            u0 = (unsigned)(v0 >> sh) & Q.dmask; u1 = (unsigned)(v1 >> sh) & Q.dmask;
            if (DIG_PARK) {                           // the digits of level LB + lb, used by this thread in the second batch
              const int sh1 = sh - LB * Bg_bit;
              dg[m & 15] = ((unsigned)(v0 >> sh1) & Q.dmask) | (((unsigned)(v1 >> sh1) & Q.dmask) << 16);
            }
          } else {
            u0 = (pk0[m] >> sh) & Q.dmask; u1 = (pk1[m] >> sh) & Q.dmask;
          }
          const double d0 = __hiloint2double(0x43300000, (int)u0) - Q.dbias;
          const double d1 = __hiloint2double(0x43300000, (int)u1) - Q.dbias;
          // fold z = d0 + i d1; the constant part of the twist, w^(64 m) = W_64^(WQ m), now or inside the butterflies (FOLD)
          x[m] = FOLD ? make_double2(d0, d1) : mul_w64(make_double2(d0, d1), WQ * m, false);
        }
        if constexpr (DIG_PARK) if (lev0 == 0) {
          tmem_st_u32x16(taddr + COL_DG, dg);
          asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
        }
        if constexpr (FOLD) reg_dft_fma<RA, WQ>(x); else reg_dif<RA>(x);
        double2 *row = buf + (pA * NB + lb) * M;
        tmem_wait(ta0);
        tmem_also(ta1);
#pragma unroll
        for (int pos = 0; pos < 8; ++pos)
          row[pos * 64 + ((pos & 1) ? qs1 : qs0)] = cmul(x[pos], tmem_c(pos < 4 ? ta0 : ta1, pos & 3));
        if (RA > 8) {
          tmem_issue(ta0, taddr + COL_TA + 32);
          tmem_issue(ta1, taddr + COL_TA + 48);
          tmem_wait(ta0);
          tmem_also(ta1);
#pragma unroll
          for (int pos = 8; pos < RA; ++pos)
            row[pos * 64 + ((pos & 1) ? qs1 : qs0)] = cmul(x[pos], tmem_c(pos < 12 ? ta0 : ta1, pos & 3));
        }
      }
      __syncthreads();
      // key position 4c + pos3 = 8c' + m' is stored at m' * (M/8) + c'; TRGSW row order of trgsw.c:394-419
      const double2 *__restrict__ kbase = key + (4 * halfC) * (M / 8) + cpC;
      auto load_half = [&](double2 (&dst)[4], int rb, int o) {        // key row of buffer rb, output polynomial o
        const int p = rb / NB, lev = lev0 + (rb - p * NB);
        const double2 *__restrict__ k0 = kbase + (size_t)((p * L + lev) * 2 + o) * M;
#pragma unroll
        for (int i = 0; i < 4; ++i) dst[i] = ldg_key(k0 + i * (M / 8));
      };
      // KPF: the first row's key values are requested before pass B (the Fourier accumulators are zero or parked, so the
      // registers are there), and inside pass C each half is re-requested for the next row as soon as its MACs are done
      double2 kv0[4], kv1[4];
      if (KPF) { load_half(kv0, 0, 0); load_half(kv1, 0, 1); }
      // ------------------------------- pass B: radix 16 in place, no twiddle --------------------------------------
      if constexpr (RB == 4) {
        double2 *blk = buf + rowB * M + pos1B * 64;
        double2 x[16];
#pragma unroll
        for (int m2 = 0; m2 < 16; ++m2) x[m2] = blk[swz_b(m2)];
        if constexpr (FOLD) reg_dft_fma<16, 0>(x); else reg_dif<16>(x);
#pragma unroll
        for (int pos = 0; pos < 16; ++pos) blk[swz_b(pos)] = x[pos];
      } else if constexpr (HYB) {
        // Two rows = 2 x M/16 full radix-16 tasks: the first M/64 warps take them (row = task / (M/16)), the others wait at
        // the block barrier.  No redundant loads and no half-filled FP64 instructions, at the price of two block barriers.
        __syncwarp();
        if (t_ < M / 8) {
          const int rowH = t_ / (M / 16), trH = t_ - rowH * (M / 16), pos1H = trH >> 2, rH = trH & 3, b4H = (pos1H & 1) << 2;
          auto swz_h = [&](int e) { return ((4 * e) ^ ((e >> 1) & 4) ^ b4H) + (rH ^ ((e >> 1) & 3)); };
          double2 *blk = buf + rowH * M + pos1H * 64;
          double2 x[16];
#pragma unroll
          for (int m2 = 0; m2 < 16; ++m2) x[m2] = blk[swz_h(m2)];
          if constexpr (FOLD) reg_dft_fma<16, 0>(x); else reg_dif<16>(x);
#pragma unroll
          for (int pos = 0; pos < 16; ++pos) blk[swz_h(pos)] = x[pos];
        }
        __syncthreads();
      } else {
        // Two rows = 16 radix-16 tasks per warp: each is split over the lane pair (lane, lane + 16) instead of leaving half
        // the warp idle (idle lanes still occupy the FP64 pipe: +12 % FP64 instructions, profiles/r2b).  Both lanes read
        // the 16 inputs; lane half h forms the first DIF stage for its half of the outputs branch-free,
        //   y_i = (x_i + s x_{i+8}) * w_i,   (s, w_i) = (+1, 1) for h = 0 and (-1, W_16^i) for h = 1,
        // finishes with a radix-8 DIF and writes outputs 8h .. 8h+7 (even / odd frequencies).
        double2 *blk = buf + rowB2 * M + pos1B * 64;
        double2 x[16], y[8];
#pragma unroll
        for (int m2 = 0; m2 < 16; ++m2) x[m2] = blk[swz_b(m2)];
        const double sg = hB ? -1.0 : 1.0;
#pragma unroll
        for (int i = 0; i < 8; ++i) y[i] = make_double2(fma(sg, x[i + 8].x, x[i].x), fma(sg, x[i + 8].y, x[i].y));
#pragma unroll
        for (int i = 1; i < 8; ++i) {
          double c, sn;
          w64_cs(4 * i, c, sn);
          const double cs = hB ? c : 1.0, ss = hB ? sn : 0.0;
          y[i] = make_double2(fma(y[i].x, cs, -y[i].y * ss), fma(y[i].x, ss, y[i].y * cs));
        }
        if constexpr (FOLD) reg_dft_fma<8, 0>(y); else reg_dif<8>(y);
        __syncwarp();                                 // every lane has read its 16 inputs before any output lands on them
#pragma unroll
        for (int j = 0; j < 8; ++j) blk[swz_b2(j)] = y[j];
      }
      __syncwarp();                                   // pass C below reads only the blocks this warp just transformed
      // ------------------------------- pass C + MAC ----------------------------------------------------------------
      {
        TmemLd tc, f0, f1;
        tmem_issue(tc, taddr + COL_TC);
        if (FA_PARK && lev0 > 0) { tmem_issue(f0, taddr + COL_FA0); tmem_issue(f1, taddr + COL_FA1); }
        tmem_wait(tc);
        if (FA_PARK && lev0 > 0) {
          tmem_also(f0); tmem_also(f1);
#pragma unroll
          for (int i = 0; i < 4; ++i) { fa[0][i] = tmem_c(f0, i); fa[1][i] = tmem_c(f1, i); }
        }
        double2 twc[4];
#pragma unroll
        for (int r = 1; r < 4; ++r) twc[r] = tmem_c(tc, r);
#pragma unroll
        for (int rb = 0; rb < RB; ++rb) {
          if (!KPF) { load_half(kv0, rb, 0); load_half(kv1, rb, 1); }
          const double2 *row = buf + rb * M + cbase;
          double2 x[4];
          x[0] = row[0 ^ cxor];
#pragma unroll
          for (int r = 1; r < 4; ++r) x[r] = cmul(row[r ^ cxor], twc[r]);
          reg_dif<4>(x);
#pragma unroll
          for (int i = 0; i < 4; ++i) cfma(fa[0][i], x[i], kv0[i]);
          if (KPF && rb + 1 < RB) load_half(kv0, rb + 1, 0);
#pragma unroll
          for (int i = 0; i < 4; ++i) cfma(fa[1][i], x[i], kv1[i]);
          if (KPF && rb + 1 < RB) load_half(kv1, rb + 1, 1);
        }
      }
      // the next batch's pass A overwrites every block of the row buffers; after the LAST batch the inverse starts in
      // this warp's own blocks (C', B'), so no block barrier is needed there
      if (lev0 + NB < L) {
        if (FA_PARK) {
          tmem_st4(taddr + COL_FA0, fa[0]); tmem_st4(taddr + COL_FA1, fa[1]);
          asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
        }
        __syncthreads();
      } else {
        __syncwarp();
      }
    };
#pragma unroll
    for (int lev0 = 0; lev0 + LB <= L; lev0 += LB) batch(std::integral_constant<int, LB>{}, lev0);
    if constexpr (L % LB != 0) batch(std::integral_constant<int, L % LB>{}, L - L % LB);

    // ---------------------------------- inverse: C' ----------------------------------------------------------------
    {
      double2 twc[4];
      tmem_ld4(twc, taddr + COL_TC);
#pragma unroll
      for (int o = 0; o < 2; ++o) {
        reg_dit_inv<4>(fa[o]);
        double2 *row = buf + o * M + cbase;
        row[0 ^ cxor] = fa[o][0];
#pragma unroll
        for (int r = 1; r < 4; ++r) row[r ^ cxor] = cmul_conj(fa[o][r], twc[r]);
      }
    }
    if constexpr (HYB) {
      // the LAST M/64 warps take the 2 x M/16 inverse tasks (the first ones took the forward two-row phase)
      __syncthreads();
      if (t_ >= T - M / 8) {
        const int th = t_ - (T - M / 8);
        const int rowH = th / (M / 16), trH = th - rowH * (M / 16), pos1H = trH >> 2, rH = trH & 3, b4H = (pos1H & 1) << 2;
        auto swz_h = [&](int e) { return ((4 * e) ^ ((e >> 1) & 4) ^ b4H) + (rH ^ ((e >> 1) & 3)); };
        double2 *blk = buf + rowH * M + pos1H * 64;
        double2 x[16];
#pragma unroll
        for (int pos = 0; pos < 16; ++pos) x[pos] = blk[swz_h(pos)];
        if constexpr (FOLD) reg_dit_inv_fma<16>(x); else reg_dit_inv<16>(x);
#pragma unroll
        for (int m2 = 0; m2 < 16; ++m2) blk[swz_h(m2)] = x[m2];
      }
    }
    __syncwarp();
    // ---------------------------------- B' (per warp, own blocks) -----------------------------------------------------
    if constexpr (!HYB) {
      // the mirror image of the split above: lane half h inverts the even (h = 0) / odd (h = 1) frequencies with a radix-8
      // DIT, the odd half applies conj(W_16^j), the lane pair swaps its 8 values and lane h forms outputs j + 8h = E_j +- O_j
      double2 *blk = buf + rowB2 * M + pos1B * 64;
      double2 y[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) y[j] = blk[swz_b2(j)];
      if constexpr (FOLD) reg_dit_inv_fma<8>(y); else reg_dit_inv<8>(y);
#pragma unroll
      for (int j = 1; j < 8; ++j) {
        double c, sn;
        w64_cs(4 * j, c, sn);
        const double cs = hB ? c : 1.0, ss = hB ? -sn : 0.0;        // conj(W_16^j)
        y[j] = make_double2(fma(y[j].x, cs, -y[j].y * ss), fma(y[j].x, ss, y[j].y * cs));
      }
      const double sg = hB ? -1.0 : 1.0;
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const double rx = __shfl_xor_sync(0xffffffffu, y[j].x, 16), ry = __shfl_xor_sync(0xffffffffu, y[j].y, 16);
        blk[swz_b2(j)] = make_double2(fma(sg, y[j].x, rx), fma(sg, y[j].y, ry));   // h = 0: E + O at j; h = 1: E - O at j + 8
      }
    }
    __syncthreads();
    // ---------------------------------- A' + accumulate ----------------------------------------------------------------
    // (N = 2048: only half of the 256 threads have a row here.  Two ways of employing the other half were measured SLOWER,
    // because both add work to pipes that the SM's other CTA would otherwise use: the radix-16 inverse split over lane pairs
    // (+54 % instructions in this phase for selects, shuffles and extra tensor-memory loads: 117.2 vs 112.3 ms, profiles/r2k)
    // and the outputs split by parity over the two slots of a polynomial (both load and untwiddle all 16 positions, then one
    // radix-8 inverse each: +18 % instructions in this phase, 114.8 vs 112.2 ms, profiles/r2q))
    if (SLOTS == 2 || slot < 2) {
      const double2 *row = buf + pA * M;
      double2 x[RA];
#pragma unroll
      for (int g8 = 0; g8 < RA / 8; ++g8) {
        TmemLd ta0, ta1;
        tmem_issue(ta0, taddr + COL_TA + 32 * g8);
        tmem_issue(ta1, taddr + COL_TA + 32 * g8 + 16);
#pragma unroll
        for (int i = 0; i < 8; ++i) x[8 * g8 + i] = row[(8 * g8 + i) * 64 + ((i & 1) ? qs1 : qs0)];
        tmem_wait(ta0);
        tmem_also(ta1);
#pragma unroll
        for (int i = 0; i < 8; ++i) x[8 * g8 + i] = cmul_conj(x[8 * g8 + i], tmem_c(i < 4 ? ta0 : ta1, i & 3));
      }
      if constexpr (FOLD) reg_dit_inv_fma<RA>(x); else reg_dit_inv<RA>(x);
      u64 *ap = acc + pA * N;
      u64 ownA[8];
#pragma unroll
      for (int m = 0; m < RA; ++m) {
        const double2 z = mul_w64(x[m], WQ * m, true);
        const int j = qA + m * 64;
        const u64 d0 = f64_to_torus_scaled(z.x, Q.out_scale), d1 = f64_to_torus_scaled(z.y, Q.out_scale);
        if (TMEM_ACC) {
          if ((m & 3) == 0) tmem_ld_u64x8(ownA, taddr + COL_ACC + 4 * m);
          const u64 n0 = ownA[2 * (m & 3)] + d0, n1 = ownA[2 * (m & 3) + 1] + d1;
          ownA[2 * (m & 3)] = n0; ownA[2 * (m & 3) + 1] = n1;
          ap[j] = n0;                                 // shared copy: read (rotated) by other threads in the next step
          ap[j + M] = n1;
          if ((m & 3) == 3) tmem_st_u64x8(taddr + COL_ACC + 4 * (m - 3), ownA);
        } else {
          ap[j] += d0;                                // trlwe_from_DFT + trlwe_addto
          ap[j + M] += d1;
        }
      }
      if (TMEM_ACC) asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
    }
    __syncthreads();
  }

  // every thread is past its last tensor-memory access before the columns are released (also when no step ran)
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  if (tid < 32) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base_s), "n"(TMEM_COLS) : "memory");
  // ---- epilogue: sample extraction at index 0 (trlwe.c:540-552) or the raw accumulator -----------------------------
  if (A.extract) {
    u64 *o = A.out + (size_t)ct * (N + 1);
    for (int c = tid; c < N; c += T) o[c] = (c == 0) ? acc[0] : (0ull - acc[N - c]);
    if (tid == 0) o[N] = acc[N];
  } else {
    u64 *o = A.out + (size_t)ct * 2 * N;
    for (int c = tid; c < 2 * N; c += T) o[c] = acc[c];
  }
}

// ---- tables: TA[RA][64] then TC[16][4] -----------------------------------------------------------------------------
static std::mutex g_k1q_mu;
static std::map<long long, double2 *> g_k1q_tab;     // (device, N) -> table

static const double2 *k1q_tables_for(int N) {
  ensure_init();
  std::lock_guard<std::mutex> lk(g_k1q_mu);
  auto it = g_k1q_tab.find(dev_key(N));
  if (it != g_k1q_tab.end()) return it->second;
  const int M = N / 2, RA = M / 64;
  std::vector<double2> h((size_t)RA * 64 + 64);
  for (int k1 = 0; k1 < RA; ++k1)
    for (int q = 0; q < 64; ++q) {
      // w^q * W_M^(q*k1) = exp(i*pi*q*(4*k1+1)/N)
      const long double ang = M_PIl * (long double)((long long)q * (4 * k1 + 1)) / (long double)N;
      h[(size_t)k1 * 64 + q] = make_double2((double)cosl(ang), (double)sinl(ang));
    }
  for (int k2 = 0; k2 < 16; ++k2)
    for (int r = 0; r < 4; ++r) {
      const long double ang = 2.0L * M_PIl * (long double)(r * k2) / 64.0L;   // W_64^(r k2)
      h[(size_t)RA * 64 + k2 * 4 + r] = make_double2((double)cosl(ang), (double)sinl(ang));
    }
  double2 *d = nullptr;
  MB_CHECK(cudaMalloc(&d, sizeof(double2) * h.size()));
  MB_CHECK(cudaMemcpy(d, h.data(), sizeof(double2) * h.size(), cudaMemcpyHostToDevice));
  MB_CHECK(cudaDeviceSynchronize());   // pageable H2D + non-blocking compute streams: fence once
  g_k1q_tab[dev_key(N)] = d;
  return d;
}

template <int LOGM, int L, int LB, bool PKALL, int VAR>
static void launch_q(const K1QArgs &a, int count, cudaStream_t st) {
  constexpr int M = 1 << LOGM;
  const size_t smem = (size_t)2 * 2 * M * 8 + (size_t)2 * LB * M * 16 + (((size_t)a.a.size * 2 + 15) & ~(size_t)15);
  static size_t configured_dev[MB_MAX_DEV] = {0};          // function attributes are per device
  size_t &configured = configured_dev[current_device()];
  if (smem > configured) {
    MB_REQUIRE(smem <= 227 * 1024, "k1q kernel: %zu B of shared memory needed (blind rotation too long)", smem);
    MB_CHECK(cudaFuncSetAttribute(blind_rotate_k1q_kernel<LOGM, L, LB, PKALL, VAR>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    configured = smem;
  }
  blind_rotate_k1q_kernel<LOGM, L, LB, PKALL, VAR><<<count, M / 4, smem, st>>>(a);
  MB_CHECK(cudaGetLastError());
  count_launch();
}

// default variant bits (see the kernel): key pipelining + hybrid two-row pass B + folded butterflies.  Measured on B200, 4096
// ciphertexts (profiles/r2g / r2h / r2l_k1q_variants.log): level 1 var 0 / 1 / 3 / 7 = 41.1 / 40.6 / 39.0 / 38.75 ms, level 2 (three
// key segments) var 3 / 7 = 112.2 / 111.1 ms.  The folded butterflies (bit 2) remove 7 % of the FP64 instructions and gain under 1 %:
// the kernel is bound by the L1 / shared-memory data pipe and by barrier skew, not by the FP64 pipe.
constexpr int K1Q_VAR = 7;

// levels per shared-memory batch: 2 when two levels' digits fit the 32-bit packed word, else 1
static int k1q_lb(int l, int Bg_bit) { return (l >= 2 && 2 * Bg_bit <= 32) ? 2 : 1; }

bool k1q_supported(const Params &p) {
  if (p.k != 1 || !(p.N == 1024 || p.N == 2048)) return false;
  return p.l >= 1 && p.l <= 4 && p.Bg_bit >= 1 && p.Bg_bit <= 31;
}

void k1q_variant_name(const Params &p, char *dst, size_t cap) {
  snprintf(dst, cap, "k1q<N=%d,l=%d,lb=%d,T=%d>", p.N, p.l, k1q_lb(p.l, p.Bg_bit), p.N / 8);
}

void launch_blind_rotate_k1q(const BlindRotateLaunch &b, cudaStream_t st) {
  const Params &p = b.bsk->p;
  MB_REQUIRE(k1q_supported(p) && !b.direct, "k1q kernel: unsupported parameters");
  upload_w64();
  K1QArgs qa;
  K1Args &a = qa.a;
  qa.off = 1ull << (64 - p.l * p.Bg_bit - 1);
  for (int i = 0; i < p.l; ++i) qa.off += 1ull << (64 - i * p.Bg_bit - 1);
  qa.dbias = 4503599627370496.0 + (double)(1ull << (p.Bg_bit - 1));
  qa.out_scale = 5.42101086242752217e-20 / (double)(p.N / 2);
  qa.dmask = (unsigned)((1ull << p.Bg_bit) - 1ull);
  qa.b_index = b.b_index > 0 ? b.b_index : b.size;
  a.bsk = b.bsk->d; a.tab = k1q_tables_for(p.N); a.tv = b.tv; a.tv_count = b.tv_count; a.in = b.in;
  a.in_stride = b.in_stride; a.in_div = b.in_div > 0 ? b.in_div : 1; a.size = b.size; a.out = b.out; a.extract = b.extract;
  a.init_rotate = b.init_rotate; a.prec_offset = b.prec_offset; a.preprocess = b.preprocess; a.kappa = b.kappa; a.theta = b.theta;
  a.Bg_bit = p.Bg_bit; a.count = b.count;
  const int logm = ilog2i(p.N) - 1, lb = k1q_lb(p.l, p.Bg_bit);
  const bool pkall = p.l * p.Bg_bit <= 32;
  // experiment knob (benchmark shapes only): MB200_K1Q_VAR = variant bits
  if (const char *e = getenv("MB200_K1Q_VAR")) {
    const int var = atoi(e);
#define MB_K1Q_VARCASE(V_) \
    if (var == V_ && var != K1Q_VAR && logm == 9 && p.l == 3 && lb == 2 && pkall) { launch_q<9, 3, 2, true, V_>(qa, b.count, st); return; } \
    if (var == V_ && var != K1Q_VAR && logm == 10 && p.l == 4 && lb == 2 && !pkall) { launch_q<10, 4, 2, false, V_>(qa, b.count, st); return; }
    MB_K1Q_VARCASE(0) MB_K1Q_VARCASE(1) MB_K1Q_VARCASE(3) MB_K1Q_VARCASE(5) MB_K1Q_VARCASE(7)
#undef MB_K1Q_VARCASE
  }
#define MB_K1Q_CASE(LM, LL, LBB) \
  if (logm == LM && p.l == LL && lb == LBB) { \
    if (pkall) launch_q<LM, LL, LBB, true, K1Q_VAR>(qa, b.count, st); else launch_q<LM, LL, LBB, false, K1Q_VAR>(qa, b.count, st); \
    return; }
  MB_K1Q_CASE(9, 1, 1) MB_K1Q_CASE(9, 2, 2) MB_K1Q_CASE(9, 2, 1) MB_K1Q_CASE(9, 3, 2) MB_K1Q_CASE(9, 3, 1) MB_K1Q_CASE(9, 4, 2) MB_K1Q_CASE(9, 4, 1)
  MB_K1Q_CASE(10, 1, 1) MB_K1Q_CASE(10, 2, 2) MB_K1Q_CASE(10, 2, 1) MB_K1Q_CASE(10, 3, 2) MB_K1Q_CASE(10, 3, 1) MB_K1Q_CASE(10, 4, 2) MB_K1Q_CASE(10, 4, 1)
#undef MB_K1Q_CASE
  MB_FATAL("k1q kernel: no instantiation for N=%d l=%d lb=%d", p.N, p.l, lb);
}

}  // namespace mb
