// k1_kernel.cuh -- the k = 1 blind-rotation / external-product kernel template (see blind_rotate_k1.cu for the
// design notes).  Included by blind_rotate_k1.cu (DIRECT = false: the bootstrap loop) and
// blind_rotate_k1_direct.cu (DIRECT = true: ONE external product, optionally as a CMUX).
#pragma once
#include <type_traits>

#include "k1_common.cuh"

namespace mb {

// G ciphertexts per CTA (G*T threads, each group of T threads owns one ciphertext and its own shared
// memory region).  The groups run in lockstep (block barriers), so their loads of the same key row
// are issued within one L2 round trip of each other and merge in L1: the key streams from L2 once per
// CTA instead of once per ciphertext (ablation: key loads are 19 % / 27 % of the kernel at level 1 / 2).
// G > 1 is an experiment knob (MB200_K1_G): it measured slower than G = 1, see launch_blind_rotate_k1.
#ifdef MB200_K1_MAXNREG
#define MB200_K1_BOUNDS __maxnreg__(MB200_K1_MAXNREG)       // experiment: explicit register cap (build-wide)
#else
#define MB200_K1_BOUNDS __launch_bounds__(G * (1 << LOGM) / 8, MINB)
#endif
// DIRECT: one external product instead of the rotation loop (trgsw_mul_trlwe_DFT + trlwe_from_DFT, trgsw.c:385 /
// trlwe.c:629, or the CMUX of vertical_packing.c:24-33): the shared-memory accumulator starts as the OPERAND
// tv - in1, its digits are taken as they are (no X^a - 1), the TRGSW is number sel_const / sel[ct] of the
// resident set, and the result is in1 + product (in1 may be null: plain external product).
template <int LOGM, int L, int LB, int MINB, bool PKALL, int PF, int G, bool DIRECT = false>
__global__ void MB200_K1_BOUNDS blind_rotate_k1_kernel(K1Args A) {
  constexpr int M = 1 << LOGM, N = 2 * M, S = M / 16, R2 = M / 128, T = M / 8, C8 = M / 8;
  constexpr int LOGR2 = clog2(R2);
  constexpr int ROWS = 2 * L, ROWS_B = 2 * LB;    // ROWS_B: shared-memory row buffers (largest batch)
  constexpr int PKL = PKALL ? L : LB;                 // gadget levels packed into one 32-bit word per coefficient
#ifdef MB200_PB_FULL
  constexpr int PB_UNROLL = 16;
#elif defined(MB200_PB_UNROLL)
  constexpr int PB_UNROLL = MB200_PB_UNROLL;
#else
  // independent pass-B butterflies in flight per thread; 4 measured 3 % slower than 2 at N = 1024 (code size:
  // profiles/r1k_k1_occupancy.log)
  constexpr int PB_UNROLL = 2;
#endif
#ifndef MB200_PA_UNROLL
#define MB200_PA_UNROLL 1
#endif
#ifndef MB200_PC_UNROLL
#define MB200_PC_UNROLL 2
#endif
  constexpr int PC_UNROLL = MB200_PC_UNROLL;          // pass-C rows unrolled together when keys are not double buffered
  constexpr int PA_UNROLL = MB200_PA_UNROLL;          // gadget levels of pass A unrolled together
  static_assert(LB >= 1 && LB <= L, "levels per batch");
  static_assert(R2 >= 2 && R2 <= 16, "supported N: 512..4096");

  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int grp = (G > 1) ? threadIdx.x / T : 0;
  const int tid = (G > 1) ? threadIdx.x - grp * T : threadIdx.x;
  const int ct_raw = blockIdx.x * G + grp;
  const bool live = ct_raw < A.count;
  const int ct = live ? ct_raw : A.count - 1;        // surplus groups of the last CTA shadow a real ciphertext
  const size_t region = (size_t)2 * N * 8 + (size_t)ROWS_B * M * 16 + (((size_t)A.size * 2 + 15) & ~(size_t)15);
  u64 *acc = reinterpret_cast<u64 *>(smem_raw + grp * region);       // [2][N]
  double2 *buf = reinterpret_cast<double2 *>(acc + 2 * N);           // [ROWS_B][M]
  unsigned short *rot = reinterpret_cast<unsigned short *>(buf + ROWS_B * M);   // [size] rotation amounts

  const int log_N2 = LOGM + 2;
  const double2 *__restrict__ TA = A.tab;
  const double2 *__restrict__ TB = A.tab + 16 * S;
  const u64 *in = DIRECT ? nullptr : A.in + (size_t)(ct / A.in_div) * A.in_stride;
  const u64 *tv = A.tv + (size_t)(A.tv_count > 1 ? ct % A.tv_count : 0) * 2 * N;
  const int Bg_bit = A.Bg_bit;

#ifndef MB200_NO_TMEM_TW
  constexpr bool TMEM_TW = (T <= 128) && (G == 1);    // one lane per thread: warps 0..3 of the CTA own TMEM lanes 32w..32w+31
#else
  constexpr bool TMEM_TW = false;
#endif
  // pass-B / B' twiddles too when they fill whole 16-word groups (R2 = 4, 8): columns 64 .. 64 + 4*R2
  // ... and the 32 accumulator words a thread owns (columns 64..127), so that only the ROTATED reads of pass A go to
  // shared memory; N >= 1024 only (at N = 512 up to 8 CTAs share the SM's 512 columns)
  // ... and, at N = 1024 where a row's pass-A threads are exactly one warp, the pass A -> pass B exchange itself: the
  // warp parks its 16 x 32 transformed values (real / imaginary parts in separate column groups) and reads them back
  // with the 16x256b shape, which hands thread t the stride-8 lanes t/4 + 8m of a radix-4 group.  That takes the
  // pass-A stores and the pass-B loads (a third of the kernel's L1 wavefronts) and one block barrier out of the batch.
#ifdef MB200_TMEM_XB
  constexpr bool TMEM_XB = TMEM_TW && LOGM == 9;
#else
  constexpr bool TMEM_XB = false;
#endif
  constexpr int COL_X = 64;
  // ... and, when the l levels do not fit one 32-bit word per coefficient (PKALL = false) but the levels after the first
  // batch do, the low digit word of every coefficient: the digits are then extracted ONCE per step (one pass over the
  // accumulator, rotated reads included) instead of once per batch.  Needs 32 spare columns: N >= 2048 (256 per CTA).
#ifndef MB200_NO_TMEM_PK
  constexpr bool TMEM_PK = TMEM_TW && !PKALL && !DIRECT && LOGM >= 9 && !TMEM_XB && (L - LB) <= LB && (L - LB) >= 1;
#else
  constexpr bool TMEM_PK = false;
#endif
  // column budget: 256 per CTA at N = 2048 (everything fits); 128 at N = 1024, where the digit words (worth 3.6 %)
  // take the place of the owned accumulator words (worth 1.8 %)
  constexpr int COL_PK = LOGM >= 10 ? 160 : 64;
#ifndef MB200_NO_TMEM_ACC
  constexpr bool TMEM_ACC = TMEM_TW && !DIRECT && LOGM >= 9 && !TMEM_XB && !(TMEM_PK && LOGM == 9);
#else
  constexpr bool TMEM_ACC = false;
#endif
  constexpr int COL_ACC = 64, COL_TB = TMEM_ACC ? 128 : 64;
  constexpr bool TMEM_TB = TMEM_TW && !TMEM_XB && (R2 == 8 || (R2 == 4 && !TMEM_ACC && !TMEM_PK));   // 4 CTAs x 128 columns at N = 1024
  constexpr int TMEM_COLS = (TMEM_PK && LOGM >= 10) ? 256 : (TMEM_PK || TMEM_XB) ? 128 : TMEM_TB ? (TMEM_ACC ? 256 : 128) : (TMEM_ACC ? 128 : 64);
  static_assert(!(TMEM_PK && LOGM >= 10) || (COL_TB + 32 <= COL_PK), "tensor-memory column layout");
  static_assert(!(TMEM_PK && TMEM_ACC && COL_PK == COL_ACC), "tensor-memory column layout");
  __shared__ unsigned tmem_base_s;
  unsigned tw_taddr = 0;
  if (TMEM_TW) {
    if (tid < 32) {
      asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;\n\t"
                   "tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;"
                   ::"r"((unsigned)__cvta_generic_to_shared(&tmem_base_s)), "n"(TMEM_COLS) : "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    tw_taddr = tmem_base_s + ((unsigned)(tid >> 5) << 21);            // lane field (bits 31:16) = 32 * warp
    const int qA0 = tid % (M / 16);
#pragma unroll
    for (int g4 = 0; g4 < 4; ++g4) {
      double2 tw[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) tw[i] = __ldg(&A.tab[brev(4 * g4 + i, 4) * (M / 16) + qA0]);
      tmem_st4(tw_taddr + 16 * g4, tw);
    }
    if (TMEM_TB) {
#pragma unroll
      for (int g4 = 0; g4 < R2 / 4; ++g4) {
        double2 tw[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) tw[i] = __ldg(&A.tab[16 * (M / 16) + (4 * g4 + i) * 8 + (tid & 7)]);   // TB[k*8 + qpB]
        tmem_st4(tw_taddr + COL_TB + 16 * g4, tw);
      }
    }
    asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
  }

  // ---- initial accumulator: tv * X^(2N - round((b + 1/(4*torus_base)) * 2N))  (bootstrap.c:194-195)
  int rot0 = 0;
  if (A.init_rotate) {
    u64 b = in[A.size];
    if (A.preprocess) b = pb_preprocess(b, A.kappa, A.theta, log_N2);
    rot0 = (2 * N - (int)torus2int(b + A.prec_offset, log_N2)) & (2 * N - 1);
  }
  if (DIRECT) {
    const u64 *sub = A.in1 ? A.in1 + (size_t)ct * 2 * N : nullptr;
    for (int c = tid; c < 2 * N; c += T) acc[c] = sub ? tv[c] - sub[c] : tv[c];
  } else {
    for (int c = tid; c < 2 * N; c += T) {
      const int p = c / N, i = c - p * N;
      acc[c] = rot0 ? rotated_coeff(tv + (size_t)p * N, i, rot0, N) : tv[c];
    }
  }
  // all rotation amounts up front: round(a_i * 2N / 2^64) (bootstrap.c:113), one 16-bit word per step
  for (int i = tid; !DIRECT && i < A.size; i += T) {
    u64 av = in[i];
    if (A.preprocess) av = pb_preprocess(av, A.kappa, A.theta, log_N2);
    rot[i] = (unsigned short)(torus2int(av, log_N2) & (2 * N - 1));
  }
  __syncthreads();
  if (TMEM_ACC) {                                     // park the accumulator words this thread owns (pass A / A' ownership)
    const int pA0 = tid / (M / 16), qA0 = tid - pA0 * (M / 16);
    const u64 *ap0 = acc + pA0 * N;
#pragma unroll
    for (int g4 = 0; g4 < 4; ++g4) {
      u64 v[8];
#pragma unroll
      for (int i = 0; i < 4; ++i) { v[2 * i] = ap0[qA0 + (4 * g4 + i) * (M / 16)]; v[2 * i + 1] = ap0[qA0 + (4 * g4 + i) * (M / 16) + M]; }
      tmem_st_u64x8(tw_taddr + COL_ACC + 16 * g4, v);
    }
    asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
  }

  const u64 off = decomp_offset(Bg_bit, L);
  const unsigned dmask = (1u << Bg_bit) - 1u;
  // digit u in [0, Bg) -> double(u - Bg/2) = hiloint2double(0x43300000, u) - (2^52 + Bg/2), exact
  const double dbias = 4503599627370496.0 + (double)(1 << (Bg_bit - 1));
  const double inv_M = 1.0 / (double)M;
  const int pA = tid / S, qA = tid - pA * S;          // pass A / A' ownership
  const int qpB = tid & 7;                            // pass B twiddle column (T is a multiple of 8)
  // Swizzled addresses spelled out so that they are (thread constant) + (compile-time constant):
  //   swz(s) = s ^ ((s >> 3) & 7) only permutes the low 3 bits, by a mask that depends on s >> 3.
  // pass A / A': element pos*S + qA -> mask ((pos*(S/8)) + (qA>>3)) & 7: NVA distinct masks
  constexpr int S8 = S / 8, NVA = (S8 >= 8) ? 1 : 8 / S8;
  int qsw[NVA];
#pragma unroll
  for (int v = 0; v < NVA; ++v) qsw[v] = qA ^ ((v * S8 + (qA >> 3)) & 7);
  // pass B / B': element b*S + 8m + qp with b = (tid>>3) + it*(T/8): mask ((tid>>3)*S8 + m) & 7 (it drops out)
  int qx[R2];
#pragma unroll
  for (int m = 0; m < R2; ++m) qx[m] = qpB ^ ((((tid >> 3) * S8) + m) & 7);
  const int bB0 = (tid >> 3) * S;                     // first element of this thread's pass-B block
  // pass-B twiddles W_S^(qp*k), k < R2: from tensor memory (TMEM_TB) or global memory
  auto load_tb = [&](double2 (&tb)[R2]) {
    if (TMEM_TB) {
#pragma unroll
      for (int g4 = 0; g4 < R2 / 4; ++g4) {
        double2 tw[4];
        tmem_ld4(tw, tw_taddr + COL_TB + 16 * g4);
#pragma unroll
        for (int i = 0; i < 4; ++i) tb[(4 * g4 + i) & (R2 - 1)] = tw[i];
      }
    } else {
#pragma unroll
      for (int k = 1; k < R2; ++k) tb[k] = __ldg(&TB[k * 8 + qpB]);
    }
  };

  for (int step = 0; step < (DIRECT ? 1 : A.size); ++step) {
    const int a_i = DIRECT ? 0 : rot[step];
    // bootstrap.c:114 skips a_i == 0.  With several ciphertexts in lockstep the step is executed
    // instead: (X^0 - 1)*acc = 0 decomposes into all-zero digits, so the accumulator is unchanged.
    if (!DIRECT && G == 1 && a_i == 0) continue;
    const int key_idx = DIRECT ? (A.sel_const >= 0 ? A.sel_const : A.sel[ct]) : step;
    const double2 *__restrict__ key = A.bsk + (size_t)key_idx * ROWS * 2 * M;

    double2 fa[2][8];                                 // Fourier accumulators: positions 8*tid .. 8*tid+7
#pragma unroll
    for (int pp = 0; pp < 2; ++pp)
#pragma unroll
      for (int i = 0; i < 8; ++i) fa[pp][i] = make_double2(0.0, 0.0);

    unsigned pk0[16], pk1[16];
    // (X^a - 1)*acc + rounding offset, top PKL*Bg_bit bits (= PKL signed digits) per coefficient
    auto pack_digits = [&](int lev_end) {
      const u64 *ap = acc + pA * N;
      const int pk_shift = 64 - lev_end * Bg_bit;
      const int base = (qA - a_i) & (2 * N - 1);       // index of coefficient qA in acc * X^a (sign in bit log2 N)
      u64 own[8];                                      // TMEM_ACC: acc[j], acc[j + M] of 4 consecutive m
      unsigned plo[16];                                // TMEM_PK: low digit words of 8 consecutive m
#pragma unroll
      for (int m = 0; m < 16; ++m) {
        const int j = qA + m * S;
        if (DIRECT) {                                  // the operand itself (trgsw.c:396-400)
          pk0[m] = (unsigned)((off + ap[j]) >> pk_shift);
          pk1[m] = (unsigned)((off + ap[j + M]) >> pk_shift);
          continue;
        }
        const int s0 = (base + m * S) & (2 * N - 1), s1 = (s0 + M) & (2 * N - 1);
        const u64 r0 = ap[s0 & (N - 1)], r1 = ap[s1 & (N - 1)];
        if (TMEM_ACC && (m & 3) == 0) tmem_ld_u64x8(own, tw_taddr + COL_ACC + 4 * m);
        const u64 t0 = off - (TMEM_ACC ? own[2 * (m & 3)] : ap[j]), t1 = off - (TMEM_ACC ? own[2 * (m & 3) + 1] : ap[j + M]);
        const u64 v0 = (s0 & N) ? t0 - r0 : t0 + r0;
        const u64 v1 = (s1 & N) ? t1 - r1 : t1 + r1;
        if (TMEM_PK) {
          // w = top L*Bg_bit bits; registers keep the first batch's levels, tensor memory the remaining ones
          const int lo_bits = (L - LB) * Bg_bit;
          const u64 w0 = v0 >> pk_shift, w1 = v1 >> pk_shift;
          pk0[m] = (unsigned)(w0 >> lo_bits);
          pk1[m] = (unsigned)(w1 >> lo_bits);
          plo[2 * (m & 7)] = (unsigned)w0 & ((1u << lo_bits) - 1u);
          plo[2 * (m & 7) + 1] = (unsigned)w1 & ((1u << lo_bits) - 1u);
          if ((m & 7) == 7) tmem_st_u32x16(tw_taddr + COL_PK + 2 * (m - 7), plo);
        } else {
          pk0[m] = (unsigned)(v0 >> pk_shift);
          pk1[m] = (unsigned)(v1 >> pk_shift);
        }
      }
      if (TMEM_PK) asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
    };
    if (PKALL || TMEM_PK) pack_digits(L);

    // One batch = NB gadget levels of both input polynomials (2*NB rows of shared-memory buffers).
    auto batch = [&](auto nb_tag, const int lev0) {
      constexpr int NB = decltype(nb_tag)::value, ROWS_B = 2 * NB;
      // ------------------------------- pass A -------------------------------------------------
      if (!PKALL && !TMEM_PK) pack_digits(lev0 + NB);
      if (TMEM_PK && lev0 > 0) {                        // the remaining levels' digit words, parked by pack_digits(L)
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          unsigned w[16];
          tmem_ld_u32x16(w, tw_taddr + COL_PK + 16 * h);
#pragma unroll
          for (int i = 0; i < 8; ++i) { pk0[8 * h + i] = w[2 * i]; pk1[8 * h + i] = w[2 * i + 1]; }
        }
      }
      // pass-A twiddles w^q * W_M^(q*k1): the same 16 values for every level of the batch -- loaded once
      // (the L1/shared-memory data pipe, not FP64, is the busiest unit of this kernel: ncu r1e)
      constexpr bool HOIST_TW = (LOGM <= 9) && !TMEM_TW;   // N = 2048+: registers are needed elsewhere (pass C key buffers)
      double2 twA[HOIST_TW ? 16 : 1];
      if (HOIST_TW) {
#pragma unroll
        for (int pos = 0; pos < 16; ++pos) twA[HOIST_TW ? pos : 0] = __ldg(&TA[brev(pos, 4) * S + qA]);
      }
#pragma unroll(PA_UNROLL)
      for (int lb = 0; lb < NB; ++lb) {
        const int sh = ((PKALL || (TMEM_PK && lev0 > 0)) ? (L - 1 - lev0 - lb) : (NB - 1 - lb)) * Bg_bit;
        double2 x[16];
#pragma unroll
        for (int m = 0; m < 16; ++m) {
          const double d0 = __hiloint2double(0x43300000, (int)((pk0[m] >> sh) & dmask)) - dbias;
          const double d1 = __hiloint2double(0x43300000, (int)((pk1[m] >> sh) & dmask)) - dbias;
          // fold z = d0 + i*d1 and the constant part of the twist, w^(m*M/16) = W_64^m
          x[m] = mul_w64(make_double2(d0, d1), m, false);
        }
        reg_dif<16>(x);
        double2 *row = buf + (pA * NB + lb) * M;
        if constexpr (TMEM_XB) {
          // twiddle, park in tensor memory: column COL_X + 2*pos = Re, COL_X + 32 + 2*pos = Im of element (block pos, lane qA)
#pragma unroll
          for (int g4 = 0; g4 < 4; ++g4) {
            double2 tw[4];
            tmem_ld4(tw, tw_taddr + 16 * g4);
#pragma unroll
            for (int i = 0; i < 4; ++i) x[4 * g4 + i] = cmul(x[4 * g4 + i], tw[i]);
          }
#pragma unroll
          for (int h = 0; h < 2; ++h) {
            tmem_st_f64x8(tw_taddr + COL_X + 16 * h, x[8 * h].x, x[8 * h + 1].x, x[8 * h + 2].x, x[8 * h + 3].x,
                          x[8 * h + 4].x, x[8 * h + 5].x, x[8 * h + 6].x, x[8 * h + 7].x);
            tmem_st_f64x8(tw_taddr + COL_X + 32 + 16 * h, x[8 * h].y, x[8 * h + 1].y, x[8 * h + 2].y, x[8 * h + 3].y,
                          x[8 * h + 4].y, x[8 * h + 5].y, x[8 * h + 6].y, x[8 * h + 7].y);
          }
          asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
          // pass B of this row by the same warp: thread (qB, cB) = (lane / 4, lane % 4) takes the radix-4 groups
          // (block p0 + cB, lanes qB + 8m), p0 = 0, 4, 8, 12, and writes the row buffer pass C reads
          const int lane = tid & 31, qB = lane >> 2, cB = lane & 3;
          double2 tbx[4];
#pragma unroll
          for (int k = 1; k < 4; ++k) tbx[k] = __ldg(&TB[k * 8 + qB]);
          {
            // all four groups of the thread at once: 4 x (16x256b.x4) = the whole 32-column Re / Im group groups of lanes
            // qB, qB + 8 (address lane 0) and qB + 16, qB + 24 (address lane 16); window i holds block 4i + cB
            double r0[4], r1[4], r2[4], r3[4], i0[4], i1[4], i2[4], i3[4];
            tmem_ld_16x256_x4(r0, r1, tw_taddr + COL_X);
            tmem_ld_16x256_x4(r2, r3, tw_taddr + (16u << 16) + COL_X);
            tmem_ld_16x256_x4(i0, i1, tw_taddr + COL_X + 32);
            tmem_ld_16x256_x4(i2, i3, tw_taddr + (16u << 16) + COL_X + 32);
            tmem_wait_ld();
#pragma unroll
            for (int g = 0; g < 4; ++g) {
              double2 y[4] = {make_double2(r0[g], i0[g]), make_double2(r1[g], i1[g]), make_double2(r2[g], i2[g]),
                              make_double2(r3[g], i3[g])};
              reg_dif<4>(y);
              const int b = 4 * g + cB;
#pragma unroll
              for (int pos = 0; pos < 4; ++pos) {
                const int k = brev(pos, 2);
                row[b * S + 8 * pos + (qB ^ ((4 * b + pos) & 7))] = k == 0 ? y[pos] : cmul(y[pos], tbx[k]);
              }
            }
          }
        } else if (TMEM_TW) {
#pragma unroll
          for (int g4 = 0; g4 < 4; ++g4) {
            double2 tw[4];
            tmem_ld4(tw, tw_taddr + 16 * g4);
#pragma unroll
            for (int i = 0; i < 4; ++i) row[(4 * g4 + i) * S + qsw[(4 * g4 + i) & (NVA - 1)]] = cmul(x[4 * g4 + i], tw[i]);
          }
        } else {
#pragma unroll
          for (int pos = 0; pos < 16; ++pos) {
            const double2 t = HOIST_TW ? twA[HOIST_TW ? pos : 0] : __ldg(&TA[brev(pos, 4) * S + qA]);
            row[pos * S + qsw[pos & (NVA - 1)]] = cmul(x[pos], t);
          }
        }
      }
      if constexpr (!TMEM_XB) __syncthreads();
      // key rows of this batch: row index of buffer rb
      auto key_row = [&](int rb) {
        const int p = rb / NB, lev = lev0 + (rb - p * NB);
        return key + (size_t)((p * L + lev) * 2) * M + tid;            // TRGSW row order of trgsw.c:394-419
      };
      double2 kv[PF == 1 ? 2 : 1][16];
      auto load_keys = [&](double2 (&dst)[16], int rb) {
        const double2 *__restrict__ k0 = key_row(rb);
#pragma unroll
#ifdef MB200_ABL_NOKEY
        for (int i = 0; i < 8; ++i) { dst[i] = make_double2(1.0 + i, 0.5 * rb); dst[8 + i] = make_double2(0.25 * i, 2.0 + rb); }
        (void)k0;
#else
        for (int i = 0; i < 8; ++i) {
          if (G > 1 || PF == 2) { dst[i] = __ldg(k0 + i * C8); dst[8 + i] = __ldg(k0 + M + i * C8); }   // L1-allocating
          else { dst[i] = ldg_key(k0 + i * C8); dst[8 + i] = ldg_key(k0 + M + i * C8); }          // streaming
        }
#endif
      };
      // PF == 2: no register double buffer; the next row is pulled into L1 (one lane per 128-byte line) while
      // the current one computes, so the demand loads hit L1 instead of waiting for L2
      auto prefetch_keys = [&](int rb) {
        if ((tid & 7) == 0) {
          const double2 *__restrict__ k0 = key_row(rb);
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            asm volatile("prefetch.global.L1 [%0];" ::"l"(k0 + i * C8));
            asm volatile("prefetch.global.L1 [%0];" ::"l"(k0 + M + i * C8));
          }
        }
      };
      // PF == 3: single register buffer, software pipelined by halves: the first row is requested before pass B,
      // and inside pass C the q = 0 (q = 1) half of the NEXT row is requested as soon as the MACs of that half are
      // done, so the loads fly under the next row's shared-memory reads and radix-8 butterflies
      auto load_keys_half = [&](double2 (&dst)[16], int rb, int q) {
        const double2 *__restrict__ k0 = key_row(rb) + q * M;
#pragma unroll
        for (int i = 0; i < 8; ++i) dst[8 * q + i] = ldg_key(k0 + i * C8);
      };
      if (PF == 1 || PF == 3) load_keys(kv[0], 0);                      // in flight across pass B
      if (PF == 2) prefetch_keys(0);
      // ------------------------------- pass B -------------------------------------------------
      constexpr int TASKS_B = ROWS_B * 128 / T;
      static_assert(TASKS_B * T == ROWS_B * 128, "pass B tasks must tile the CTA");
      double2 tb[R2];
      if constexpr (!TMEM_XB) load_tb(tb);
#ifdef MB200_ABL_NOPASSB
      if (a_i < 0)
#endif
      if constexpr (!TMEM_XB) {
#ifdef MB200_PB_PIPE
      constexpr bool PB_PIPE = (TASKS_B % 2 == 0);
#else
      constexpr bool PB_PIPE = false;
#endif
      if constexpr (PB_PIPE) {
        // software pipeline: the loads of task it+1 are issued before the butterflies of task it (tasks touch
        // disjoint blocks, so in-place is safe); pass B is shared-memory latency bound (ncu: 41-49 % short scoreboard)
        auto blk_of = [&](int it) { return buf + ((it * T) >> 7) * M + (((it * T) & 127) >> 3) * S + bB0; };
        auto ld = [&](double2 (&x)[R2], int it) {
          const double2 *blk = blk_of(it);
#pragma unroll
          for (int m = 0; m < R2; ++m) x[m] = blk[8 * m + qx[m]];
        };
        auto fin = [&](double2 (&x)[R2], int it) {
          double2 *blk = blk_of(it);
          reg_dif<R2>(x);
#pragma unroll
          for (int pos = 0; pos < R2; ++pos) {
            const int k = brev(pos, LOGR2);
            blk[8 * pos + qx[pos]] = k == 0 ? x[pos] : cmul(x[pos], tb[k]);
          }
        };
        double2 xa[R2], xb[R2];
        ld(xa, 0);
#pragma unroll 1
        for (int it = 0; it < TASKS_B; it += 2) {
          ld(xb, it + 1);
          fin(xa, it);
          if (it + 2 < TASKS_B) ld(xa, it + 2);
          fin(xb, it + 1);
        }
      } else {
#pragma unroll(PB_UNROLL)
      for (int it = 0; it < TASKS_B; ++it) {
        // task = tid + it*T: row = task >> 7, block b = (task & 127) >> 3 = (tid >> 3) + it*(T/8) (mod 16)
        double2 *blk = buf + ((it * T) >> 7) * M + (((it * T) & 127) >> 3) * S + bB0;
        double2 x[R2];
#pragma unroll
        for (int m = 0; m < R2; ++m) x[m] = blk[8 * m + qx[m]];
        reg_dif<R2>(x);
#pragma unroll
        for (int pos = 0; pos < R2; ++pos) {
          const int k = brev(pos, LOGR2);
          const double2 y = k == 0 ? x[pos] : cmul(x[pos], tb[k]);
          blk[8 * pos + qx[pos]] = y;
        }
      }
      }
      }   // !TMEM_XB
      __syncthreads();
      // ------------------------------- pass C + MAC ----------------------------------------------
      if (PF == 1) {
        // fully unrolled on purpose: a 2-row ping-pong loop (smaller code) measured 10 % slower
#pragma unroll
        for (int rb = 0; rb < ROWS_B; ++rb) {
          if (rb + 1 < ROWS_B) load_keys(kv[(rb + 1) & 1], rb + 1);     // next row's keys while this row computes
          const double2 *row = buf + rb * M;
          double2 x[8];
#pragma unroll
          for (int m = 0; m < 8; ++m) x[m] = row[8 * tid + (m ^ (tid & 7))];
          reg_dif<8>(x);
#pragma unroll
          for (int i = 0; i < 8; ++i) { cfma(fa[0][i], x[i], kv[rb & 1][i]); cfma(fa[1][i], x[i], kv[rb & 1][8 + i]); }
        }
      } else if (PF == 3) {
#pragma unroll(PC_UNROLL)
        for (int rb = 0; rb < ROWS_B; ++rb) {
          const double2 *row = buf + rb * M;
          double2 x[8];
#pragma unroll
          for (int m = 0; m < 8; ++m) x[m] = row[8 * tid + (m ^ (tid & 7))];
          reg_dif<8>(x);
#pragma unroll
          for (int i = 0; i < 8; ++i) cfma(fa[0][i], x[i], kv[0][i]);
          if (rb + 1 < ROWS_B) load_keys_half(kv[0], rb + 1, 0);
#pragma unroll
          for (int i = 0; i < 8; ++i) cfma(fa[1][i], x[i], kv[0][8 + i]);
          if (rb + 1 < ROWS_B) load_keys_half(kv[0], rb + 1, 1);
        }
      } else {
#pragma unroll(PC_UNROLL)
        for (int rb = 0; rb < ROWS_B; ++rb) {
          load_keys(kv[0], rb);
          if (PF == 2 && rb + 1 < ROWS_B) prefetch_keys(rb + 1);
          const double2 *row = buf + rb * M;
          double2 x[8];
#pragma unroll
          for (int m = 0; m < 8; ++m) x[m] = row[8 * tid + (m ^ (tid & 7))];
          reg_dif<8>(x);
#pragma unroll
          for (int i = 0; i < 8; ++i) { cfma(fa[0][i], x[i], kv[0][i]); cfma(fa[1][i], x[i], kv[0][8 + i]); }
        }
      }
      __syncthreads();
    };
    // The same batch with the number of levels as a run-time value: ONE copy of the pass A/B/C code for the
    // full and the ragged batch (the kernel is instruction-cache sensitive: ncu shows 64-72 % GCC instruction
    // requests and 12 % no_instruction stalls in pass A).  Used when the batches are ragged and the keys are
    // not double buffered in registers (that path needs compile-time buffer indices).
    auto batch_rt = [&](const int nb, const int lev0) {
      if (!PKALL) pack_digits(lev0 + nb);
      constexpr bool HOIST_TW = (LOGM <= 9);
      double2 twA[HOIST_TW ? 16 : 1];
      if (HOIST_TW) {
#pragma unroll
        for (int pos = 0; pos < 16; ++pos) twA[HOIST_TW ? pos : 0] = __ldg(&TA[brev(pos, 4) * S + qA]);
      }
#pragma unroll 1
      for (int lb = 0; lb < nb; ++lb) {
        const int sh = (PKALL ? (L - 1 - lev0 - lb) : (nb - 1 - lb)) * Bg_bit;
        double2 x[16];
#pragma unroll
        for (int m = 0; m < 16; ++m) {
          const double d0 = __hiloint2double(0x43300000, (int)((pk0[m] >> sh) & dmask)) - dbias;
          const double d1 = __hiloint2double(0x43300000, (int)((pk1[m] >> sh) & dmask)) - dbias;
          x[m] = mul_w64(make_double2(d0, d1), m, false);
        }
        reg_dif<16>(x);
        double2 *row = buf + (pA * nb + lb) * M;
#pragma unroll
        for (int pos = 0; pos < 16; ++pos) {
          const double2 t = HOIST_TW ? twA[HOIST_TW ? pos : 0] : __ldg(&TA[brev(pos, 4) * S + qA]);
          row[pos * S + qsw[pos & (NVA - 1)]] = cmul(x[pos], t);
        }
      }
      __syncthreads();
      const int tasks_b = 2 * nb * 128 / T;
#pragma unroll(PB_UNROLL)
      for (int it = 0; it < tasks_b; ++it) {
        double2 *blk = buf + ((it * T) >> 7) * M + (((it * T) & 127) >> 3) * S + bB0;
        double2 x[R2];
#pragma unroll
        for (int m = 0; m < R2; ++m) x[m] = blk[8 * m + qx[m]];
        reg_dif<R2>(x);
#pragma unroll
        for (int pos = 0; pos < R2; ++pos) {
          const int k = brev(pos, LOGR2);
          const double2 y = k == 0 ? x[pos] : cmul(x[pos], __ldg(&TB[k * 8 + qpB]));
          blk[8 * pos + qx[pos]] = y;
        }
      }
      __syncthreads();
#pragma unroll 1
      for (int p = 0; p < 2; ++p) {
#pragma unroll(PC_UNROLL)
        for (int lv = 0; lv < nb; ++lv) {
          const double2 *__restrict__ k0 = key + (size_t)((p * L + lev0 + lv) * 2) * M + tid;
          double2 kv[16];
#pragma unroll
          for (int i = 0; i < 8; ++i) { kv[i] = ldg_key(k0 + i * C8); kv[8 + i] = ldg_key(k0 + M + i * C8); }
          const double2 *row = buf + (p * nb + lv) * M;
          double2 x[8];
#pragma unroll
          for (int m = 0; m < 8; ++m) x[m] = row[8 * tid + (m ^ (tid & 7))];
          reg_dif<8>(x);
#pragma unroll
          for (int i = 0; i < 8; ++i) { cfma(fa[0][i], x[i], kv[i]); cfma(fa[1][i], x[i], kv[8 + i]); }
        }
      }
      __syncthreads();
    };
    // measured: no faster than the two unrolled copies (47.2-47.8 ms vs 46.5-48.0 ms at level 1) -> opt-in
#ifdef MB200_ROLLED_BATCH
    constexpr bool ROLLED = (LB < L) && PF == 0 && G == 1;
#else
    constexpr bool ROLLED = false;
#endif
    if constexpr (ROLLED) {
#pragma unroll 1
      for (int lev0 = 0; lev0 < L; lev0 += LB) batch_rt(min(LB, L - lev0), lev0);
    } else {
      // full batches of LB levels, then the ragged remainder (compile-time structure: constant shifts and rows)
#pragma unroll
      for (int lev0 = 0; lev0 + LB <= L; lev0 += LB) batch(std::integral_constant<int, LB>{}, lev0);
      if constexpr (L % LB != 0) batch(std::integral_constant<int, L % LB>{}, L - L % LB);
    }

    // ---------------------------------- inverse: C' ------------------------------------------------
#pragma unroll
    for (int pp = 0; pp < 2; ++pp) {
      reg_dit_inv<8>(fa[pp]);
      double2 *row = buf + pp * M;
#pragma unroll
      for (int m = 0; m < 8; ++m) row[8 * tid + (m ^ (tid & 7))] = fa[pp][m];
    }
    __syncthreads();
    // ---------------------------------- B' ---------------------------------------------------------
    constexpr int TASKS_BI = 2 * 128 / T > 0 ? 2 * 128 / T : 1;
    double2 tbi[R2];
    load_tb(tbi);
#ifdef MB200_ABL_NOPASSB
    if (a_i < 0)
#endif
#pragma unroll
    for (int it = 0; it < TASKS_BI; ++it) {
      double2 *blk = buf + ((it * T) >> 7) * M + (((it * T) & 127) >> 3) * S + bB0;
      double2 x[R2];
#pragma unroll
      for (int pos = 0; pos < R2; ++pos) {
        const int k = brev(pos, LOGR2);
        const double2 y = blk[8 * pos + qx[pos]];
        x[pos] = k == 0 ? y : cmul_conj(y, tbi[k]);
      }
      reg_dit_inv<R2>(x);
#pragma unroll
      for (int m = 0; m < R2; ++m) blk[8 * m + qx[m]] = x[m];
    }
    __syncthreads();
    // ---------------------------------- A' + accumulate --------------------------------------------
    {
      const double2 *row = buf + pA * M;
      double2 x[16];
      if (TMEM_TW) {
#pragma unroll
        for (int g4 = 0; g4 < 4; ++g4) {
          double2 tw[4];
          tmem_ld4(tw, tw_taddr + 16 * g4);
#pragma unroll
          for (int i = 0; i < 4; ++i) x[4 * g4 + i] = cmul_conj(row[(4 * g4 + i) * S + qsw[(4 * g4 + i) & (NVA - 1)]], tw[i]);
        }
      } else {
#pragma unroll
        for (int pos = 0; pos < 16; ++pos) {
          const double2 t = __ldg(&TA[brev(pos, 4) * S + qA]);
          x[pos] = cmul_conj(row[pos * S + qsw[pos & (NVA - 1)]], t);
        }
      }
      reg_dit_inv<16>(x);
      u64 *ap = acc + pA * N;
      const u64 *addp = (DIRECT && A.in1) ? A.in1 + (size_t)ct * 2 * N + pA * N : nullptr;
      u64 ownA[8];
#pragma unroll
      for (int m = 0; m < 16; ++m) {
        const double2 z = mul_w64(x[m], m, true);
        const int j = qA + m * S;
        if (DIRECT) {                                  // trlwe_from_DFT (+ in1 for a CMUX, vertical_packing.c:30)
          ap[j] = f64_to_torus_fast(z.x * inv_M) + (addp ? addp[j] : 0ull);
          ap[j + M] = f64_to_torus_fast(z.y * inv_M) + (addp ? addp[j + M] : 0ull);
        } else if (TMEM_ACC) {
          if ((m & 3) == 0) tmem_ld_u64x8(ownA, tw_taddr + COL_ACC + 4 * m);
          const u64 n0 = ownA[2 * (m & 3)] + f64_to_torus_fast(z.x * inv_M);
          const u64 n1 = ownA[2 * (m & 3) + 1] + f64_to_torus_fast(z.y * inv_M);
          ownA[2 * (m & 3)] = n0; ownA[2 * (m & 3) + 1] = n1;
          ap[j] = n0;                                  // shared copy: read (rotated) by other threads in the next step
          ap[j + M] = n1;
          if ((m & 3) == 3) tmem_st_u64x8(tw_taddr + COL_ACC + 4 * (m - 3), ownA);
        } else {
          ap[j] += f64_to_torus_fast(z.x * inv_M);     // trlwe_from_DFT + trlwe_addto
          ap[j + M] += f64_to_torus_fast(z.y * inv_M);
        }
      }
      if (TMEM_ACC) asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
    }
    __syncthreads();
  }

  if (TMEM_TW) {                                       // every thread is past its last tcgen05.ld (block barrier above)
    if (tid < 32) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base_s), "n"(TMEM_COLS) : "memory");
  }
  // ---- epilogue: sample extraction at index 0 (trlwe.c:540-552) or the raw accumulator -------------
  if (!live) return;
  if (A.extract) {
    u64 *o = A.out + (size_t)ct * (N + 1);
    for (int c = tid; c < N; c += T) o[c] = (c == 0) ? acc[0] : (0ull - acc[N - c]);
    if (tid == 0) o[N] = acc[N];
  } else {
    u64 *o = A.out + (size_t)ct * 2 * N;
    for (int c = tid; c < 2 * N; c += T) o[c] = acc[c];
  }
}


const double2 *k1_tables_for(int N);

}  // namespace mb
