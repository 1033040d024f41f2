// blind_rotate_k1h.cu -- k = 1 blind rotation with HALF the per-thread state of blind_rotate_k1.cu:
// T = M/4 threads per ciphertext (128 at N=1024), radix 8 x R2 x 8 (R2 = M/64), so that twice as
// many warps are resident per SM at <= 168 registers (ncu on the T = M/8 kernel: 1.5 warps per
// scheduler, issue slots 40 % busy, the rest fixed-latency / shared-memory / L2 waits).
//
//   pass A   thread (p, q), q < M/8: coefficients q + m*M/8 (+M), m < 8 -> digits -> radix-8 DIF
//   pass B   radix-R2 DIF on 8-strided groups inside blocks of M/8 positions, in place
//   pass C   thread (c, h): slot set c = positions 8c..8c+7, h = input polynomial whose rows it
//            consumes; radix-8 DIF + MAC against both output polynomials' key rows into a private
//            partial accumulator; the lane pair (c,0),(c,1) then swaps the halves it does not keep
//            (one shuffle per word) so that thread (c, h) ends up with output polynomial h
//   C'/B'/A' inverse, A' as in the T = M/8 kernel with 8 coefficient pairs per thread.
//
// Same arithmetic, tables, key layout, swizzle and prologue/epilogue as blind_rotate_k1.cu; see that
// file for the reference citations.
#include <map>
#include <mutex>
#include <type_traits>
#include <vector>

#include "k1_common.cuh"

namespace mb {

template <int LOGM, int L, int LB, int MINB>
__global__ void __launch_bounds__((1 << LOGM) / 4, MINB) blind_rotate_k1h_kernel(K1Args A) {
  constexpr int M = 1 << LOGM, N = 2 * M, S = M / 8, R2 = M / 64, T = M / 4, C8 = M / 8;
  constexpr int LOGR2 = clog2(R2);
  constexpr int ROWS = 2 * L, ROWS_B = 2 * LB;
  static_assert(R2 >= 2 && R2 <= 16, "supported N: 256..2048");
  static_assert(LB >= 1 && LB <= L, "levels per batch");

  extern __shared__ __align__(16) unsigned char smem_raw[];
  u64 *acc = reinterpret_cast<u64 *>(smem_raw);                       // [2][N]
  double2 *buf = reinterpret_cast<double2 *>(acc + 2 * N);           // [ROWS_B][M]
  unsigned short *rot = reinterpret_cast<unsigned short *>(buf + ROWS_B * M);

  const int tid = threadIdx.x, ct = blockIdx.x;
  const int log_N2 = LOGM + 2;
  const double2 *__restrict__ TA = A.tab;              // [8][S]
  const double2 *__restrict__ TB = A.tab + 8 * S;      // [R2][8]
  const u64 *in = A.in + (size_t)(ct / A.in_div) * A.in_stride;
  const u64 *tv = A.tv + (size_t)(A.tv_count > 1 ? ct % A.tv_count : 0) * 2 * N;
  const int Bg_bit = A.Bg_bit;

  int rot0 = 0;
  if (A.init_rotate) {
    u64 b = in[A.size];
    if (A.preprocess) b = pb_preprocess(b, A.kappa, A.theta, log_N2);
    rot0 = (2 * N - (int)torus2int(b + A.prec_offset, log_N2)) & (2 * N - 1);
  }
  for (int c = tid; c < 2 * N; c += T) {
    const int p = c / N, i = c - p * N;
    acc[c] = rot0 ? rotated_coeff(tv + (size_t)p * N, i, rot0, N) : tv[c];
  }
  for (int i = tid; i < A.size; i += T) {
    u64 av = in[i];
    if (A.preprocess) av = pb_preprocess(av, A.kappa, A.theta, log_N2);
    rot[i] = (unsigned short)(torus2int(av, log_N2) & (2 * N - 1));
  }
  __syncthreads();

  const u64 off = decomp_offset(Bg_bit, L);
  const unsigned dmask = (1u << Bg_bit) - 1u;
  const double dbias = 4503599627370496.0 + (double)(1 << (Bg_bit - 1));
  const double inv_M = 1.0 / (double)M;
  const int pA = tid / S, qA = tid - pA * S;           // pass A / A' ownership (T = 2S)
  const int qpB = tid & 7;
  const int cC = tid >> 1, hC = tid & 1;               // pass C: slot set and input-polynomial half
  // explicit swizzle: phys(s) = s ^ ((s >> 3) & 7)
  constexpr int S8 = S / 8, NVA = (S8 >= 8) ? 1 : 8 / S8;
  int qsw[NVA];
#pragma unroll
  for (int v = 0; v < NVA; ++v) qsw[v] = qA ^ ((v * S8 + (qA >> 3)) & 7);
  int qx[R2];
#pragma unroll
  for (int m = 0; m < R2; ++m) qx[m] = qpB ^ ((((tid >> 3) * R2) + m) & 7);
  const int bB0 = (tid >> 3) * S;                      // pass-B block of task `tid` (block size S = 8*R2)

  for (int step = 0; step < A.size; ++step) {
    const int a_i = rot[step];
    if (a_i == 0) continue;
    const double2 *__restrict__ key = A.bsk + (size_t)step * ROWS * 2 * M;

    double2 fa[2][8];
#pragma unroll
    for (int pp = 0; pp < 2; ++pp)
#pragma unroll
      for (int i = 0; i < 8; ++i) fa[pp][i] = make_double2(0.0, 0.0);

    auto batch = [&](auto nb_tag, const int lev0) {
      constexpr int NB = decltype(nb_tag)::value, RB = 2 * NB;
      // ------------------------------- pass A -------------------------------------------------
      {
        const u64 *ap = acc + pA * N;
        const int pk_shift = 64 - (lev0 + NB) * Bg_bit;
        const int base = (qA - a_i) & (2 * N - 1);
        unsigned pk0[8], pk1[8];
#pragma unroll
        for (int m = 0; m < 8; ++m) {
          const int j = qA + m * S;
          const int s0 = (base + m * S) & (2 * N - 1), s1 = (s0 + M) & (2 * N - 1);
          const u64 r0 = ap[s0 & (N - 1)], r1 = ap[s1 & (N - 1)];
          const u64 t0 = off - ap[j], t1 = off - ap[j + M];
          const u64 v0 = (s0 & N) ? t0 - r0 : t0 + r0;
          const u64 v1 = (s1 & N) ? t1 - r1 : t1 + r1;
          pk0[m] = (unsigned)(v0 >> pk_shift);
          pk1[m] = (unsigned)(v1 >> pk_shift);
        }
#pragma unroll
        for (int lb = 0; lb < NB; ++lb) {
          const int sh = (NB - 1 - lb) * Bg_bit;
          double2 x[8];
#pragma unroll
          for (int m = 0; m < 8; ++m) {
            const double d0 = __hiloint2double(0x43300000, (int)((pk0[m] >> sh) & dmask)) - dbias;
            const double d1 = __hiloint2double(0x43300000, (int)((pk1[m] >> sh) & dmask)) - dbias;
            x[m] = mul_w64(make_double2(d0, d1), 2 * m, false);       // w^(m*M/8) = W_32^m
          }
          reg_dif<8>(x);
          double2 *row = buf + (pA * NB + lb) * M;
#pragma unroll
          for (int pos = 0; pos < 8; ++pos) {
            const double2 t = __ldg(&TA[brev(pos, 3) * S + qA]);       // w^q * W_M^(q*k1)
            row[pos * S + qsw[pos & (NVA - 1)]] = cmul(x[pos], t);
          }
        }
      }
      __syncthreads();
      // key rows double buffered in registers (this is the small-batch kernel: few resident CTAs, nothing else hides
      // the L2 latency): row 0 of the batch is requested here and flies under pass B, row lb + 1 under the butterflies of lb
      double2 kv[2][16];
      auto load_keys = [&](double2 (&dst)[16], int lb) {
        const int r = hC * L + lev0 + lb;                               // TRGSW row (trgsw.c:394-419 order)
        const double2 *__restrict__ k0 = key + (size_t)(r * 2) * M + cC;
#pragma unroll
        for (int i = 0; i < 8; ++i) { dst[i] = ldg_key(k0 + i * C8); dst[8 + i] = ldg_key(k0 + M + i * C8); }
      };
      load_keys(kv[0], 0);
      // ------------------------------- pass B: RB*64 radix-R2 butterflies ------------------------
      constexpr int TASKS_B = (RB * 64 + T - 1) / T;
#pragma unroll
      for (int it = 0; it < TASKS_B; ++it) {
        // task = tid + it*T -> row = task >> 6, block = (task & 63) >> 3; rows are M = 8*S apart
        if ((RB * 64) % T != 0 && tid + it * T >= RB * 64) break;
        double2 *blk = buf + (size_t)(it * (T >> 3)) * S + bB0;
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
      // ------------------------------- pass C + MAC: rows of input polynomial hC -----------------
#pragma unroll
      for (int lb = 0; lb < NB; ++lb) {
        if (lb + 1 < NB) load_keys(kv[(lb + 1) & 1], lb + 1);
        const double2 *row = buf + (hC * NB + lb) * M + 8 * cC;
        double2 x[8];
#pragma unroll
        for (int m = 0; m < 8; ++m) x[m] = row[m ^ (cC & 7)];
        reg_dif<8>(x);
#pragma unroll
        for (int i = 0; i < 8; ++i) { cfma(fa[0][i], x[i], kv[lb & 1][i]); cfma(fa[1][i], x[i], kv[lb & 1][8 + i]); }
      }
      __syncthreads();
    };
#pragma unroll
    for (int lev0 = 0; lev0 + LB <= L; lev0 += LB) batch(std::integral_constant<int, LB>{}, lev0);
    if constexpr (L % LB != 0) batch(std::integral_constant<int, L % LB>{}, L - L % LB);

    // ---- lane pair exchange: thread (c, h) keeps output polynomial h = own partial + partner's ----
    double2 keep[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const double2 give = hC ? fa[0][i] : fa[1][i];
      const double gx = __shfl_xor_sync(0xffffffffu, give.x, 1), gy = __shfl_xor_sync(0xffffffffu, give.y, 1);
      const double2 mine = hC ? fa[1][i] : fa[0][i];
      keep[i] = make_double2(mine.x + gx, mine.y + gy);
    }
    // ---------------------------------- inverse: C' ------------------------------------------------
    reg_dit_inv<8>(keep);
    {
      double2 *row = buf + hC * M + 8 * cC;
#pragma unroll
      for (int m = 0; m < 8; ++m) row[m ^ (cC & 7)] = keep[m];
    }
    __syncthreads();
    // ---------------------------------- B': 2*64 butterflies -----------------------------------------
    if (tid < 128) {
      double2 *blk = buf + bB0;                         // task = tid: row = tid >> 6 (rows are 8*S apart)
      double2 x[R2];
#pragma unroll
      for (int pos = 0; pos < R2; ++pos) {
        const int k = brev(pos, LOGR2);
        const double2 y = blk[8 * pos + qx[pos]];
        x[pos] = k == 0 ? y : cmul_conj(y, __ldg(&TB[k * 8 + qpB]));
      }
      reg_dit_inv<R2>(x);
#pragma unroll
      for (int m = 0; m < R2; ++m) blk[8 * m + qx[m]] = x[m];
    }
    if (T < 128) {                                       // M = 256: 64 threads, second half of the 128 tasks
      double2 *blk = buf + (size_t)(T >> 3) * S + bB0;
      double2 x[R2];
#pragma unroll
      for (int pos = 0; pos < R2; ++pos) {
        const int k = brev(pos, LOGR2);
        const double2 y = blk[8 * pos + qx[pos]];
        x[pos] = k == 0 ? y : cmul_conj(y, __ldg(&TB[k * 8 + qpB]));
      }
      reg_dit_inv<R2>(x);
#pragma unroll
      for (int m = 0; m < R2; ++m) blk[8 * m + qx[m]] = x[m];
    }
    __syncthreads();
    // ---------------------------------- A' + accumulate --------------------------------------------
    {
      const double2 *row = buf + pA * M;
      double2 x[8];
#pragma unroll
      for (int pos = 0; pos < 8; ++pos) {
        const double2 t = __ldg(&TA[brev(pos, 3) * S + qA]);
        x[pos] = cmul_conj(row[pos * S + qsw[pos & (NVA - 1)]], t);
      }
      reg_dit_inv<8>(x);
      u64 *ap = acc + pA * N;
#pragma unroll
      for (int m = 0; m < 8; ++m) {
        const double2 z = mul_w64(x[m], 2 * m, true);
        const int j = qA + m * S;
        ap[j] += f64_to_torus_fast(z.x * inv_M);
        ap[j + M] += f64_to_torus_fast(z.y * inv_M);
      }
    }
    __syncthreads();
  }

  if (A.extract) {
    u64 *o = A.out + (size_t)ct * (N + 1);
    for (int c = tid; c < N; c += T) o[c] = (c == 0) ? acc[0] : (0ull - acc[N - c]);
    if (tid == 0) o[N] = acc[N];
  } else {
    u64 *o = A.out + (size_t)ct * 2 * N;
    for (int c = tid; c < 2 * N; c += T) o[c] = acc[c];
  }
}

// ---- tables: TA[8][S] then TB[R2][8], S = M/8 -------------------------------------------------------
static std::mutex g_k1h_mu;
static std::map<long long, double2 *> g_k1h_tab;     // (device, N) -> table

static const double2 *k1h_tables_for(int N) {
  ensure_init();
  std::lock_guard<std::mutex> lk(g_k1h_mu);
  auto it = g_k1h_tab.find(dev_key(N));
  if (it != g_k1h_tab.end()) return it->second;
  const int M = N / 2, S = M / 8, R2 = M / 64;
  std::vector<double2> h((size_t)8 * S + (size_t)R2 * 8);
  for (int k1 = 0; k1 < 8; ++k1)
    for (int q = 0; q < S; ++q) {
      const long double ang = M_PIl * (long double)((long long)q * (4 * k1 + 1)) / (long double)N;
      h[(size_t)k1 * S + q] = make_double2((double)cosl(ang), (double)sinl(ang));
    }
  for (int k = 0; k < R2; ++k)
    for (int qp = 0; qp < 8; ++qp) {
      const long double ang = 2.0L * M_PIl * (long double)(qp * k) / (long double)S;
      h[(size_t)8 * S + k * 8 + qp] = make_double2((double)cosl(ang), (double)sinl(ang));
    }
  double2 *d = nullptr;
  MB_CHECK(cudaMalloc(&d, sizeof(double2) * h.size()));
  MB_CHECK(cudaMemcpy(d, h.data(), sizeof(double2) * h.size(), cudaMemcpyHostToDevice));
  MB_CHECK(cudaDeviceSynchronize());
  g_k1h_tab[dev_key(N)] = d;
  return d;
}

template <int LOGM, int L, int LB, int MINB>
static void launch_h(const K1Args &a, int count, cudaStream_t st) {
  constexpr int M = 1 << LOGM;
  const size_t smem = (size_t)2 * 2 * M * 8 + (size_t)2 * LB * M * 16 + (((size_t)a.size * 2 + 15) & ~(size_t)15);
  static size_t configured_dev[MB_MAX_DEV] = {0};          // function attributes are per device
  size_t &configured = configured_dev[current_device()];
  if (smem > configured) {
    MB_REQUIRE(smem <= 227 * 1024, "k1h kernel: %zu B of shared memory needed", smem);
    MB_CHECK(cudaFuncSetAttribute(blind_rotate_k1h_kernel<LOGM, L, LB, MINB>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    configured = smem;
  }
  blind_rotate_k1h_kernel<LOGM, L, LB, MINB><<<count, M / 4, smem, st>>>(a);
  MB_CHECK(cudaGetLastError());
  count_launch();
}

// Levels per shared-memory batch: all of them when they fit the 32-bit packed-digit word, else 2, else 1
static int k1h_lb(int logm, int l, int Bg_bit) {
  int lb = l;
  if (logm == 10 && lb > 2) lb = 2;
  while (lb > 1 && lb * Bg_bit > 32) --lb;
  if (lb == 3 && l == 4) lb = 2;
  return lb;
}

bool k1h_supported(const Params &p) {
  if (p.k != 1) return false;
  const int logm = ilog2i(p.N) - 1;
  if (!(logm >= 8 && logm <= 10 && (1 << (logm + 1)) == p.N && p.l >= 1 && p.l <= 4)) return false;
  return p.Bg_bit >= 1 && p.Bg_bit <= 31;   // see k1_supported
}

void k1h_variant_name(const Params &p, char *dst, size_t cap) {
  snprintf(dst, cap, "k1h<N=%d,l=%d,lb=%d,T=%d>", p.N, p.l, k1h_lb(ilog2i(p.N) - 1, p.l, p.Bg_bit), p.N / 8);
}

void launch_blind_rotate_k1h(const BlindRotateLaunch &b, cudaStream_t st) {
  MB_REQUIRE(b.b_index == 0 || b.b_index == b.size, "segmented launches exist for the k1q kernel only");
  const Params &p = b.bsk->p;
  MB_REQUIRE(k1h_supported(p) && !b.direct, "k1h kernel: unsupported parameters");
  upload_w64();
  K1Args a;
  a.bsk = b.bsk->d; a.tab = k1h_tables_for(p.N); a.tv = b.tv; a.tv_count = b.tv_count; a.in = b.in;
  a.in_stride = b.in_stride; a.in_div = b.in_div > 0 ? b.in_div : 1; a.size = b.size; a.out = b.out; a.extract = b.extract; a.init_rotate = b.init_rotate;
  a.prec_offset = b.prec_offset; a.preprocess = b.preprocess; a.kappa = b.kappa; a.theta = b.theta; a.Bg_bit = p.Bg_bit;
  a.count = b.count;
  const int logm = ilog2i(p.N) - 1;
  const int lb = k1h_lb(logm, p.l, p.Bg_bit);
  // one register-rich variant per shape: this kernel is the small-batch / low-latency path
  // (profiles/r1j_latency.log: 2.7 ms vs 4.4 ms per PBS at batch <= 148 for N = 1024)
#define MB_K1H_CASE(LM, LL, LBB, MB_) if (logm == LM && p.l == LL && lb == LBB) { launch_h<LM, LL, LBB, MB_>(a, b.count, st); return; }
  MB_K1H_CASE(8, 1, 1, 2) MB_K1H_CASE(8, 2, 2, 2) MB_K1H_CASE(8, 3, 3, 2) MB_K1H_CASE(8, 4, 4, 2)
  MB_K1H_CASE(8, 2, 1, 2) MB_K1H_CASE(8, 3, 2, 2) MB_K1H_CASE(8, 3, 1, 2) MB_K1H_CASE(8, 4, 2, 2) MB_K1H_CASE(8, 4, 1, 2)
  MB_K1H_CASE(9, 1, 1, 2) MB_K1H_CASE(9, 2, 2, 2) MB_K1H_CASE(9, 3, 3, 2) MB_K1H_CASE(9, 4, 4, 2)
  MB_K1H_CASE(9, 2, 1, 2) MB_K1H_CASE(9, 3, 2, 2) MB_K1H_CASE(9, 3, 1, 2) MB_K1H_CASE(9, 4, 2, 2) MB_K1H_CASE(9, 4, 1, 2)
  MB_K1H_CASE(10, 1, 1, 1) MB_K1H_CASE(10, 2, 2, 1) MB_K1H_CASE(10, 3, 2, 1) MB_K1H_CASE(10, 4, 2, 1)
  MB_K1H_CASE(10, 2, 1, 1) MB_K1H_CASE(10, 3, 1, 1) MB_K1H_CASE(10, 4, 1, 1)
#undef MB_K1H_CASE
  MB_FATAL("k1h kernel: no instantiation for N=%d l=%d lb=%d", p.N, p.l, lb);
}

}  // namespace mb
