// blind_rotate_k1.cu -- the specialised k = 1 blind-rotation kernel (the hot one).
//
// One CTA of T = M/8 threads per ciphertext (M = N/2 complex points).  The accumulator lives in
// shared memory for all n steps; every N/2-point negacyclic transform is three register-resident
// passes (radix 16 x R2 x 8, R2 = M/128) separated by two shared-memory exchanges:
//
//   pass A  thread (p, q): reads acc_p[q + m*M/16] (and +M, and the X^a-rotated copies),
//           forms (X^a - 1)*acc, extracts the signed gadget digits of all levels from the same
//           registers, folds/twists, radix-16 DIF in registers, twiddle, -> smem
//   pass B  radix-R2 DIF on 8-strided groups, twiddle, in place in smem
//   pass C  thread c owns FFT positions 8c..8c+7: radix-8 DIF in registers and the Fourier-domain
//           multiply-accumulate against the key rows straight from registers; the two output
//           polynomials' accumulators (2 x 8 complex) never leave registers until the inverse
//   inverse C' -> B' -> A' mirrors the above (DIT, conjugate twiddles); A' untwists, scales by 2/N,
//           reduces mod 2^64 on the integer pipe and adds into the accumulator words the thread owns.
//
// Key rows are read with coalesced 128-bit loads: the resident layout stores position 8c+m at
// index m*(M/8)+c (keys.cu), so lane c of a warp reads consecutive 16-byte words.
// Shared-memory exchanges use 16-byte complex elements with the XOR swizzle s ^ ((s>>3)&7), which
// makes pass A writes, pass B strided accesses and pass C's 128-byte-per-thread reads all
// bank-conflict free.
//
// Reference functions fused here: bootstrap.c:107-122 (blind_rotate loop), polynomial.c:220-235,
// polynomial.c:74-89, polynomial.c:359-375 (+ src/fft), trlwe.c:491-505, trlwe.c:629-634,
// trlwe.c:437, and the prologue/epilogue bootstrap.c:192-206 + trlwe.c:540-552.
#include <map>
#include <mutex>
#include <vector>

#include "common.cuh"
#include "device_math.cuh"
#include "w64_constants.cuh"

namespace mb {

// x * W_64^idx (or its conjugate).  idx is a compile-time constant after unrolling, so the trivial
// cases cost nothing and the 8th roots cost 2 mul + 2 add.
__device__ __forceinline__ double2 mul_w64(double2 x, int idx, bool conj) {
  idx &= 63;
  if (conj) idx = (64 - idx) & 63;
  const double h = 0.70710678118654752440;
  if (idx == 0) return x;
  if (idx == 16) return make_double2(-x.y, x.x);
  if (idx == 32) return make_double2(-x.x, -x.y);
  if (idx == 48) return make_double2(x.y, -x.x);
  if (idx == 8) return make_double2((x.x - x.y) * h, (x.x + x.y) * h);
  if (idx == 24) return make_double2((-x.x - x.y) * h, (x.x - x.y) * h);
  if (idx == 40) return make_double2((x.y - x.x) * h, (-x.x - x.y) * h);
  if (idx == 56) return make_double2((x.x + x.y) * h, (x.y - x.x) * h);
  const double c = W64C[idx], s = W64S[idx];
  return make_double2(fma(x.x, c, -x.y * s), fma(x.x, s, x.y * c));
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

__device__ __forceinline__ int swz(int s) { return s ^ ((s >> 3) & 7); }

struct K1Args {
  const double2 *bsk;
  const double2 *tab;     // TA[16][S] then TB[R2][8]
  const u64 *tv;
  int tv_count;
  const u64 *in;
  int in_stride;
  int size;
  u64 *out;
  int extract, init_rotate;
  u64 prec_offset;
  int preprocess, kappa, theta;
  int Bg_bit;
};

template <int LOGM, int L, int LB>
__global__ void __launch_bounds__((1 << LOGM) / 8) blind_rotate_k1_kernel(K1Args A) {
  constexpr int M = 1 << LOGM, N = 2 * M, S = M / 16, R2 = M / 128, T = M / 8, C8 = M / 8;
  constexpr int LOGR2 = clog2(R2);
  constexpr int ROWS = 2 * L, ROWS_B = 2 * LB;
  static_assert(L % LB == 0, "levels per batch must divide l");
  static_assert(R2 >= 2 && R2 <= 16, "supported N: 512..4096");

  extern __shared__ __align__(16) unsigned char smem_raw[];
  u64 *acc = reinterpret_cast<u64 *>(smem_raw);                       // [2][N]
  double2 *buf = reinterpret_cast<double2 *>(acc + 2 * N);           // [ROWS_B][M]

  const int tid = threadIdx.x, ct = blockIdx.x;
  const int log_N2 = LOGM + 2;
  const double2 *__restrict__ TA = A.tab;
  const double2 *__restrict__ TB = A.tab + 16 * S;
  const u64 *in = A.in + (size_t)ct * A.in_stride;
  const u64 *tv = A.tv + (size_t)(A.tv_count > 1 ? ct : 0) * 2 * N;
  const int Bg_bit = A.Bg_bit;

  // ---- initial accumulator: tv * X^(2N - round((b + 1/(4*torus_base)) * 2N))  (bootstrap.c:194-195)
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
  __syncthreads();

  const u64 off = decomp_offset(Bg_bit, L);
  const u64 dmask = (1ull << Bg_bit) - 1ull;
  const int half_bg = 1 << (Bg_bit - 1);
  const double inv_M = 1.0 / (double)M;
  const int pA = tid / S, qA = tid - pA * S;          // pass A / A' ownership
  const int qpB = tid & 7;                            // pass B twiddle column (T is a multiple of 8)

  for (int step = 0; step < A.size; ++step) {
    u64 av = in[step];
    if (A.preprocess) av = pb_preprocess(av, A.kappa, A.theta, log_N2);
    const int a_i = (int)torus2int(av, log_N2) & (2 * N - 1);
    if (a_i == 0) continue;                           // bootstrap.c:114
    const double2 *__restrict__ key = A.bsk + (size_t)step * ROWS * 2 * M;

    double2 fa[2][8];                                 // Fourier accumulators: positions 8*tid .. 8*tid+7
#pragma unroll
    for (int pp = 0; pp < 2; ++pp)
#pragma unroll
      for (int i = 0; i < 8; ++i) fa[pp][i] = make_double2(0.0, 0.0);

#pragma unroll
    for (int lev0 = 0; lev0 < L; lev0 += LB) {
      // ------------------------------- pass A -------------------------------------------------
      {
        const u64 *ap = acc + pA * N;
        u64 v0[16], v1[16];
#pragma unroll
        for (int m = 0; m < 16; ++m) {
          const int j = qA + m * S;
          v0[m] = rotated_coeff(ap, j, a_i, N) - ap[j] + off;          // (X^a - 1) * acc, + rounding offset
          v1[m] = rotated_coeff(ap, j + M, a_i, N) - ap[j + M] + off;
        }
#pragma unroll
        for (int lb = 0; lb < LB; ++lb) {
          const int sh = 64 - (lev0 + lb + 1) * Bg_bit;
          double2 x[16];
#pragma unroll
          for (int m = 0; m < 16; ++m) {
            const int d0 = (int)((v0[m] >> sh) & dmask) - half_bg;
            const int d1 = (int)((v1[m] >> sh) & dmask) - half_bg;
            // fold z = d0 + i*d1 and the constant part of the twist, w^(m*M/16) = W_64^m
            x[m] = mul_w64(make_double2((double)d0, (double)d1), m, false);
          }
          reg_dif<16>(x);
          double2 *row = buf + (pA * LB + lb) * M;
#pragma unroll
          for (int pos = 0; pos < 16; ++pos) {
            const double2 t = __ldg(&TA[brev(pos, 4) * S + qA]);       // w^q * W_M^(q*k1)
            row[swz(pos * S + qA)] = cmul(x[pos], t);
          }
        }
      }
      __syncthreads();
      // ------------------------------- pass B -------------------------------------------------
#pragma unroll 1
      for (int task = tid; task < ROWS_B * 128; task += T) {
        double2 *row = buf + (task >> 7) * M;
        const int t = task & 127, b = t >> 3;
        double2 x[R2];
#pragma unroll
        for (int m = 0; m < R2; ++m) x[m] = row[swz(b * S + qpB + 8 * m)];
        reg_dif<R2>(x);
#pragma unroll
        for (int pos = 0; pos < R2; ++pos) {
          const int k = brev(pos, LOGR2);
          const double2 y = k == 0 ? x[pos] : cmul(x[pos], __ldg(&TB[k * 8 + qpB]));
          row[swz(b * S + pos * 8 + qpB)] = y;
        }
      }
      __syncthreads();
      // ------------------------------- pass C + MAC ----------------------------------------------
#pragma unroll
      for (int rb = 0; rb < ROWS_B; ++rb) {
        const int p = rb / LB, lev = lev0 + (rb - p * LB);
        const int r = p * L + lev;                                      // TRGSW row (trgsw.c:394-419 order)
        const double2 *__restrict__ k0 = key + (size_t)(r * 2 + 0) * M + tid;
        const double2 *__restrict__ k1 = key + (size_t)(r * 2 + 1) * M + tid;
        double2 kv0[8], kv1[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) { kv0[i] = __ldg(k0 + i * C8); kv1[i] = __ldg(k1 + i * C8); }
        const double2 *row = buf + rb * M;
        double2 x[8];
#pragma unroll
        for (int m = 0; m < 8; ++m) x[m] = row[8 * tid + (m ^ (tid & 7))];
        reg_dif<8>(x);
#pragma unroll
        for (int i = 0; i < 8; ++i) { cfma(fa[0][i], x[i], kv0[i]); cfma(fa[1][i], x[i], kv1[i]); }
      }
      __syncthreads();
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
#pragma unroll 1
    for (int task = tid; task < 2 * 128; task += T) {
      double2 *row = buf + (task >> 7) * M;
      const int t = task & 127, b = t >> 3;
      double2 x[R2];
#pragma unroll
      for (int pos = 0; pos < R2; ++pos) {
        const int k = brev(pos, LOGR2);
        const double2 y = row[swz(b * S + pos * 8 + qpB)];
        x[pos] = k == 0 ? y : cmul_conj(y, __ldg(&TB[k * 8 + qpB]));
      }
      reg_dit_inv<R2>(x);
#pragma unroll
      for (int m = 0; m < R2; ++m) row[swz(b * S + qpB + 8 * m)] = x[m];
    }
    __syncthreads();
    // ---------------------------------- A' + accumulate --------------------------------------------
    {
      const double2 *row = buf + pA * M;
      double2 x[16];
#pragma unroll
      for (int pos = 0; pos < 16; ++pos) {
        const double2 t = __ldg(&TA[brev(pos, 4) * S + qA]);
        x[pos] = cmul_conj(row[swz(pos * S + qA)], t);
      }
      reg_dit_inv<16>(x);
      u64 *ap = acc + pA * N;
#pragma unroll
      for (int m = 0; m < 16; ++m) {
        const double2 z = mul_w64(x[m], m, true);
        const int j = qA + m * S;
        ap[j] += f64_to_torus(z.x * inv_M);            // trlwe_from_DFT + trlwe_addto
        ap[j + M] += f64_to_torus(z.y * inv_M);
      }
    }
    __syncthreads();
  }

  // ---- epilogue: sample extraction at index 0 (trlwe.c:540-552) or the raw accumulator -------------
  if (A.extract) {
    u64 *o = A.out + (size_t)ct * (N + 1);
    for (int c = tid; c < N; c += T) o[c] = (c == 0) ? acc[0] : (0ull - acc[N - c]);
    if (tid == 0) o[N] = acc[N];
  } else {
    u64 *o = A.out + (size_t)ct * 2 * N;
    for (int c = tid; c < 2 * N; c += T) o[c] = acc[c];
  }
}

// ---- per-N tables ------------------------------------------------------------------------------
static std::mutex g_k1_mu;
static std::map<int, double2 *> g_k1_tab;

const double2 *k1_tables_for(int N) {
  ensure_init();
  std::lock_guard<std::mutex> lk(g_k1_mu);
  auto it = g_k1_tab.find(N);
  if (it != g_k1_tab.end()) return it->second;
  const int M = N / 2, S = M / 16, R2 = M / 128;
  std::vector<double2> h((size_t)16 * S + (size_t)R2 * 8);
  for (int k1 = 0; k1 < 16; ++k1)
    for (int q = 0; q < S; ++q) {
      // w^q * W_M^(q*k1) = exp(i*pi*q*(4*k1+1)/N)
      const long double ang = M_PIl * (long double)((long long)q * (4 * k1 + 1)) / (long double)N;
      h[(size_t)k1 * S + q] = make_double2((double)cosl(ang), (double)sinl(ang));
    }
  for (int k = 0; k < R2; ++k)
    for (int qp = 0; qp < 8; ++qp) {
      // W_S^(qp*k) = exp(2*pi*i*qp*k/S)
      const long double ang = 2.0L * M_PIl * (long double)(qp * k) / (long double)S;
      h[(size_t)16 * S + k * 8 + qp] = make_double2((double)cosl(ang), (double)sinl(ang));
    }
  double2 *d = nullptr;
  MB_CHECK(cudaMalloc(&d, sizeof(double2) * h.size()));
  MB_CHECK(cudaMemcpy(d, h.data(), sizeof(double2) * h.size(), cudaMemcpyHostToDevice));
  g_k1_tab[N] = d;
  return d;
}

// ---- dispatch --------------------------------------------------------------------------------------
static int levels_per_batch(int logm, int l) {
  if (logm == 11) return 1;
  if (logm == 10 && l == 4) return 2;
  return l;
}

bool k1_supported(const Params &p) {
  if (p.k != 1) return false;
  const int logm = ilog2i(p.N) - 1;
  return logm >= 8 && logm <= 11 && p.l >= 1 && p.l <= 4 && (1 << (logm + 1)) == p.N;
}

static char g_name[64];
const char *k1_variant_name(const Params &p) {
  const int logm = ilog2i(p.N) - 1;
  snprintf(g_name, sizeof(g_name), "k1<N=%d,l=%d,lb=%d>", p.N, p.l, levels_per_batch(logm, p.l));
  return g_name;
}

template <int LOGM, int L, int LB>
static void launch_one(const K1Args &a, int count, cudaStream_t st) {
  constexpr int M = 1 << LOGM;
  constexpr size_t smem = (size_t)2 * 2 * M * 8 + (size_t)2 * LB * M * 16;
  static bool configured = false;
  if (!configured) {
    MB_CHECK(cudaFuncSetAttribute(blind_rotate_k1_kernel<LOGM, L, LB>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    configured = true;
  }
  blind_rotate_k1_kernel<LOGM, L, LB><<<count, M / 8, smem, st>>>(a);
  MB_CHECK(cudaGetLastError());
  count_launch();
}

void launch_blind_rotate_k1(const BlindRotateLaunch &b, cudaStream_t st) {
  const Params &p = b.bsk->p;
  MB_REQUIRE(k1_supported(p) && !b.direct, "k1 kernel: unsupported parameters");
  K1Args a;
  a.bsk = b.bsk->d; a.tab = k1_tables_for(p.N); a.tv = b.tv; a.tv_count = b.tv_count; a.in = b.in;
  a.in_stride = b.in_stride; a.size = b.size; a.out = b.out; a.extract = b.extract; a.init_rotate = b.init_rotate;
  a.prec_offset = b.prec_offset; a.preprocess = b.preprocess; a.kappa = b.kappa; a.theta = b.theta; a.Bg_bit = p.Bg_bit;
  const int logm = ilog2i(p.N) - 1;
#define MB_K1_CASE(LM, LL, LBB) if (logm == LM && p.l == LL) { launch_one<LM, LL, LBB>(a, b.count, st); return; }
  MB_K1_CASE(8, 1, 1) MB_K1_CASE(8, 2, 2) MB_K1_CASE(8, 3, 3) MB_K1_CASE(8, 4, 4)
  MB_K1_CASE(9, 1, 1) MB_K1_CASE(9, 2, 2) MB_K1_CASE(9, 3, 3) MB_K1_CASE(9, 4, 4)
  MB_K1_CASE(10, 1, 1) MB_K1_CASE(10, 2, 2) MB_K1_CASE(10, 3, 3) MB_K1_CASE(10, 4, 2)
  MB_K1_CASE(11, 1, 1) MB_K1_CASE(11, 2, 1) MB_K1_CASE(11, 3, 1) MB_K1_CASE(11, 4, 1)
#undef MB_K1_CASE
  MB_FATAL("k1 kernel: no instantiation for N=%d l=%d", p.N, p.l);
}

}  // namespace mb
