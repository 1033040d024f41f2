// blind_rotate_k1c.cu -- the LATENCY kernel: one ciphertext per thread-block CLUSTER of two CTAs (two SMs).
//
// The blind rotation of one ciphertext is 632 strictly sequential steps, so a single bootstrap cannot use more
// than the SMs that work on one step.  Here CTA p of the pair owns accumulator polynomial p (0 = a, 1 = b):
//   * it decomposes ITS polynomial ((X^a - 1)*acc_p, all l gadget levels), runs the l forward transforms and
//     multiplies them with key rows p*l .. p*l + l - 1 into partial Fourier sums for BOTH output polynomials;
//   * the partial sum that belongs to the other polynomial is written straight into the peer CTA's shared
//     memory (distributed shared memory, st.shared::cluster), one cluster barrier per step;
//   * it adds the peer's contribution, inverts ONE polynomial and accumulates into acc_p.
// Per step each SM does l forward + 1 inverse transform instead of 2l + 2, at the price of 8*M bytes over the
// SM-to-SM network and one cluster barrier.  Same arithmetic per transform as blind_rotate_k1_kernel (passes
// radix 16 x R2 x 8, same tables, same rounding), so results agree with it to f64 reassociation of the two
// partial sums.  Used when the batch is at most half the SM count (api.cu: run_blind_rotate).
//
// Reference functions fused: as blind_rotate_k1.cu (bootstrap.c:107-122, 192-206; trgsw.c:385-423; ...).
#include <map>
#include <mutex>

#include "k1_common.cuh"

namespace mb {

const double2 *k1_tables_for(int N);
bool k1_supported(const Params &p);

__device__ __forceinline__ void cluster_barrier() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ void st_cluster(unsigned addr, double2 v) {
  asm volatile("st.shared::cluster.v2.f64 [%0], {%1, %2};" ::"r"(addr), "d"(v.x), "d"(v.y) : "memory");
}

template <int LOGM, int L>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__((1 << LOGM) / 8, 1) blind_rotate_k1c_kernel(K1Args A) {
  constexpr int M = 1 << LOGM, N = 2 * M, S = M / 16, R2 = M / 128, T = M / 8, C8 = M / 8;
  constexpr int LOGR2 = clog2(R2);
  constexpr int ROWS = 2 * L;
  static_assert(R2 >= 2 && R2 <= 16 && L >= 1 && L <= 4, "supported: N = 512..4096, l = 1..4");

  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int tid = threadIdx.x;
  unsigned rank;
  asm("mov.u32 %0, %%cluster_ctarank;" : "=r"(rank));
  const int p_own = (int)rank;                       // polynomial owned by this CTA: 0 = a, 1 = b
  const int ct = blockIdx.x >> 1;

  u64 *acc = reinterpret_cast<u64 *>(smem_raw);                       // [N]   the owned polynomial
  double2 *buf = reinterpret_cast<double2 *>(acc + N);               // [L][M] forward rows; row 0 reused by the inverse
  double2 *xbuf = buf + L * M;                                        // [2][M] partial sums received from the peer
  unsigned short *rot = reinterpret_cast<unsigned short *>(xbuf + 2 * M);

  const int log_N2 = LOGM + 2;
  const double2 *__restrict__ TA = A.tab;
  const double2 *__restrict__ TB = A.tab + 16 * S;
  const u64 *in = A.in + (size_t)(ct / A.in_div) * A.in_stride;
  const u64 *tv = A.tv + (size_t)(A.tv_count > 1 ? ct % A.tv_count : 0) * 2 * N + (size_t)p_own * N;
  const int Bg_bit = A.Bg_bit;

  // The cluster barrier of every step invalidates L1 (it is a cluster-scope acquire), so twiddles re-read from global
  // memory miss L1 every step (ncu: long scoreboard in pass A / B): they live in the thread's tensor-memory lane instead
  // (columns 0..63: pass A / A', 64..64+4*R2: pass B / B').
  __shared__ unsigned tmem_base_s;
  if (tid < 32) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 128;\n\t"
                 "tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;"
                 ::"r"((unsigned)__cvta_generic_to_shared(&tmem_base_s)) : "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const unsigned tw_taddr = tmem_base_s + ((unsigned)(tid >> 5) << 21);
  {
    const int qA0 = tid % S;
#pragma unroll
    for (int g4 = 0; g4 < 4; ++g4) {
      double2 tw[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) tw[i] = __ldg(&TA[brev(4 * g4 + i, 4) * S + qA0]);
      tmem_st4(tw_taddr + 16 * g4, tw);
    }
#pragma unroll
    for (int g4 = 0; g4 < (R2 + 3) / 4; ++g4) {
      double2 tw[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) tw[i] = __ldg(&TB[((4 * g4 + i) & (R2 - 1)) * 8 + (tid & 7)]);
      tmem_st4(tw_taddr + 64 + 16 * g4, tw);
    }
    asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
  }
  auto load_tb = [&](double2 (&tb)[R2 < 4 ? 4 : R2]) {
#pragma unroll
    for (int g4 = 0; g4 < (R2 + 3) / 4; ++g4) {
      double2 tw[4];
      tmem_ld4(tw, tw_taddr + 64 + 16 * g4);
#pragma unroll
      for (int i = 0; i < 4; ++i) tb[4 * g4 + i] = tw[i];
    }
  };

  int rot0 = 0;
  if (A.init_rotate) {
    u64 b = in[A.size];
    if (A.preprocess) b = pb_preprocess(b, A.kappa, A.theta, log_N2);
    rot0 = (2 * N - (int)torus2int(b + A.prec_offset, log_N2)) & (2 * N - 1);
  }
  for (int c = tid; c < N; c += T) acc[c] = rot0 ? rotated_coeff(tv, c, rot0, N) : tv[c];
  for (int i = tid; i < A.size; i += T) {
    u64 av = in[i];
    if (A.preprocess) av = pb_preprocess(av, A.kappa, A.theta, log_N2);
    rot[i] = (unsigned short)(torus2int(av, log_N2) & (2 * N - 1));
  }
  // the peer's receive buffer in the cluster shared-memory window
  const unsigned xbuf_local = (unsigned)__cvta_generic_to_shared(xbuf);
  unsigned xbuf_peer;
  asm("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(xbuf_peer) : "r"(xbuf_local), "r"(rank ^ 1u));
  cluster_barrier();                                  // both CTAs resident and initialised before any remote store

  const u64 off = decomp_offset(Bg_bit, L);
  const unsigned dmask = (1u << Bg_bit) - 1u;
  const double dbias = 4503599627370496.0 + (double)(1 << (Bg_bit - 1));
  const double inv_M = 1.0 / (double)M;
  const int pA = tid / S, qA = tid - pA * S;          // pass A: half pA of the CTA transforms levels 2*pA, 2*pA + 1
  const int qpB = tid & 7;
  constexpr int S8 = S / 8, NVA = (S8 >= 8) ? 1 : 8 / S8;
  int qsw[NVA];
#pragma unroll
  for (int v = 0; v < NVA; ++v) qsw[v] = qA ^ ((v * S8 + (qA >> 3)) & 7);
  int qx[R2];
#pragma unroll
  for (int m = 0; m < R2; ++m) qx[m] = qpB ^ ((((tid >> 3) * S8) + m) & 7);
  const int bB0 = (tid >> 3) * S;
  const int n_lev = max(0, min(2, L - 2 * pA));       // levels this thread transforms in pass A
  const int pk_shift = 64 - (2 * pA + n_lev) * Bg_bit;
  int parity = 0;

  for (int step = 0; step < A.size; ++step) {
    const int a_i = rot[step];
    if (a_i == 0) continue;                           // bootstrap.c:114 (both CTAs see the same mask)
    const double2 *__restrict__ key = A.bsk + (size_t)step * ROWS * 2 * M + (size_t)(p_own * L) * 2 * M;
    // a single bootstrap reads every key row exactly once (L2 hit rate 4 % at batch 1): pull the NEXT step's rows of
    // this CTA (l rows x 2 polynomials x M x 16 B) into L2 now, one 128-byte line per thread and iteration
    if (step + 1 < A.size) {
      const char *nxt = reinterpret_cast<const char *>(key + (size_t)ROWS * 2 * M);
      for (int off = tid * 128; off < L * 2 * M * 16; off += T * 128) asm volatile("prefetch.global.L2 [%0];" ::"l"(nxt + off));
    }

    double2 f0[8], f1[8];                             // partial Fourier sums for output polynomials 0 and 1
#pragma unroll
    for (int i = 0; i < 8; ++i) { f0[i] = make_double2(0.0, 0.0); f1[i] = make_double2(0.0, 0.0); }

    // ------------------------------- pass A: digits of (X^a - 1)*acc_p, fold/twist, radix 16 ---------------
    if (n_lev > 0) {
      unsigned pk0[16], pk1[16];
      const int base = (qA - a_i) & (2 * N - 1);
#pragma unroll
      for (int m = 0; m < 16; ++m) {
        const int j = qA + m * S;
        const int s0 = (base + m * S) & (2 * N - 1), s1 = (s0 + M) & (2 * N - 1);
        const u64 r0 = acc[s0 & (N - 1)], r1 = acc[s1 & (N - 1)];
        const u64 t0 = off - acc[j], t1 = off - acc[j + M];
        const u64 v0 = (s0 & N) ? t0 - r0 : t0 + r0;
        const u64 v1 = (s1 & N) ? t1 - r1 : t1 + r1;
        pk0[m] = (unsigned)(v0 >> pk_shift);
        pk1[m] = (unsigned)(v1 >> pk_shift);
      }
#pragma unroll 1
      for (int it = 0; it < n_lev; ++it) {
        const int sh = (n_lev - 1 - it) * Bg_bit;
        double2 x[16];
#pragma unroll
        for (int m = 0; m < 16; ++m) {
          const double d0 = __hiloint2double(0x43300000, (int)((pk0[m] >> sh) & dmask)) - dbias;
          const double d1 = __hiloint2double(0x43300000, (int)((pk1[m] >> sh) & dmask)) - dbias;
          x[m] = mul_w64(make_double2(d0, d1), m, false);
        }
        reg_dif<16>(x);
        double2 *row = buf + (2 * pA + it) * M;
#pragma unroll
        for (int g4 = 0; g4 < 4; ++g4) {
          double2 tw[4];
          tmem_ld4(tw, tw_taddr + 16 * g4);
#pragma unroll
          for (int i = 0; i < 4; ++i) row[(4 * g4 + i) * S + qsw[(4 * g4 + i) & (NVA - 1)]] = cmul(x[4 * g4 + i], tw[i]);
        }
      }
    }
    __syncthreads();
    // key rows: double buffered in registers -- row 0 is requested here and flies under pass B, row rb + 1 under the
    // butterflies of row rb (with one CTA per SM nothing else hides the L2 latency; the throughput kernel does not need this)
    double2 kv[2][16];
    auto load_keys = [&](double2 (&dst)[16], int rb) {
      const double2 *__restrict__ k0 = key + (size_t)rb * 2 * M + tid;
#pragma unroll
      for (int i = 0; i < 8; ++i) { dst[i] = ldg_key(k0 + i * C8); dst[8 + i] = ldg_key(k0 + M + i * C8); }
    };
    load_keys(kv[0], 0);
    // ------------------------------- pass B: radix R2 in shared memory ---------------------------------------
    constexpr int TASKS_B = L * 128 / T > 0 ? L * 128 / T : 1;
    double2 tb[R2 < 4 ? 4 : R2];
    load_tb(tb);
#pragma unroll 2
    for (int it = 0; it < TASKS_B; ++it) {
      double2 *blk = buf + ((it * T) >> 7) * M + (((it * T) & 127) >> 3) * S + bB0;
      double2 x[R2];
#pragma unroll
      for (int m = 0; m < R2; ++m) x[m] = blk[8 * m + qx[m]];
      reg_dif<R2>(x);
#pragma unroll
      for (int pos = 0; pos < R2; ++pos) {
        const int k = brev(pos, LOGR2);
        blk[8 * pos + qx[pos]] = k == 0 ? x[pos] : cmul(x[pos], tb[k]);
      }
    }
    __syncthreads();
    // ------------------------------- pass C + MAC against key rows p*l + lev -------------------------------
#pragma unroll
    for (int rb = 0; rb < L; ++rb) {
      if (rb + 1 < L) load_keys(kv[(rb + 1) & 1], rb + 1);
      const double2 *row = buf + rb * M;
      double2 x[8];
#pragma unroll
      for (int m = 0; m < 8; ++m) x[m] = row[8 * tid + (m ^ (tid & 7))];
      reg_dif<8>(x);
#pragma unroll
      for (int i = 0; i < 8; ++i) { cfma(f0[i], x[i], kv[rb & 1][i]); cfma(f1[i], x[i], kv[rb & 1][8 + i]); }
    }
    // ------------------------------- exchange: the other polynomial's partial sum goes to the peer ------------
    double2 mine[8];
#pragma unroll
    for (int m = 0; m < 8; ++m) {
      const double2 give = p_own ? f0[m] : f1[m];
      mine[m] = p_own ? f1[m] : f0[m];
      st_cluster(xbuf_peer + (unsigned)((parity * M + m * T + tid) * sizeof(double2)), give);
    }
    cluster_barrier();                                // also orders this CTA's pass-C reads before the row-0 reuse
#pragma unroll
    for (int m = 0; m < 8; ++m) mine[m] = cadd(mine[m], xbuf[parity * M + m * T + tid]);
    parity ^= 1;
    // ------------------------------- inverse of the owned polynomial: C' -> B' -> A' ------------------------
    reg_dit_inv<8>(mine);
#pragma unroll
    for (int m = 0; m < 8; ++m) buf[8 * tid + (m ^ (tid & 7))] = mine[m];
    __syncthreads();
    constexpr int TASKS_BI = 128 / T > 0 ? 128 / T : 1;
    static_assert(T <= 128, "B' assumes at least one task per thread");
#pragma unroll
    for (int it = 0; it < TASKS_BI; ++it) {
      double2 *blk = buf + (((it * T) & 127) >> 3) * S + bB0;
      double2 x[R2];
#pragma unroll
      for (int pos = 0; pos < R2; ++pos) {
        const int k = brev(pos, LOGR2);
        const double2 y = blk[8 * pos + qx[pos]];
        x[pos] = k == 0 ? y : cmul_conj(y, tb[k]);
      }
      reg_dit_inv<R2>(x);
#pragma unroll
      for (int m = 0; m < R2; ++m) blk[8 * m + qx[m]] = x[m];
    }
    __syncthreads();
    if (pA == 0) {
      double2 x[16];
#pragma unroll
      for (int g4 = 0; g4 < 4; ++g4) {
        double2 tw[4];
        tmem_ld4(tw, tw_taddr + 16 * g4);
#pragma unroll
        for (int i = 0; i < 4; ++i) x[4 * g4 + i] = cmul_conj(buf[(4 * g4 + i) * S + qsw[(4 * g4 + i) & (NVA - 1)]], tw[i]);
      }
      reg_dit_inv<16>(x);
#pragma unroll
      for (int m = 0; m < 16; ++m) {
        const double2 z = mul_w64(x[m], m, true);
        const int j = qA + m * S;
        acc[j] += f64_to_torus_fast(z.x * inv_M);
        acc[j + M] += f64_to_torus_fast(z.y * inv_M);
      }
    }
    __syncthreads();
  }

  // every thread is past its last tensor-memory access before the columns are released -- also when every step was
  // skipped and no step barrier ran
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  if (tid < 32) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 128;" ::"r"(tmem_base_s) : "memory");
  // ---- epilogue: extraction at index 0 (a part from polynomial 0, b from coefficient 0 of polynomial 1) or raw ----
  if (A.extract) {
    u64 *o = A.out + (size_t)ct * (N + 1);
    if (p_own == 0) {
      for (int c = tid; c < N; c += T) o[c] = (c == 0) ? acc[0] : (0ull - acc[N - c]);
    } else if (tid == 0) {
      o[N] = acc[0];
    }
  } else {
    u64 *o = A.out + (size_t)ct * 2 * N + (size_t)p_own * N;
    for (int c = tid; c < N; c += T) o[c] = acc[c];
  }
}

bool k1c_supported(const Params &p) {
  if (!k1_supported(p)) return false;
  const int logm = ilog2i(p.N) - 1;
  // pass A keeps the digits of two levels in one 32-bit word per coefficient
  return logm >= 9 && logm <= 10 && (p.l == 1 || 2 * p.Bg_bit <= 32);
}

void k1c_variant_name(const Params &p, char *dst, size_t cap) { snprintf(dst, cap, "k1c<N=%d,l=%d,cluster=2>", p.N, p.l); }

template <int LOGM, int L>
static void launch_k1c_one(const K1Args &a, int count, cudaStream_t st) {
  constexpr int M = 1 << LOGM;
  const size_t smem = (size_t)2 * M * 8 + (size_t)L * M * 16 + (size_t)2 * M * 16 + (((size_t)a.size * 2 + 15) & ~(size_t)15);
  static size_t configured_dev[MB_MAX_DEV] = {0};          // function attributes are per device
  size_t &configured = configured_dev[current_device()];
  if (smem > configured) {
    MB_REQUIRE(smem <= 227 * 1024, "k1c kernel: %zu B of shared memory needed (blind rotation too long)", smem);
    MB_CHECK(cudaFuncSetAttribute(blind_rotate_k1c_kernel<LOGM, L>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    configured = smem;
  }
  blind_rotate_k1c_kernel<LOGM, L><<<2 * count, M / 8, smem, st>>>(a);
  MB_CHECK(cudaGetLastError());
  count_launch();
}

void launch_blind_rotate_k1c(const BlindRotateLaunch &b, cudaStream_t st) {
  MB_REQUIRE(b.b_index == 0 || b.b_index == b.size, "segmented launches exist for the k1q kernel only");
  const Params &p = b.bsk->p;
  MB_REQUIRE(k1c_supported(p) && !b.direct, "k1c kernel: unsupported parameters");
  upload_w64();
  K1Args a{};
  a.bsk = b.bsk->d; a.tab = k1_tables_for(p.N); a.tv = b.tv; a.tv_count = b.tv_count; a.in = b.in;
  a.in_stride = b.in_stride; a.in_div = b.in_div > 0 ? b.in_div : 1; a.size = b.size; a.out = b.out; a.extract = b.extract;
  a.init_rotate = b.init_rotate; a.prec_offset = b.prec_offset; a.preprocess = b.preprocess; a.kappa = b.kappa;
  a.theta = b.theta; a.Bg_bit = p.Bg_bit; a.count = b.count;
  const int logm = ilog2i(p.N) - 1;
#define MB_K1C_CASE(LM, LL) if (logm == LM && p.l == LL) { launch_k1c_one<LM, LL>(a, b.count, st); return; }
  MB_K1C_CASE(9, 1) MB_K1C_CASE(9, 2) MB_K1C_CASE(9, 3) MB_K1C_CASE(9, 4)
  MB_K1C_CASE(10, 1) MB_K1C_CASE(10, 2) MB_K1C_CASE(10, 3) MB_K1C_CASE(10, 4)
#undef MB_K1C_CASE
  MB_FATAL("k1c kernel: no instantiation for N=%d l=%d", p.N, p.l);
}

}  // namespace mb
