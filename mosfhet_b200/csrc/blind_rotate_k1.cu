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

#include "k1_kernel.cuh"

namespace mb {

// ---- per-N tables ------------------------------------------------------------------------------
static std::mutex g_k1_mu;
static std::map<long long, double2 *> g_k1_tab;     // (device, N) -> table

const double2 *k1_tables_for(int N) {
  ensure_init();
  std::lock_guard<std::mutex> lk(g_k1_mu);
  auto it = g_k1_tab.find(dev_key(N));
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
  MB_CHECK(cudaDeviceSynchronize());   // pageable H2D + non-blocking compute streams: fence once
  g_k1_tab[dev_key(N)] = d;
  return d;
}

// ---- dispatch --------------------------------------------------------------------------------------
// Variant = (levels per smem batch LB, min resident CTAs per SM MINB -> register cap, PKALL = all
// levels' digits packed once per step, PF = pass-C key prefetch mode).  The default per (N, l) is the
// fastest measured on B200 (profiles/); MB200_K1_LB / _PF select another instantiated one.
struct K1Variant { int lb, minb, pf; };

// Levels per shared-memory batch: as many as fit (a) the 32-bit packed-digit word, lb*Bg_bit <= 32, and
// (b) shared memory with >= 2 CTAs per SM at N = 2048 / 1 CTA at N = 4096.
static K1Variant default_variant(int logm, int l, int Bg_bit) {
  int lb = l;
  if (logm == 10 && lb > 2) lb = 2;
  if (logm == 11) lb = 1;
  while (lb > 1 && lb * Bg_bit > 32) --lb;
  if (lb == 3 && l == 4) lb = 2;                        // instantiated batch sizes: l, 2, 1
  // N = 1024, l = 3: batches of 2 + 1 levels need 49 KB of shared memory instead of 65 KB -> 4 CTAs per SM
  // instead of 3; 47.7 ms vs 51.7 ms per 4096 bootstraps (profiles/r1k_k1_occupancy.log)
  if (logm == 9 && l == 3 && lb == 3) lb = 2;
  // double-buffered key rows in pass C pay off while they fit the register file (profiles/r1b_k1_variants.log)
  const int pf = (logm <= 9 && lb == l && l <= 3) ? 1 : 0;
  return {lb, 1, pf};
}

static K1Variant chosen_variant(int logm, int l, int Bg_bit) {
  K1Variant v = default_variant(logm, l, Bg_bit);
  if (const char *e = getenv("MB200_K1_LB")) v.lb = atoi(e);
  if (const char *e = getenv("MB200_K1_PF")) v.pf = atoi(e);
  return v;
}

bool k1_supported(const Params &p) {
  if (p.k != 1) return false;
  const int logm = ilog2i(p.N) - 1;
  if (!(logm >= 8 && logm <= 11 && p.l >= 1 && p.l <= 4 && (1 << (logm + 1)) == p.N)) return false;
  // Bg_bit = 32 takes the generic kernel: the 32-bit digit mask and the 2^52 + Bg/2 bias of the specialised kernels are
  // formed in 32-bit arithmetic
  return p.Bg_bit >= 1 && p.Bg_bit <= 31;
}

void k1_variant_name(const Params &p, char *dst, size_t cap) {
  const int logm = ilog2i(p.N) - 1;
  const K1Variant v = chosen_variant(logm, p.l, p.Bg_bit);
  snprintf(dst, cap, "k1<N=%d,l=%d,lb=%d,minb=%d,pkall=%d,pf=%d>", p.N, p.l, v.lb, v.minb, (int)(p.l * p.Bg_bit <= 32), v.pf);
}

template <int LOGM, int L, int LB, int MINB, bool PKALL, int PF>
static void launch_one(const K1Args &a, int count, cudaStream_t st) {
  constexpr int M = 1 << LOGM;
  const size_t smem = (size_t)2 * 2 * M * 8 + (size_t)2 * LB * M * 16 + (((size_t)a.size * 2 + 15) & ~(size_t)15);
  static size_t configured_dev[MB_MAX_DEV] = {0};          // function attributes are per device
  size_t &configured = configured_dev[current_device()];
  if (smem > configured) {
    MB_REQUIRE(smem <= 227 * 1024, "k1 kernel: %zu B of shared memory needed (blind rotation too long)", smem);
    MB_CHECK(cudaFuncSetAttribute(blind_rotate_k1_kernel<LOGM, L, LB, MINB, PKALL, PF>,
                                  cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    configured = smem;
  }
  blind_rotate_k1_kernel<LOGM, L, LB, MINB, PKALL, PF><<<count, M / 8, smem, st>>>(a);
  MB_CHECK(cudaGetLastError());
  count_launch();
}

template <int LOGM, int L, int LB, int MINB, int PF>
static void launch_pk(const K1Args &a, int count, cudaStream_t st) {
  // all l levels fit one 32-bit word per coefficient -> pack once per step; otherwise once per batch
  if (L * a.Bg_bit <= 32) launch_one<LOGM, L, LB, MINB, true, PF>(a, count, st);
  else launch_one<LOGM, L, LB, MINB, false, PF>(a, count, st);
}

void launch_blind_rotate_k1(const BlindRotateLaunch &b, cudaStream_t st) {
  MB_REQUIRE(b.b_index == 0 || b.b_index == b.size, "segmented launches exist for the k1q kernel only");
  const Params &p = b.bsk->p;
  MB_REQUIRE(k1_supported(p) && !b.direct, "k1 kernel: unsupported parameters");
  upload_w64();
  K1Args a;
  a.bsk = b.bsk->d; a.tab = k1_tables_for(p.N); a.tv = b.tv; a.tv_count = b.tv_count; a.in = b.in;
  a.in_stride = b.in_stride; a.in_div = b.in_div > 0 ? b.in_div : 1; a.size = b.size; a.out = b.out; a.extract = b.extract; a.init_rotate = b.init_rotate;
  a.prec_offset = b.prec_offset; a.preprocess = b.preprocess; a.kappa = b.kappa; a.theta = b.theta; a.Bg_bit = p.Bg_bit;
  a.count = b.count;
  const int logm = ilog2i(p.N) - 1;
  const K1Variant v = chosen_variant(logm, p.l, p.Bg_bit);
  MB_REQUIRE(v.lb * p.Bg_bit <= 32, "k1 kernel: digits of one batch must fit 32 bits");
#define MB_K1_CASE(LM, LL, LBB, MB_, PF_) \
  if (logm == LM && p.l == LL && v.lb == LBB && v.minb == MB_ && v.pf == PF_) { launch_pk<LM, LL, LBB, MB_, PF_>(a, b.count, st); return; }
  // N = 512, 1024: batch = all levels (double-buffered keys for l <= 3), or 2 / 1 levels when l*Bg_bit > 32
  MB_K1_CASE(8, 1, 1, 1, 1) MB_K1_CASE(8, 2, 2, 1, 1) MB_K1_CASE(8, 3, 3, 1, 1) MB_K1_CASE(8, 4, 4, 1, 0)
  MB_K1_CASE(8, 2, 1, 1, 0) MB_K1_CASE(8, 3, 2, 1, 0) MB_K1_CASE(8, 3, 1, 1, 0) MB_K1_CASE(8, 4, 2, 1, 0) MB_K1_CASE(8, 4, 1, 1, 0)
  MB_K1_CASE(9, 1, 1, 1, 1) MB_K1_CASE(9, 2, 2, 1, 1) MB_K1_CASE(9, 3, 3, 1, 1) MB_K1_CASE(9, 4, 4, 1, 0)
  MB_K1_CASE(9, 2, 1, 1, 0) MB_K1_CASE(9, 3, 2, 1, 0) MB_K1_CASE(9, 3, 1, 1, 0) MB_K1_CASE(9, 4, 2, 1, 0) MB_K1_CASE(9, 4, 1, 1, 0)
  // N = 2048: at most 2 levels per batch (2 CTAs per SM); N = 4096: 1
  MB_K1_CASE(10, 1, 1, 1, 0) MB_K1_CASE(10, 2, 2, 1, 0) MB_K1_CASE(10, 3, 2, 1, 0) MB_K1_CASE(10, 4, 2, 1, 0)
  MB_K1_CASE(10, 2, 1, 1, 0) MB_K1_CASE(10, 3, 1, 1, 0) MB_K1_CASE(10, 4, 1, 1, 0)
  MB_K1_CASE(11, 1, 1, 1, 0) MB_K1_CASE(11, 2, 1, 1, 0) MB_K1_CASE(11, 3, 1, 1, 0) MB_K1_CASE(11, 4, 1, 1, 0)
#undef MB_K1_CASE
  MB_FATAL("k1 kernel: no instantiation for N=%d l=%d lb=%d minb=%d pf=%d", p.N, p.l, v.lb, v.minb, v.pf);
}

}  // namespace mb
