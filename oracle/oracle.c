/*
 * oracle.c -- TEST INFRASTRUCTURE ONLY.  NOT part of the product, never linked into or called
 * from libmosfhet_b200.so / the mosfhet_b200 package.  Only tests/, __graft_entry__.smoke()
 * and bench.py's cpu_baseline / --impl reference legs may load this.
 *
 * A plain-C restatement of the reference's (antoniocgj/MOSFHET) programmable-bootstrap hot
 * path on flat arrays: gadget decomposition, negacyclic FFT pair, external product, blind
 * rotation, sample extraction, TLWE key switch -- each function cites the reference file:line
 * it follows.  It also carries an EXACT integer version of the external product / blind
 * rotation (no floating point) used as a second oracle for the FFT error.
 *
 * Parity is PINNED: tests/test_oracle_golden.py checks this file against fixtures produced by
 * the unmodified reference library built by oracle/build_ref.sh (tests/golden/make_golden.py),
 * and, when oracle/_ref is present, against the live reference on fresh random inputs.
 *
 * Flat layouts (same as include/mosfhet_b200.h section 3):
 *   TLWE(n)      : a[0..n) , b                       (n+1 words)
 *   TRLWE(k,N)   : a[0][N] .. a[k-1][N] , b[N]
 *   DFT poly     : N doubles = Re[0..N/2) | Im[0..N/2), slot s holds p(w^(1+4s)), w = e^{i pi/N}
 *                  ("natural" order; oracle_permute_* converts from/to a host FFT backend order)
 *   TRGSW_DFT    : [(k+1)*l rows][(k+1) polys][N doubles]
 *   BSK          : [n] TRGSW_DFT
 *   KSK          : [N_in][t][2^base_bit-1][n_out+1]
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <math.h>

typedef uint64_t Torus;

#ifndef M_PI
#define M_PI 3.14159265358979323846
#endif

/* ---------- scalar helpers ------------------------------------------------------------- */

/* misc.c:18-22  torus2int: round x * 2^log_scale / 2^64 to nearest */
uint64_t oracle_torus2int(Torus x, int log_scale) {
  const Torus round_offset = 1ULL << (64 - log_scale - 1);
  return (x + round_offset) >> (64 - log_scale);
}

/* misc.c:13-15  double2torus */
Torus oracle_double2torus(double x) { return (Torus)((int64_t)(18446744073709551616.0 * x)); }

static int ilog2(int x) { int r = 0; while ((1 << r) < x) r++; return r; }

static int bitrev(int x, int bits) {
  int r = 0;
  for (int i = 0; i < bits; i++) { r = (r << 1) | (x & 1); x >>= 1; }
  return r;
}

/* ---------- gadget decomposition -------------------------------------------------------- */

/* polynomial.c:74-89  polynomial_decompose_i -- signed digit i, rounding term included */
void oracle_decompose_i(int64_t *out, const Torus *in, int N, int Bg_bit, int l, int i) {
  const uint64_t half_Bg = 1ULL << (Bg_bit - 1);
  const uint64_t h_mask = (1ULL << Bg_bit) - 1;
  const uint64_t h_bit = 64 - (i + 1) * Bg_bit;
  uint64_t offset = 1ULL << (64 - l * Bg_bit - 1);
  for (int j = 0; j < l; j++) offset += 1ULL << (64 - j * Bg_bit - 1);
  for (int c = 0; c < N; c++) {
    const uint64_t v = in[c] + offset;
    out[c] = (int64_t)(((v >> h_bit) & h_mask) - half_Bg);
  }
}

/* ---------- rotations -------------------------------------------------------------------- */

/* polynomial.c:184-199  torus_polynomial_mul_by_xai: out = in * X^a, a taken mod 2N */
void oracle_mul_by_xai(Torus *out, const Torus *in, int N, int a) {
  a &= (2 * N - 1);
  for (int i = 0; i < N; i++) {
    int src = i - a;              /* in [-2N+1, N-1] */
    int neg = 0;
    while (src < 0) { src += N; neg ^= 1; }
    out[i] = neg ? (Torus)(0 - in[src]) : in[src];
  }
}

/* polynomial.c:220-235  torus_polynomial_mul_by_xai_minus_1: out = in * (X^a - 1) */
void oracle_mul_by_xai_minus_1(Torus *out, const Torus *in, int N, int a) {
  a &= (2 * N - 1);
  for (int i = 0; i < N; i++) {
    int src = i - a, neg = 0;
    while (src < 0) { src += N; neg ^= 1; }
    const Torus r = neg ? (Torus)(0 - in[src]) : in[src];
    out[i] = r - in[i];
  }
}

/* ---------- negacyclic FFT pair ----------------------------------------------------------- */
/* Follows polynomial.c:359-375 -> fft_processor_spqlios.c:81-97 (torus -> Fourier: cast to
 * signed, fold z_j = a_j + i a_{j+N/2}, twist by w^j, N/2-point transform, unscaled) and
 * fft_processor_spqlios.c:128-180 (Fourier -> torus: scale 2/N, inverse transform, untwist,
 * unfold, reduce mod 2^64).  The transform itself is restated as a textbook iterative radix-2
 * FFT; slot order here is the natural one (slot s <-> root exponent 1+4s).                   */

typedef struct { double re, im; } cplx;

static void fft_inplace(cplx *x, int M, int sign) { /* X_k = sum_j x_j e^{sign*2*pi*i*jk/M} */
  const int bits = ilog2(M);
  for (int i = 0; i < M; i++) {
    int j = bitrev(i, bits);
    if (j > i) { cplx t = x[i]; x[i] = x[j]; x[j] = t; }
  }
  for (int len = 2; len <= M; len <<= 1) {
    const int half = len >> 1;
    for (int blk = 0; blk < M; blk += len) {
      for (int j = 0; j < half; j++) {
        const double ang = sign * 2.0 * M_PI * (double)j / (double)len;
        const double wr = cos(ang), wi = sin(ang);
        cplx *u = &x[blk + j], *v = &x[blk + j + half];
        const double tr = v->re * wr - v->im * wi, ti = v->re * wi + v->im * wr;
        v->re = u->re - tr; v->im = u->im - ti;
        u->re += tr;        u->im += ti;
      }
    }
  }
}

void oracle_int_to_dft(double *out, const int64_t *in, int N) {
  const int M = N / 2;
  cplx *z = (cplx *)malloc(sizeof(cplx) * M);
  for (int j = 0; j < M; j++) {
    const double ang = M_PI * (double)j / (double)N;
    const double wr = cos(ang), wi = sin(ang);
    const double a = (double)in[j], b = (double)in[j + M];
    z[j].re = a * wr - b * wi;
    z[j].im = a * wi + b * wr;
  }
  fft_inplace(z, M, +1);
  for (int s = 0; s < M; s++) { out[s] = z[s].re; out[s + M] = z[s].im; }
  free(z);
}

/* polynomial_torus_to_DFT (polynomial.c:368-375) */
void oracle_torus_to_dft(double *out, const Torus *in, int N) {
  oracle_int_to_dft(out, (const int64_t *)in, N);
}

/* f64 -> u64 mod 2^64.  mode 0: round to nearest (AVX-512 path, fft_processor_spqlios.c:158-164);
 * mode 1: truncate the magnitude (scalar / FFNT path, fft_processor_spqlios.c:166-178,
 * ffnt.c:860-872).  Identical whenever |x| >= 2^52.                                           */
Torus oracle_f64_to_torus(double x, int mode) {
  uint64_t bits; memcpy(&bits, &x, 8);
  const uint64_t mant = (bits & 0x000FFFFFFFFFFFFFull) | 0x0010000000000000ull;
  const int expo = (int)((bits >> 52) & 0x7FF);
  const int sh = expo - 1075;
  uint64_t v;
  if (expo == 0) v = 0;
  else if (sh >= 64) v = 0;
  else if (sh >= 0) v = mant << sh;
  else if (sh <= -64) v = 0;
  else {
    v = mant >> (-sh);
    if (mode == 0) {
      const uint64_t rem = mant & ((1ULL << (-sh)) - 1), half = 1ULL << (-sh - 1);
      if (rem > half || (rem == half && (v & 1))) v++;   /* nearest, ties to even */
    }
  }
  return (bits >> 63) ? (Torus)(0 - v) : v;
}

/* polynomial_DFT_to_torus (polynomial.c:359-366) */
void oracle_dft_to_torus(Torus *out, const double *in, int N, int mode) {
  const int M = N / 2;
  const double scale = 2.0 / (double)N;
  cplx *z = (cplx *)malloc(sizeof(cplx) * M);
  for (int s = 0; s < M; s++) { z[s].re = in[s] * scale; z[s].im = in[s + M] * scale; }
  fft_inplace(z, M, -1);
  for (int j = 0; j < M; j++) {
    const double ang = -M_PI * (double)j / (double)N;
    const double wr = cos(ang), wi = sin(ang);
    const double re = z[j].re * wr - z[j].im * wi, im = z[j].re * wi + z[j].im * wr;
    out[j] = oracle_f64_to_torus(re, mode);
    out[j + M] = oracle_f64_to_torus(im, mode);
  }
  free(z);
}

/* Root exponent e_h of host slot h for each backend (verified against the
 * live reference in tests/test_oracle_golden.py).  layout: 1 SPQLIOS, 2 FFNT, 3 natural.      */
void oracle_slot_exponents(int layout, int N, int32_t *e) {
  const int M = N / 2, bits = ilog2(M);
  for (int h = 0; h < M; h++) {
    int v;
    if (layout == 1) v = 1 + 4 * bitrev(h, bits);
    else if (layout == 2) v = 1 - 4 * bitrev(h, bits);
    else v = 1 + 4 * h;
    e[h] = ((v % (2 * N)) + 2 * N) % (2 * N);
  }
}

/* host-order DFT polynomial -> natural order (exact: permutation, conjugation if e = 3 mod 4) */
void oracle_permute_from_host(double *nat, const double *host, int N, const int32_t *e) {
  const int M = N / 2;
  for (int h = 0; h < M; h++) {
    int ex = e[h];
    if ((ex & 3) == 1) { int s = (ex - 1) / 4; nat[s] = host[h]; nat[s + M] = host[h + M]; }
    else { int s = ((2 * N - ex) - 1) / 4; nat[s] = host[h]; nat[s + M] = -host[h + M]; }
  }
}

void oracle_permute_to_host(double *host, const double *nat, int N, const int32_t *e) {
  const int M = N / 2;
  for (int h = 0; h < M; h++) {
    int ex = e[h];
    if ((ex & 3) == 1) { int s = (ex - 1) / 4; host[h] = nat[s]; host[h + M] = nat[s + M]; }
    else { int s = ((2 * N - ex) - 1) / 4; host[h] = nat[s]; host[h + M] = -nat[s + M]; }
  }
}

/* ---------- Fourier-domain multiply-accumulate -------------------------------------------- */

/* polynomial.c:379-426  polynomial_mul_DFT / polynomial_mul_addto_DFT (scalar branch) */
static void dft_mul(double *out, const double *a, const double *b, int N, int addto) {
  const int M = N / 2;
  for (int i = 0; i < M; i++) {
    const double re = a[i] * b[i] - a[i + M] * b[i + M];
    const double im = a[i + M] * b[i] + a[i] * b[i + M];
    if (addto) { out[i] += re; out[i + M] += im; } else { out[i] = re; out[i + M] = im; }
  }
}

/* trgsw.c:385-423  trgsw_mul_trlwe_DFT: out_dft = sum_rows FFT(digit_j(poly_i)) * row[i*l+j] */
void oracle_trgsw_mul_trlwe_dft(double *out_dft, const Torus *in, const double *trgsw,
                                int N, int k, int l, int Bg_bit) {
  int64_t *dec = (int64_t *)malloc(sizeof(int64_t) * N);
  double *dec_dft = (double *)malloc(sizeof(double) * N);
  int first = 1;
  for (int i = 0; i <= k; i++) {            /* a[0..k) then b, rows i*l + j */
    for (int j = 0; j < l; j++) {
      oracle_decompose_i(dec, in + (size_t)i * N, N, Bg_bit, l, j);
      oracle_int_to_dft(dec_dft, dec, N);
      const double *row = trgsw + (size_t)(i * l + j) * (k + 1) * N;
      for (int q = 0; q <= k; q++)
        dft_mul(out_dft + (size_t)q * N, row + (size_t)q * N, dec_dft, N, !first);
      first = 0;
    }
  }
  free(dec); free(dec_dft);
}

/* trlwe.c:629-634  trlwe_from_DFT */
void oracle_trlwe_from_dft(Torus *out, const double *in_dft, int N, int k, int mode) {
  for (int q = 0; q <= k; q++) oracle_dft_to_torus(out + (size_t)q * N, in_dft + (size_t)q * N, N, mode);
}

/* ---------- blind rotation / bootstraps ---------------------------------------------------- */

/* bootstrap.c:107-122  blind_rotate (in place on acc) */
void oracle_blind_rotate(Torus *acc, const Torus *a, const double *bsk, int size,
                         int N, int k, int l, int Bg_bit, int mode) {
  const int log_N2 = ilog2(2 * N);
  const size_t trgsw_sz = (size_t)(k + 1) * l * (k + 1) * N;
  Torus *rot = (Torus *)malloc(sizeof(Torus) * (k + 1) * N);
  double *tmp = (double *)malloc(sizeof(double) * (k + 1) * N);
  for (int i = 0; i < size; i++) {
    const int ai = (int)oracle_torus2int(a[i], log_N2);
    if (!ai) continue;
    for (int q = 0; q <= k; q++) oracle_mul_by_xai_minus_1(rot + (size_t)q * N, acc + (size_t)q * N, N, ai);
    oracle_trgsw_mul_trlwe_dft(tmp, rot, bsk + (size_t)i * trgsw_sz, N, k, l, Bg_bit);
    oracle_trlwe_from_dft(rot, tmp, N, k, mode);
    for (int c = 0; c < (k + 1) * N; c++) acc[c] += rot[c];   /* trlwe_addto, trlwe.c:437 */
  }
  free(rot); free(tmp);
}

/* bootstrap.c:192-198  functional_bootstrap_wo_extract (unfolding == 1 branch) */
void oracle_functional_bootstrap_wo_extract(Torus *out_trlwe, const Torus *tv, const Torus *in_tlwe,
                                            const double *bsk, int n, int N, int k, int l,
                                            int Bg_bit, int torus_base, int mode) {
  const int log_N2 = ilog2(2 * N);
  const Torus prec_offset = oracle_double2torus(1.0 / (4 * torus_base));
  const int rot = 2 * N - (int)oracle_torus2int(in_tlwe[n] + prec_offset, log_N2);
  for (int q = 0; q <= k; q++) oracle_mul_by_xai(out_trlwe + (size_t)q * N, tv + (size_t)q * N, N, rot);
  oracle_blind_rotate(out_trlwe, in_tlwe, bsk, n, N, k, l, Bg_bit, mode);
}

/* trlwe.c:540-552  trlwe_extract_tlwe */
void oracle_extract_tlwe(Torus *out_tlwe, const Torus *in_trlwe, int N, int k, int idx) {
  for (int i = 0; i < k; i++) {
    for (int j = 0; j <= idx; j++) out_tlwe[i * N + j] = in_trlwe[(size_t)i * N + idx - j];
    for (int j = idx + 1; j < N; j++) out_tlwe[i * N + j] = (Torus)(0 - in_trlwe[(size_t)i * N + N + idx - j]);
  }
  out_tlwe[(size_t)k * N] = in_trlwe[(size_t)k * N + idx];
}

/* bootstrap.c:200-206  functional_bootstrap */
void oracle_functional_bootstrap(Torus *out_tlwe, const Torus *tv, const Torus *in_tlwe,
                                 const double *bsk, int n, int N, int k, int l, int Bg_bit,
                                 int torus_base, int mode) {
  Torus *acc = (Torus *)malloc(sizeof(Torus) * (k + 1) * N);
  oracle_functional_bootstrap_wo_extract(acc, tv, in_tlwe, bsk, n, N, k, l, Bg_bit, torus_base, mode);
  oracle_extract_tlwe(out_tlwe, acc, N, k, 0);
  free(acc);
}

/* bootstrap.c:208-220  programmable_bootstrap: input pre-processing */
void oracle_programmable_preprocess(Torus *out_tlwe, const Torus *in_tlwe, int n, int N,
                                    int kappa, int theta) {
  const int log_N2 = ilog2(2 * N);
  const Torus rnd_os = 1ULL << (64 - log_N2 + theta - 1);
  const Torus theta_mask = ~((1ULL << (64 - log_N2 + theta)) - 1);
  for (int i = 0; i <= n; i++) out_tlwe[i] = ((in_tlwe[i] << kappa) + rnd_os) & theta_mask;
}

void oracle_programmable_bootstrap(Torus *out_tlwe, const Torus *tv, const Torus *in_tlwe,
                                   const double *bsk, int n, int N, int k, int l, int Bg_bit,
                                   int precision, int kappa, int theta, int mode) {
  Torus *tmp = (Torus *)malloc(sizeof(Torus) * (n + 1));
  oracle_programmable_preprocess(tmp, in_tlwe, n, N, kappa, theta);
  oracle_functional_bootstrap(out_tlwe, tv, tmp, bsk, n, N, k, l, Bg_bit, 1 << (precision - 1), mode);
  free(tmp);
}

/* bootstrap.c:222-230  multivalue_bootstrap_CLOT21: out[i] = extract(acc, i*slot_size) */
void oracle_multivalue_bootstrap_CLOT21(Torus *out_tlwes, const Torus *tv, const Torus *in_tlwe,
                                        const double *bsk, int n, int N, int k, int l, int Bg_bit,
                                        int torus_base, int n_luts, int mode) {
  const int slot_size = N / (n_luts * torus_base);
  Torus *acc = (Torus *)malloc(sizeof(Torus) * (k + 1) * N);
  oracle_functional_bootstrap_wo_extract(acc, tv, in_tlwe, bsk, n, N, k, l, Bg_bit, torus_base * n_luts, mode);
  for (int i = 0; i < n_luts; i++) oracle_extract_tlwe(out_tlwes + (size_t)i * (k * N + 1), acc, N, k, i * slot_size);
  free(acc);
}

/* ---------- TLWE key switch ------------------------------------------------------------------ */

/* tlwe.c:289-303  tlwe_keyswitch: out = (0, in.b) - sum_{i,j: d!=0} KSK[i][j][d-1] */
void oracle_tlwe_keyswitch(Torus *out, const Torus *in, const Torus *ksk,
                           int n_in, int n_out, int t, int base_bit) {
  const Torus prec_offset = 1ULL << (64 - (1 + base_bit * t));
  const Torus mask = (1ULL << base_bit) - 1;
  const int base_m1 = (1 << base_bit) - 1;
  memset(out, 0, sizeof(Torus) * n_out);
  out[n_out] = in[n_in];
  for (int i = 0; i < n_in; i++) {
    const Torus ai = in[i] + prec_offset;
    for (int j = 0; j < t; j++) {
      const Torus aij = (ai >> (64 - (j + 1) * base_bit)) & mask;
      if (aij != 0) {
        const Torus *row = ksk + (((size_t)i * t + j) * base_m1 + (aij - 1)) * (n_out + 1);
        for (int c = 0; c <= n_out; c++) out[c] -= row[c];
      }
    }
  }
}

/* ---------- phases (test helpers; tlwe.c:135-141, trlwe.c:324-331 restated exactly) -------- */

Torus oracle_tlwe_phase(const Torus *c, const Torus *s, int n) {
  Torus sa = 0;
  for (int i = 0; i < n; i++) sa += s[i] * c[i];
  return c[n] - sa;
}

/* exact negacyclic product accumulate: out += a * b (all mod 2^64), polynomial.c:253-263 */
static void naive_mul_addto(Torus *out, const Torus *a, const Torus *b, int N) {
  for (int i = 0; i < N; i++) {
    const Torus bi = b[i];
    if (!bi) continue;
    for (int j = i; j < N; j++) out[j] += a[j - i] * bi;
    for (int j = 0; j < i; j++) out[j] -= a[N + j - i] * bi;
  }
}

void oracle_trlwe_phase(Torus *out, const Torus *c, const Torus *s /* k*N */, int N, int k) {
  Torus *acc = (Torus *)calloc(N, sizeof(Torus));
  for (int i = 0; i < k; i++) naive_mul_addto(acc, c + (size_t)i * N, s + (size_t)i * N, N);
  for (int j = 0; j < N; j++) out[j] = c[(size_t)k * N + j] - acc[j];
  free(acc);
}

/* ---------- EXACT second oracle (no floating point) ------------------------------------------ */
/* External product with polynomial_decompose_i digits (rounding term included, unlike
 * trgsw_naive_mul_trlwe, trgsw.c:452-470 -- SURVEY 8(c)) against the TORUS-domain TRGSW.       */
void oracle_trgsw_mul_trlwe_exact(Torus *out, const Torus *in, const Torus *trgsw_torus,
                                  int N, int k, int l, int Bg_bit) {
  int64_t *dec = (int64_t *)malloc(sizeof(int64_t) * N);
  memset(out, 0, sizeof(Torus) * (k + 1) * N);
  for (int i = 0; i <= k; i++)
    for (int j = 0; j < l; j++) {
      oracle_decompose_i(dec, in + (size_t)i * N, N, Bg_bit, l, j);
      const Torus *row = trgsw_torus + (size_t)(i * l + j) * (k + 1) * N;
      for (int q = 0; q <= k; q++) naive_mul_addto(out + (size_t)q * N, row + (size_t)q * N, (const Torus *)dec, N);
    }
  free(dec);
}

void oracle_blind_rotate_exact(Torus *acc, const Torus *a, const Torus *bsk_torus, int size,
                               int N, int k, int l, int Bg_bit) {
  const int log_N2 = ilog2(2 * N);
  const size_t trgsw_sz = (size_t)(k + 1) * l * (k + 1) * N;
  Torus *rot = (Torus *)malloc(sizeof(Torus) * (k + 1) * N);
  Torus *prod = (Torus *)malloc(sizeof(Torus) * (k + 1) * N);
  for (int i = 0; i < size; i++) {
    const int ai = (int)oracle_torus2int(a[i], log_N2);
    if (!ai) continue;
    for (int q = 0; q <= k; q++) oracle_mul_by_xai_minus_1(rot + (size_t)q * N, acc + (size_t)q * N, N, ai);
    oracle_trgsw_mul_trlwe_exact(prod, rot, bsk_torus + (size_t)i * trgsw_sz, N, k, l, Bg_bit);
    for (int c = 0; c < (k + 1) * N; c++) acc[c] += prod[c];
  }
  free(rot); free(prod);
}

/* ---------- multi-value bootstrap ------------------------------------- */

/* bootstrap.c:232-243  multivalue_bootstrap_phase1: out[0..torus_base] TRLWEs */
void oracle_multivalue_phase1(Torus *out, const Torus *in_tlwe, const double *bsk, int n, int N, int k, int l,
                              int Bg_bit, int torus_base, int mode) {
  const size_t W = (size_t)(k + 1) * N;
  Torus *tv = (Torus *)calloc(W, sizeof(Torus));
  for (int i = 0; i < N; i++) tv[(size_t)k * N + i] = oracle_double2torus(1.0 / (4 * torus_base));
  oracle_functional_bootstrap_wo_extract(out, tv, in_tlwe, bsk, n, N, k, l, Bg_bit, torus_base, mode);
  for (int i = 1; i < torus_base; i++)
    for (int q = 0; q <= k; q++) oracle_mul_by_xai(out + i * W + (size_t)q * N, out + (size_t)q * N, N, i * N / torus_base);
  for (int q = 0; q <= k; q++) {
    oracle_mul_by_xai(out + torus_base * W + (size_t)q * N, out + (size_t)q * N, N, torus_base);
    for (int c = 0; c < N; c++) out[torus_base * W + (size_t)q * N + c] += out[(size_t)q * N + c];   /* trlwe_addto */
  }
  free(tv);
}

/* trlwe.c:554-578  trlwe_extract_tlwe_addto / _subto */
static void extract_acc(Torus *out, const Torus *in, int N, int k, int idx, int sign) {
  Torus *e = (Torus *)malloc(sizeof(Torus) * (k * N + 1));
  oracle_extract_tlwe(e, in, N, k, idx);
  for (int c = 0; c <= k * N; c++) out[c] = sign > 0 ? out[c] + e[c] : out[c] - e[c];
  free(e);
}

/* trlwe.c:554-578 as exported entry points: out (+|-)= trlwe_extract_tlwe(in, idx) */
void oracle_extract_tlwe_addto(Torus *out, const Torus *in, int N, int k, int idx) { extract_acc(out, in, N, k, idx, +1); }
void oracle_extract_tlwe_subto(Torus *out, const Torus *in, int N, int k, int idx) { extract_acc(out, in, N, k, idx, -1); }

/* trlwe.c:580-589  trlwe_mv_extract_tlwe: out[i] = extract(i) for i < amount/2, -extract(N-1-(i-amount/2)) above */
void oracle_mv_extract_tlwe(Torus *outs, const Torus *in, int N, int k, int amount) {
  const int W = k * N + 1;
  for (int i = 0; i < amount / 2; i++) oracle_extract_tlwe(outs + (size_t)i * W, in, N, k, i);
  for (int i = amount / 2; i < amount; i++) {
    Torus *o = outs + (size_t)i * W;
    oracle_extract_tlwe(o, in, N, k, N - 1 - (i - amount / 2));
    for (int c = 0; c < W; c++) o[c] = 0 - o[c];                 /* tlwe_negate */
  }
}

/* trlwe.c:591-600  trlwe_mv_extract_tlwe_scaling */
void oracle_mv_extract_tlwe_scaling(Torus *out, const Torus *in, int N, int k, int scale) {
  const int amount = scale;
  oracle_extract_tlwe(out, in, N, k, amount / 2);
  for (int i = amount / 2 + 1; i < amount; i++) extract_acc(out, in, N, k, N - 1 - (i - amount / 2), -1);
  for (int i = 0; i < amount / 2; i++) extract_acc(out, in, N, k, i, +1);
}

/* trlwe.c:602-610 / 612-620  trlwe_mv_extract_tlwe_scaling_addto / _subto (sign = +1 / -1) */
void oracle_mv_extract_tlwe_scaling_acc(Torus *out, const Torus *in, int N, int k, int scale, int sign) {
  const int amount = scale;
  for (int i = amount / 2; i < amount; i++) extract_acc(out, in, N, k, N - 1 - (i - amount / 2), -sign);
  for (int i = 0; i < amount / 2; i++) extract_acc(out, in, N, k, i, sign);
}

/* bootstrap.c:245-265  multivalue_bootstrap_phase2 (+ trlwe_mv_extract_tlwe_scaling_addto, trlwe.c:602-610) */
void oracle_multivalue_phase2(Torus *out_tlwe, const int *in, const Torus *rot, int N, int k, int torus_base,
                              int log_torus_base) {
  const size_t W = (size_t)(k + 1) * N;
  Torus *tmp = (Torus *)malloc(sizeof(Torus) * W);
  memset(out_tlwe, 0, sizeof(Torus) * (k * N + 1));
  for (int j = 0; j < log_torus_base; j++) {
    const int s0 = ((in[0] >> j) & 1) + ((in[torus_base - 1] >> j) & 1);
    if (s0 == 2) memcpy(tmp, rot + torus_base * W, sizeof(Torus) * W);
    else if (s0 == 1) memcpy(tmp, rot, sizeof(Torus) * W);
    else memset(tmp, 0, sizeof(Torus) * W);
    for (int i = 1; i < torus_base; i++) {
      const int d = ((in[i] >> j) & 1) - ((in[i - 1] >> j) & 1);
      if (d == 1) for (size_t c = 0; c < W; c++) tmp[c] += rot[i * W + c];
      else if (d == -1) for (size_t c = 0; c < W; c++) tmp[c] -= rot[i * W + c];
    }
    const int amount = 1 << j;
    for (int i = amount / 2; i < amount; i++) extract_acc(out_tlwe, tmp, N, k, N - 1 - (i - amount / 2), -1);
    for (int i = 0; i < amount / 2; i++) extract_acc(out_tlwe, tmp, N, k, i, +1);
  }
  free(tmp);
}

/* ---------- circuit bootstrap ------------------------------------------- */

/* keyswitch.c:458-475 trlwe_packing1_keyswitch (include_b = 0) and keyswitch.c:639-656 trlwe_priv_keyswitch
 * (include_b = 1): out = (0, in.b at b[0] | nothing) - sum over entries, digits of KSK[i][j][d-1]
 * table: [n_in + include_b][t][2^base_bit-1][(k+1)*N] */
void oracle_table_keyswitch_trlwe(Torus *out, const Torus *in_tlwe, const Torus *table, int n_in, int include_b,
                                  int N, int k, int t, int base_bit) {
  const size_t W = (size_t)(k + 1) * N;
  const Torus prec_offset = 1ULL << (64 - (1 + base_bit * t));
  const Torus mask = (1ULL << base_bit) - 1;
  const int bm1 = (1 << base_bit) - 1;
  memset(out, 0, sizeof(Torus) * W);
  if (!include_b) out[(size_t)k * N] = in_tlwe[n_in];
  for (int i = 0; i < n_in + include_b; i++) {
    const Torus ai = (i < n_in ? in_tlwe[i] : in_tlwe[n_in]) + prec_offset;
    for (int j = 0; j < t; j++) {
      const Torus aij = (ai >> (64 - (j + 1) * base_bit)) & mask;
      if (aij != 0) {
        const Torus *row = table + (((size_t)i * t + j) * bm1 + (aij - 1)) * W;
        for (size_t c = 0; c < W; c++) out[c] -= row[c];
      }
    }
  }
}

/* bootstrap.c:324-345 circuit_bootstrap_2 (out->l == key->l == l, gadget of the output = Bg_out) */
void oracle_circuit_bootstrap_2(Torus *out_trgsw, const Torus *in_tlwe, const double *bsk, const Torus *kska,
                                const Torus *kskb, int n, int N, int k, int l, int Bg_bit, int Bg_out, int t,
                                int base_bit, int mode) {
  const size_t W = (size_t)(k + 1) * N;
  const int slot = N / (2 * l);
  Torus *tv = (Torus *)calloc(W, sizeof(Torus));
  Torus *acc = (Torus *)malloc(sizeof(Torus) * W);
  Torus *tl = (Torus *)malloc(sizeof(Torus) * (k * N + 1));
  for (int i = 0; i < N; i++) {                        /* trlwe_torus_packing(tv, lut, 2l), trlwe.c:662-667 */
    const int s = i / slot;
    tv[(size_t)k * N + i] = (s >= l && s < 2 * l) ? (1ULL << (64 - (s - l + 1) * Bg_out)) : 0;
  }
  oracle_functional_bootstrap_wo_extract(acc, tv, in_tlwe, bsk, n, N, k, l, Bg_bit, 2 * l, mode);
  for (int i = 0; i < l; i++) {
    oracle_extract_tlwe(tl, acc, N, k, i * slot);
    oracle_table_keyswitch_trlwe(out_trgsw + (size_t)i * W, tl, kska, k * N, 1, N, k, t, base_bit);
    oracle_table_keyswitch_trlwe(out_trgsw + (size_t)(l + i) * W, tl, kskb, k * N, 0, N, k, t, base_bit);
  }
  free(tv); free(acc); free(tl);
}

/* keyswitch.c:162-193 trlwe_keyswitch: out = (0, in.b) - sum_{i<k_in, j<t} FFT(Dec_j(in.a[i])) * KSK[i][j]
 * ksk: natural-slot-order doubles [k_in][t][k_out+1][N]; in has k_in mask polynomials, out k_out. */
void oracle_trlwe_keyswitch(Torus *out, const Torus *in, const double *ksk, int N, int k_in, int k_out, int t,
                            int base_bit, int mode) {
  int64_t *dec = (int64_t *)malloc(sizeof(int64_t) * N);
  double *dec_dft = (double *)malloc(sizeof(double) * N);
  double *acc = (double *)calloc((size_t)(k_out + 1) * N, sizeof(double));
  Torus *as = (Torus *)malloc(sizeof(Torus) * (k_out + 1) * N);
  Torus *b = (Torus *)malloc(sizeof(Torus) * N);
  memcpy(b, in + (size_t)k_in * N, sizeof(Torus) * N);              /* in may alias out (keyswitch.c:57, 60) */
  for (int i = 0; i < k_in; i++) {
    for (int j = 0; j < t; j++) {
      oracle_decompose_i(dec, in + (size_t)i * N, N, base_bit, t, j);
      oracle_int_to_dft(dec_dft, dec, N);
      const double *row = ksk + (size_t)(i * t + j) * (k_out + 1) * N;
      for (int q = 0; q <= k_out; q++) dft_mul(acc + (size_t)q * N, row + (size_t)q * N, dec_dft, N, 1);
    }
  }
  oracle_trlwe_from_dft(as, acc, N, k_out, mode);
  for (int q = 0; q < k_out; q++)
    for (int c = 0; c < N; c++) out[(size_t)q * N + c] = 0 - as[(size_t)q * N + c];
  for (int c = 0; c < N; c++) out[(size_t)k_out * N + c] = b[c] - as[(size_t)k_out * N + c];
  free(dec); free(dec_dft); free(acc); free(as); free(b);
}

/* keyswitch.c:52-63 trlwe_priv_keyswitch_2 (k = 1): ksk2 = [2][t][2][N], [0] switches a, [1] switches -b */
void oracle_trlwe_priv_keyswitch_2(Torus *out, const Torus *in, const double *ksk2, int N, int t, int base_bit,
                                   int mode) {
  const size_t key_sz = (size_t)t * 2 * N;
  Torus *tmp = (Torus *)calloc((size_t)2 * N, sizeof(Torus));
  Torus *o = (Torus *)calloc((size_t)2 * N, sizeof(Torus));
  for (int c = 0; c < N; c++) { tmp[c] = 0 - in[N + c]; o[c] = in[c]; }
  oracle_trlwe_keyswitch(tmp, tmp, ksk2 + key_sz, N, 1, 1, t, base_bit, mode);
  oracle_trlwe_keyswitch(o, o, ksk2, N, 1, 1, t, base_bit, mode);
  for (int c = 0; c < 2 * N; c++) out[c] = o[c] + tmp[c];
  free(tmp); free(o);
}

/* bootstrap.c:309-322 circuit_bootstrap: one functional bootstrap (torus_base 2, LUT {0, h_i}) per level */
void oracle_circuit_bootstrap(Torus *out_trgsw, const Torus *in_tlwe, const double *bsk, const Torus *kska,
                              const Torus *kskb, int n, int N, int k, int l, int Bg_bit, int l_out, int Bg_out,
                              int t, int base_bit, int mode) {
  const size_t W = (size_t)(k + 1) * N;
  Torus *tv = (Torus *)calloc(W, sizeof(Torus));
  Torus *tl = (Torus *)malloc(sizeof(Torus) * (k * N + 1));
  for (int i = 0; i < l_out; i++) {
    for (int c = 0; c < N; c++) tv[(size_t)k * N + c] = (c >= N / 2) ? (1ULL << (64 - (i + 1) * Bg_out)) : 0;
    oracle_functional_bootstrap(tl, tv, in_tlwe, bsk, n, N, k, l, Bg_bit, 2, mode);
    oracle_table_keyswitch_trlwe(out_trgsw + (size_t)i * W, tl, kska, k * N, 1, N, k, t, base_bit);
    oracle_table_keyswitch_trlwe(out_trgsw + (size_t)(l_out + i) * W, tl, kskb, k * N, 0, N, k, t, base_bit);
  }
  free(tv); free(tl);
}

/* bootstrap.c:347-366 circuit_bootstrap_3 (k = 1): packing key switch, then the FFT-based private one */
void oracle_circuit_bootstrap_3(Torus *out_trgsw, const Torus *in_tlwe, const double *bsk, const double *kska2,
                                const Torus *kskb, int n, int N, int l, int Bg_bit, int Bg_out, int t_a,
                                int base_bit_a, int t_b, int base_bit_b, int mode) {
  const int k = 1;
  const size_t W = (size_t)(k + 1) * N;
  const int slot = N / (2 * l);
  Torus *tv = (Torus *)calloc(W, sizeof(Torus));
  Torus *acc = (Torus *)malloc(sizeof(Torus) * W);
  Torus *tl = (Torus *)malloc(sizeof(Torus) * (k * N + 1));
  for (int i = 0; i < N; i++) {
    const int s = i / slot;
    tv[(size_t)k * N + i] = (s >= l && s < 2 * l) ? (1ULL << (64 - (s - l + 1) * Bg_out)) : 0;
  }
  oracle_functional_bootstrap_wo_extract(acc, tv, in_tlwe, bsk, n, N, k, l, Bg_bit, 2 * l, mode);
  for (int i = 0; i < l; i++) {
    oracle_extract_tlwe(tl, acc, N, k, i * slot);
    oracle_table_keyswitch_trlwe(out_trgsw + (size_t)(l + i) * W, tl, kskb, k * N, 0, N, k, t_b, base_bit_b);
    oracle_trlwe_priv_keyswitch_2(out_trgsw + (size_t)i * W, out_trgsw + (size_t)(l + i) * W, kska2, N, t_a,
                                  base_bit_a, mode);
  }
  free(tv); free(acc); free(tl);
}

/* ---------- TRGSW-accumulator bootstrap (bootstrap.c:267-306) -------------------------------------- */

/* functional_bootstrap_trgsw_phase1: every row of the trivial TRGSW(1) (trgsw.c:130-142) goes through the
 * same blind rotation (blind_rotate_trgsw is blind_rotate row by row, trgsw.c:425-431).
 * out_torus: [(k+1)*l_out][(k+1)][N] torus rows; out_dft: the same rows through polynomial_torus_to_DFT
 * (natural slot order), either may be NULL. */
void oracle_functional_bootstrap_trgsw_phase1(Torus *out_torus, double *out_dft, const Torus *in_tlwe,
                                              const double *bsk, int n, int N, int k, int l, int Bg_bit, int l_out,
                                              int Bg_out, int torus_base, int mode) {
  const size_t W = (size_t)(k + 1) * N;
  const int rows = (k + 1) * l_out, log_N2 = ilog2(2 * N);
  const Torus prec_offset = oracle_double2torus(1.0 / (4 * torus_base));
  const int rot0 = 2 * N - (int)oracle_torus2int(in_tlwe[n] + prec_offset, log_N2);
  Torus *triv = (Torus *)malloc(sizeof(Torus) * W), *acc = (Torus *)malloc(sizeof(Torus) * W);
  for (int r = 0; r < rows; r++) {
    const int q = r / l_out, i = r - q * l_out;
    memset(triv, 0, sizeof(Torus) * W);
    triv[(size_t)q * N] = 1ULL << (64 - (i + 1) * Bg_out);
    for (int p = 0; p <= k; p++) oracle_mul_by_xai(acc + (size_t)p * N, triv + (size_t)p * N, N, rot0);
    oracle_blind_rotate(acc, in_tlwe, bsk, n, N, k, l, Bg_bit, mode);
    if (out_torus) memcpy(out_torus + (size_t)r * W, acc, sizeof(Torus) * W);
    if (out_dft)
      for (int p = 0; p <= k; p++) oracle_torus_to_dft(out_dft + (size_t)r * W + (size_t)p * N, acc + (size_t)p * N, N);
  }
  free(triv); free(acc);
}

/* functional_bootstrap_trgsw_phase2 (bootstrap.c:298-306): out = extract_0(TRGSW (.) tv) */
void oracle_functional_bootstrap_trgsw_phase2(Torus *out_tlwe, const double *trgsw_dft, const Torus *tv, int N, int k,
                                              int l, int Bg_bit, int mode) {
  double *tmp = (double *)malloc(sizeof(double) * (k + 1) * N);
  Torus *res = (Torus *)malloc(sizeof(Torus) * (k + 1) * N);
  oracle_trgsw_mul_trlwe_dft(tmp, tv, trgsw_dft, N, k, l, Bg_bit);
  oracle_trlwe_from_dft(res, tmp, N, k, mode);
  oracle_extract_tlwe(out_tlwe, res, N, k, 0);
  free(tmp); free(res);
}

/* ---------- unfolded blind rotation (bootstrap.c:23-48 key layout, 124-148 loop) --------------------- */

/* The TRGSW of one group of `unfolding` key bits: su[g*2^u + 0] + sum_{j>=1} X^{round(sum of selected a)} su[g*2^u + j]
 * (exact integer arithmetic), in the torus domain.  su: [size/u * 2^u][(k+1)l][(k+1)][N]. */
void oracle_unfold_group(Torus *xai, const Torus *a, const Torus *su, int group, int unfolding, int N, int k, int l) {
  const int key_exp = 1 << unfolding, log_N2 = ilog2(2 * N);
  const size_t T = (size_t)(k + 1) * l * (k + 1) * N, npoly = (size_t)(k + 1) * l * (k + 1);
  const Torus *base = su + (size_t)group * key_exp * T;
  Torus *rot = (Torus *)malloc(sizeof(Torus) * N);
  memcpy(xai, base, sizeof(Torus) * T);
  for (int j = 1; j < key_exp; j++) {
    Torus a_i = 0;
    for (int u = 0, j_ = j; u < unfolding; u++, j_ >>= 1)
      if (j_ & 1) a_i += a[group * unfolding + u];
    const int e = (int)oracle_torus2int(a_i, log_N2);
    for (size_t p = 0; p < npoly; p++) {
      oracle_mul_by_xai(rot, base + (size_t)j * T + p * N, N, e);
      for (int c = 0; c < N; c++) xai[p * N + c] += rot[c];
    }
  }
  free(rot);
}

void oracle_blind_rotate_unfolded(Torus *acc, const Torus *a, const Torus *su, int size, int unfolding, int N, int k,
                                  int l, int Bg_bit, int mode) {
  const size_t T = (size_t)(k + 1) * l * (k + 1) * N;
  Torus *xai = (Torus *)malloc(sizeof(Torus) * T);
  double *xai_dft = (double *)malloc(sizeof(double) * T);
  double *tmp = (double *)malloc(sizeof(double) * (k + 1) * N);
  for (int g = 0; g < size / unfolding; g++) {
    oracle_unfold_group(xai, a, su, g, unfolding, N, k, l);
    for (size_t p = 0; p < T / N; p++) oracle_torus_to_dft(xai_dft + p * N, xai + p * N, N);
    oracle_trgsw_mul_trlwe_dft(tmp, acc, xai_dft, N, k, l, Bg_bit);
    oracle_trlwe_from_dft(acc, tmp, N, k, mode);
  }
  free(xai); free(xai_dft); free(tmp);
}

/* functional_bootstrap_wo_extract with an unfolding > 1 key (bootstrap.c:192-198) */
void oracle_functional_bootstrap_unfolded_wo_extract(Torus *out_trlwe, const Torus *tv, const Torus *in_tlwe,
                                                     const Torus *su, int n, int unfolding, int N, int k, int l,
                                                     int Bg_bit, int torus_base, int mode) {
  const int log_N2 = ilog2(2 * N);
  const Torus prec_offset = oracle_double2torus(1.0 / (4 * torus_base));
  const int rot0 = 2 * N - (int)oracle_torus2int(in_tlwe[n] + prec_offset, log_N2);
  for (int p = 0; p <= k; p++) oracle_mul_by_xai(out_trlwe + (size_t)p * N, tv + (size_t)p * N, N, rot0);
  oracle_blind_rotate_unfolded(out_trlwe, in_tlwe, su, n, unfolding, N, k, l, Bg_bit, mode);
}
