// keys.cu -- key residency: import of host-form keys into the resident HBM layouts, and
// device-side synthetic key generation for benchmarks/tests.
//
// Resident BSK layout (double2 = one complex slot):
//     bsk[i][r][q][idx],  i < n (LWE key bit), r < (k+1)*l (TRGSW row, reference order
//     trgsw.c:152-168), q < k+1 (a[0..k), b), idx < M = N/2
// with   idx = (s & 7) * (M/8) + (s >> 3)   <->   FFT position s   <->   root exponent
// e = 1 + 4*bitrev(s).  Position order is what a decimation-in-frequency transform produces
// without a reordering pass; the 8-way tiling makes the key loads of the thread that owns
// positions 8c..8c+7 (the radix-8 last pass of the k=1 kernel) land on consecutive 16-byte
// words across lanes, i.e. fully coalesced 128-bit loads.
#include <vector>

#include "common.cuh"
#include "device_math.cuh"

namespace mb {

static int bitrev_host(int x, int bits) {
  int r = 0;
  for (int i = 0; i < bits; ++i) { r = (r << 1) | (x & 1); x >>= 1; }
  return r;
}

int stored_index_of_position(int s, int M) { return M >= 8 ? (s & 7) * (M >> 3) + (s >> 3) : s; }
int position_of_stored_index(int idx, int M) { return M >= 8 ? ((idx % (M >> 3)) << 3) + idx / (M >> 3) : idx; }

// Root exponent of host slot h for each CPU FFT backend
void host_slot_exponents(int layout, int N, int32_t *e) {
  const int M = N / 2, bits = ilog2i(M);
  for (int h = 0; h < M; ++h) {
    int v;
    switch (layout) {
      case 1: v = 1 + 4 * bitrev_host(h, bits); break;   // SPQLIOS
      case 2: v = 1 - 4 * bitrev_host(h, bits); break;   // FFNT
      case 3: v = 1 + 4 * h; break;                      // natural
      default: MB_FATAL("host_slot_exponents: unknown layout %d", layout);
    }
    e[h] = ((v % (2 * N)) + 2 * N) % (2 * N);
  }
}

void slot_maps(int N, const int32_t *e, int *stored_to_host, int *stored_conj) {
  const int M = N / 2, bits = ilog2i(M);
  std::vector<char> seen(M, 0);
  for (int h = 0; h < M; ++h) {
    int ex = e[h], cj = 0;
    MB_REQUIRE((ex & 1) == 1 && ex > 0 && ex < 2 * N, "host slot %d: exponent %d is not an odd residue mod 2N", h, ex);
    if ((ex & 3) == 3) { ex = 2 * N - ex; cj = 1; }       // p(w^-e) = conj(p(w^e)) for real p
    const int kf = (ex - 1) / 4;
    const int s = bitrev_host(kf, bits);
    const int idx = stored_index_of_position(s, M);
    MB_REQUIRE(!seen[idx], "host slot order is not a permutation of the N/2 roots (slot %d)", h);
    seen[idx] = 1;
    stored_to_host[idx] = h;
    stored_conj[idx] = cj;
  }
}

// ---- import ----------------------------------------------------------------------------------
__global__ void import_bsk_kernel(double2 *dst, const double *src, const int *to_host, const int *conj,
                                  int N, size_t npolys) {
  const int M = N >> 1;
  for (size_t g = (size_t)blockIdx.x * blockDim.x + threadIdx.x; g < npolys * M; g += (size_t)gridDim.x * blockDim.x) {
    const size_t poly = g / M;
    const int idx = (int)(g - poly * M);
    const int h = to_host[idx];
    const double re = src[poly * N + h], im = src[poly * N + M + h];
    dst[g] = make_double2(re, conj[idx] ? -im : im);
  }
}

void import_bsk(BskDev *dst, const double *d_host_layout, const int32_t *h_exponents, cudaStream_t st) {
  const Params &p = dst->p;
  const int M = p.N / 2;
  std::vector<int> to_host(M), cj(M);
  slot_maps(p.N, h_exponents, to_host.data(), cj.data());
  int *d_maps = nullptr;
  MB_CHECK(cudaMalloc(&d_maps, sizeof(int) * 2 * M));
  MB_CHECK(cudaMemcpyAsync(d_maps, to_host.data(), sizeof(int) * M, cudaMemcpyHostToDevice, st));
  MB_CHECK(cudaMemcpyAsync(d_maps + M, cj.data(), sizeof(int) * M, cudaMemcpyHostToDevice, st));
  const size_t npolys = (size_t)p.n * (p.k + 1) * p.l * (p.k + 1);
  import_bsk_kernel<<<sm_count() * 8, 256, 0, st>>>(dst->d, d_host_layout, d_maps, d_maps + M, p.N, npolys);
  MB_CHECK(cudaGetLastError());
  count_launch();
  MB_CHECK(cudaStreamSynchronize(st));
  MB_CHECK(cudaFree(d_maps));
}

// ---- synthetic bootstrapping key ------------------------------------------------------------------
__device__ __forceinline__ double gaussian(u64 seed, u64 idx, double sigma) {
  const u64 r0 = splitmix64(seed ^ (idx * 2 + 0x1234567ull)), r1 = splitmix64(seed ^ (idx * 2 + 0x7654321ull));
  const double u1 = ((double)(r0 >> 11) + 1.0) * (1.0 / 9007199254740993.0);
  const double u2 = (double)(r1 >> 11) * (1.0 / 9007199254740992.0);
  return sigma * sqrt(-2.0 * log(u1)) * cospi(2.0 * u2);
}

// One CTA per TRLWE row of the torus-domain key: row = (i, r).  Writes (k+1) polynomials of N words.
//   a_q uniform, b = sum_q a_q * s_q + e (exact negacyclic product with a binary key), then the
//   gadget term m*h_lev on polynomial r/l, coefficient 0 (trgsw.c:152-168 / bootstrap.c:15-18).
__global__ void synth_trgsw_rows_kernel(u64 *out, const u64 *lwe_key, const u64 *rlwe_key, int N, int k, int l,
                                        int Bg_bit, double sigma, u64 seed) {
  extern __shared__ u64 sm[];          // a_q[N], s_q[N]
  u64 *sa = sm, *ss = sm + N;
  const int rows = (k + 1) * l;
  const size_t row = blockIdx.x;
  const int i = (int)(row / rows), r = (int)(row - (size_t)i * rows);
  u64 *o = out + row * (k + 1) * N;
  const int T = blockDim.x;
  // b starts as the noise
  for (int j = threadIdx.x; j < N; j += T)
    o[(size_t)k * N + j] = (u64)(i64)(gaussian(seed, row * N + j, sigma) * 18446744073709551616.0);
  for (int q = 0; q < k; ++q) {
    __syncthreads();
    for (int j = threadIdx.x; j < N; j += T) {
      const u64 v = splitmix64(seed + 0xABCDEFull + ((row * (k + 1) + q) * (u64)N + j) * 0x9E3779B97F4A7C15ull);
      sa[j] = v;
      ss[j] = rlwe_key[(size_t)q * N + j];
      o[(size_t)q * N + j] = v;
    }
    __syncthreads();
    for (int j = threadIdx.x; j < N; j += T) {
      u64 acc = 0;
      for (int t = 0; t < N; ++t) {
        if (ss[t]) {
          const int src = j - t;
          acc += (src >= 0) ? sa[src] : (0ull - sa[src + N]);
        }
      }
      o[(size_t)k * N + j] += acc;
    }
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    const int q = r / l, lev = r - q * l;
    const u64 h = 1ull << (64 - (lev + 1) * Bg_bit);
    o[(size_t)q * N] += lwe_key[i] * h;
  }
}

// position order [poly][N] (Re|Im) -> tiled resident order
__global__ void tile_bsk_kernel(double2 *dst, const double *src, int N, size_t npolys) {
  const int M = N >> 1, C8 = M >> 3;
  for (size_t g = (size_t)blockIdx.x * blockDim.x + threadIdx.x; g < npolys * M; g += (size_t)gridDim.x * blockDim.x) {
    const size_t poly = g / M;
    const int idx = (int)(g - poly * M);
    const int s = ((idx % C8) << 3) + idx / C8;
    dst[g] = make_double2(src[poly * N + s], src[poly * N + M + s]);
  }
}

// torus-domain TRGSW samples [n][(k+1)l][(k+1)][N] on the device -> resident Fourier layout (trgsw_to_DFT,
// trgsw.c:345, for a whole set at once); d_dft is caller scratch of npolys*N doubles
void bsk_from_torus(BskDev *dst, const u64 *d_torus, double *d_dft, cudaStream_t st) {
  const Params &p = dst->p;
  const size_t npolys = (size_t)p.n * (p.k + 1) * p.l * (p.k + 1);
  launch_torus_to_dft(d_dft, d_torus, p.N, (int)npolys, st);
  tile_bsk_kernel<<<sm_count() * 8, 256, 0, st>>>(dst->d, d_dft, p.N, npolys);
  MB_CHECK(cudaGetLastError());
  count_launch();
}

void synth_bsk(BskDev *dst, const u64 *h_lwe_key, const u64 *h_rlwe_key, double sigma, u64 seed, cudaStream_t st) {
  const Params &p = dst->p;
  MB_REQUIRE(p.N >= 16, "synth_bsk: N too small");
  const size_t rows_total = (size_t)p.n * (p.k + 1) * p.l;
  const size_t npolys = rows_total * (p.k + 1);
  u64 *d_lwe, *d_rlwe, *d_torus;
  double *d_dft;
  MB_CHECK(cudaMalloc(&d_lwe, sizeof(u64) * p.n));
  MB_CHECK(cudaMalloc(&d_rlwe, sizeof(u64) * p.k * p.N));
  MB_CHECK(cudaMalloc(&d_torus, sizeof(u64) * npolys * p.N));
  MB_CHECK(cudaMalloc(&d_dft, sizeof(double) * npolys * p.N));
  MB_CHECK(cudaMemcpyAsync(d_lwe, h_lwe_key, sizeof(u64) * p.n, cudaMemcpyHostToDevice, st));
  MB_CHECK(cudaMemcpyAsync(d_rlwe, h_rlwe_key, sizeof(u64) * p.k * p.N, cudaMemcpyHostToDevice, st));
  const size_t smem = sizeof(u64) * 2 * p.N;
  if (smem > 48 * 1024) MB_CHECK(cudaFuncSetAttribute(synth_trgsw_rows_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  synth_trgsw_rows_kernel<<<(unsigned)rows_total, 256, smem, st>>>(d_torus, d_lwe, d_rlwe, p.N, p.k, p.l, p.Bg_bit, sigma, seed);
  MB_CHECK(cudaGetLastError());
  count_launch();
  bsk_from_torus(dst, d_torus, d_dft, st);
  MB_CHECK(cudaStreamSynchronize(st));
  cudaFree(d_lwe); cudaFree(d_rlwe); cudaFree(d_torus); cudaFree(d_dft);
}

// ---- synthetic TRLWE-row key-switching keys (circuit bootstrap; k = 1) ------------------------------------
// One CTA per row [i][j][d-1], i < n_in + include_b:  a uniform, b = a*s_out + e + msg  with
//   packing1 (include_b = 0): msg = s_in[i]*d*2^(64-(j+1)b) at coefficient 0   (keyswitch.c:368-390)
//   private  (include_b = 1): msg = (-s_out) * (s_i*d*2^(64-(j+1)b)), s_i = s_in[i] or -1 for the b entry (:611-637)
__global__ void synth_gksk_kernel(u64 *out, const u64 *in_key, const u64 *out_key, int n_in, int include_b, int N,
                                  int t, int base_bit, double sigma, u64 seed) {
  extern __shared__ u64 sm[];
  u64 *sa = sm, *ss = sm + N;
  const int bm1 = (1 << base_bit) - 1;
  const size_t row = blockIdx.x;
  const int d = (int)(row % bm1) + 1, j = (int)((row / bm1) % t), i = (int)(row / ((size_t)bm1 * t));
  u64 *o = out + row * 2 * N;
  const int T = blockDim.x;
  const u64 s_i = i < n_in ? in_key[i] : ~0ull;                       // -1 for the b entry
  const u64 dec_key = s_i * (u64)d * (1ull << (64 - (j + 1) * base_bit));
  for (int c = threadIdx.x; c < N; c += T) {
    const u64 v = splitmix64(seed + 0x6b6bull + (row * (u64)N + c) * 0x9E3779B97F4A7C15ull);
    sa[c] = v; ss[c] = out_key[c]; o[c] = v;
  }
  __syncthreads();
  for (int c = threadIdx.x; c < N; c += T) {
    u64 acc = (u64)(i64)(gaussian(seed ^ 0x99ull, row * N + c, sigma) * 18446744073709551616.0);
    for (int q = 0; q < N; ++q)
      if (ss[q]) { const int src = c - q; acc += (src >= 0) ? sa[src] : (0ull - sa[src + N]); }
    if (include_b) acc += (0ull - ss[c]) * dec_key;
    else if (c == 0) acc += dec_key;
    o[N + c] = acc;
  }
}

void synth_gksk(u64 *d_out, const u64 *h_in_key, const u64 *h_out_key, int n_in, int include_b, int N, int t,
                int base_bit, double sigma, u64 seed, cudaStream_t st) {
  u64 *d_in, *d_ok;
  MB_CHECK(cudaMalloc(&d_in, sizeof(u64) * n_in));
  MB_CHECK(cudaMalloc(&d_ok, sizeof(u64) * N));
  MB_CHECK(cudaMemcpyAsync(d_in, h_in_key, sizeof(u64) * n_in, cudaMemcpyHostToDevice, st));
  MB_CHECK(cudaMemcpyAsync(d_ok, h_out_key, sizeof(u64) * N, cudaMemcpyHostToDevice, st));
  const size_t rows = (size_t)(n_in + include_b) * t * ((1 << base_bit) - 1);
  const size_t smem = sizeof(u64) * 2 * N;
  if (smem > 48 * 1024) MB_CHECK(cudaFuncSetAttribute(synth_gksk_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  synth_gksk_kernel<<<(unsigned)rows, 256, smem, st>>>(d_out, d_in, d_ok, n_in, include_b, N, t, base_bit, sigma, seed);
  MB_CHECK(cudaGetLastError());
  count_launch();
  MB_CHECK(cudaStreamSynchronize(st));
  cudaFree(d_in); cudaFree(d_ok);
}

// ---- synthetic key-switching table -------------------------------------------------------------------
// One warp per row KSK[i][j][d-1] = TLWE_out( s_in[i] * d * 2^(64-(j+1)*base_bit) )   (tlwe.c:202-209)
__global__ void synth_ksk_kernel(u64 *out, const u64 *in_key, const u64 *out_key, int n_in, int n_out, int t,
                                 int base_bit, int row_stride, double sigma, u64 seed) {
  const int bm1 = (1 << base_bit) - 1;
  const size_t nrows = (size_t)n_in * t * bm1;
  const int lane = threadIdx.x & 31;
  for (size_t row = ((size_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5; row < nrows; row += ((size_t)gridDim.x * blockDim.x) >> 5) {
    const int d = (int)(row % bm1) + 1;
    const int j = (int)((row / bm1) % t);
    const int i = (int)(row / ((size_t)bm1 * t));
    u64 *o = out + row * row_stride;
    u64 dot = 0;
    for (int c = lane; c < n_out; c += 32) {
      const u64 a = splitmix64(seed + 0x5151ull + (row * (u64)n_out + c) * 0x9E3779B97F4A7C15ull);
      o[c] = a;
      dot += a * out_key[c];
    }
    for (int off = 16; off; off >>= 1) dot += __shfl_xor_sync(0xffffffffu, dot, off);
    if (lane == 0) {
      const u64 msg = in_key[i] * (u64)d * (1ull << (64 - (j + 1) * base_bit));
      o[n_out] = dot + msg + (u64)(i64)(gaussian(seed ^ 0x77ull, row, sigma) * 18446744073709551616.0);
    }
    for (int c = n_out + 1 + lane; c < row_stride; c += 32) o[c] = 0;
  }
}

void synth_ksk(KskDev *dst, const u64 *h_in_key, const u64 *h_out_key, double sigma, u64 seed, cudaStream_t st) {
  const Params &p = dst->p;
  const int n_in = p.k * p.N;
  u64 *d_in, *d_out;
  MB_CHECK(cudaMalloc(&d_in, sizeof(u64) * n_in));
  MB_CHECK(cudaMalloc(&d_out, sizeof(u64) * p.n));
  MB_CHECK(cudaMemcpyAsync(d_in, h_in_key, sizeof(u64) * n_in, cudaMemcpyHostToDevice, st));
  MB_CHECK(cudaMemcpyAsync(d_out, h_out_key, sizeof(u64) * p.n, cudaMemcpyHostToDevice, st));
  synth_ksk_kernel<<<sm_count() * 8, 256, 0, st>>>(dst->d, d_in, d_out, n_in, p.n, p.t, p.base_bit, dst->row_stride, sigma, seed);
  MB_CHECK(cudaGetLastError());
  count_launch();
  MB_CHECK(cudaStreamSynchronize(st));
  cudaFree(d_in); cudaFree(d_out);
}

}  // namespace mb
