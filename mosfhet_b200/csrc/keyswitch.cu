// keyswitch.cu -- batched TLWE key switch (reference tlwe_keyswitch, tlwe.c:289-303):
//     out = (0, in.b) - sum_{i < N_in} sum_{j < t, d_ij != 0} KSK[i][j][d_ij - 1]
//     d_ij = digit j (base 2^base_bit, most significant first) of in.a[i] + 2^(63 - t*base_bit)
// Pure u64 wrap-around arithmetic: bit-exact with the reference whatever the summation order.
//
// Mapping: a CTA owns KS_G ciphertexts; thread c owns output columns c, c+T, ... of all of them in
// registers.  The table is swept in (i, j) order; for each (i, j) the CTA walks the digit values
// d = 1 .. 2^base_bit-1 that at least one of its ciphertexts selected, loads that row ONCE
// (coalesced, lanes on consecutive words) and subtracts it from every ciphertext that selected it.
// Across the grid all CTAs sweep the table in the same order, so the rows stream from HBM once
// per wave and are otherwise served by L2.
#include "common.cuh"
#include "device_math.cuh"

namespace mb {

constexpr int KS_THREADS = 256;
constexpr int KS_MAXCOLS = 4;      // columns per thread: supports n+1 <= 1024
constexpr int KS_CHUNK = 64;       // input coefficients staged per round

template <int G>
__global__ void __launch_bounds__(KS_THREADS) keyswitch_kernel(u64 *__restrict__ out, const u64 *__restrict__ in,
                                                               const u64 *__restrict__ ksk, int count, int n_in,
                                                               int n_out, int t, int base_bit, int row_stride) {
  // digit table for the current chunk: dig[g][e], e = (i_local * t + j)
  extern __shared__ unsigned char ks_smem[];
  unsigned char *dig = ks_smem;                                   // [G][KS_CHUNK * t]
  unsigned int *used = reinterpret_cast<unsigned int *>(ks_smem + ((G * KS_CHUNK * t + 15) & ~15));  // [KS_CHUNK * t] bitmask of digits in use
  const int ct0 = blockIdx.x * G;
  const int width = n_out + 1;
  const int bm1 = (1 << base_bit) - 1;
  const u64 prec_offset = 1ull << (64 - (1 + base_bit * t));

  u64 acc[G][KS_MAXCOLS];
#pragma unroll
  for (int g = 0; g < G; ++g)
#pragma unroll
    for (int q = 0; q < KS_MAXCOLS; ++q) {
      const int c = threadIdx.x + q * KS_THREADS;
      acc[g][q] = (c == n_out && ct0 + g < count) ? in[(size_t)(ct0 + g) * (n_in + 1) + n_in] : 0ull;
    }

  for (int i0 = 0; i0 < n_in; i0 += KS_CHUNK) {
    const int ni = min(KS_CHUNK, n_in - i0);
    __syncthreads();
    for (int e = threadIdx.x; e < ni * t; e += KS_THREADS) used[e] = 0u;
    __syncthreads();
    for (int w = threadIdx.x; w < G * ni; w += KS_THREADS) {
      const int g = w / ni, il = w - g * ni;
      const bool live = ct0 + g < count;
      const u64 ai = live ? in[(size_t)(ct0 + g) * (n_in + 1) + i0 + il] + prec_offset : 0ull;
      for (int j = 0; j < t; ++j) {
        const unsigned d = live ? (unsigned)((ai >> (64 - (j + 1) * base_bit)) & (u64)bm1) : 0u;
        dig[g * KS_CHUNK * t + il * t + j] = (unsigned char)d;
        if (d) atomicOr(&used[il * t + j], 1u << d);
      }
    }
    __syncthreads();
    for (int e = 0; e < ni * t; ++e) {
      unsigned m = used[e];
      const size_t base_row = ((size_t)(i0 * t + e)) * bm1;
      while (m) {
        const int d = __ffs(m) - 1;
        m &= m - 1;
        const u64 *row = ksk + (base_row + d - 1) * row_stride;
        u64 v[KS_MAXCOLS];
#pragma unroll
        for (int q = 0; q < KS_MAXCOLS; ++q) {
          const int c = threadIdx.x + q * KS_THREADS;
          v[q] = (c < width) ? __ldg(row + c) : 0ull;
        }
#pragma unroll
        for (int g = 0; g < G; ++g) {
          if (dig[g * KS_CHUNK * t + e] == d) {
#pragma unroll
            for (int q = 0; q < KS_MAXCOLS; ++q) acc[g][q] -= v[q];
          }
        }
      }
    }
  }
#pragma unroll
  for (int g = 0; g < G; ++g) {
    if (ct0 + g >= count) continue;
#pragma unroll
    for (int q = 0; q < KS_MAXCOLS; ++q) {
      const int c = threadIdx.x + q * KS_THREADS;
      if (c < width) out[(size_t)(ct0 + g) * width + c] = acc[g][q];
    }
  }
}

template <int G>
static void launch_ks_g(const KskDev *ksk, u64 *out, const u64 *in, int count, cudaStream_t st) {
  const Params &p = ksk->p;
  const size_t smem = ((size_t)G * KS_CHUNK * p.t + 15 & ~(size_t)15) + sizeof(unsigned) * KS_CHUNK * p.t;
  const int grid = (count + G - 1) / G;
  keyswitch_kernel<G><<<grid, KS_THREADS, smem, st>>>(out, in, ksk->d, count, p.k * p.N, p.n, p.t, p.base_bit,
                                                     ksk->row_stride);
  MB_CHECK(cudaGetLastError());
  count_launch();
}

void launch_keyswitch(const KskDev *ksk, u64 *out, const u64 *in, int count, cudaStream_t st) {
  const Params &p = ksk->p;
  MB_REQUIRE(p.n + 1 <= KS_THREADS * KS_MAXCOLS, "keyswitch: output dimension n=%d too large (max %d)", p.n,
             KS_THREADS * KS_MAXCOLS - 1);
  MB_REQUIRE(p.base_bit >= 1 && p.base_bit <= 5, "keyswitch: base_bit=%d unsupported (1..5)", p.base_bit);
  MB_REQUIRE(p.t * p.base_bit < 64, "keyswitch: t*base_bit must be < 64");
  if (count <= 0) return;
  // enough CTAs to fill the machine first, then amortise row loads over more ciphertexts per CTA
  const int sms = sm_count();
  if (count >= sms * 8 * 2) launch_ks_g<8>(ksk, out, in, count, st);
  else if (count >= sms * 4) launch_ks_g<4>(ksk, out, in, count, st);
  else if (count >= sms * 2) launch_ks_g<2>(ksk, out, in, count, st);
  else launch_ks_g<1>(ksk, out, in, count, st);
}

}  // namespace mb
