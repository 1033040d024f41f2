// multivalue.cu -- integer epilogues of the multi-value bootstrap:
//   multivalue_bootstrap_phase1 (bootstrap.c:232-243): one blind rotation of the constant test vector
//     1/(4*torus_base), then torus_base rotated copies  out[i] = out[0] * X^(i*N/torus_base)  and
//     out[torus_base] = out[0] * X^torus_base + out[0]
//   multivalue_bootstrap_phase2 (bootstrap.c:245-265): per LUT bit j, a +/- combination of the rotated
//     copies followed by trlwe_mv_extract_tlwe_scaling_addto(out, tmp, 1 << j) (trlwe.c:602-610).
// Pure u64 arithmetic: bit-exact with the reference given the same phase-1 accumulator.
#include "common.cuh"
#include "device_math.cuh"

namespace mb {

// out[ct][i][(k+1)][N], i <= torus_base, from src[ct][(k+1)][N]
__global__ void mv_phase1_rotations_kernel(u64 *out, const u64 *src, int N, int k, int torus_base) {
  const int ct = blockIdx.x, polys = k + 1;
  const u64 *s = src + (size_t)ct * polys * N;
  u64 *o = out + (size_t)ct * (torus_base + 1) * polys * N;
  for (int c = threadIdx.x; c < polys * N; c += blockDim.x) {
    const int p = c / N, i = c - p * N;
    const u64 v = s[c];
    o[c] = v;
    for (int r = 1; r < torus_base; ++r)
      o[(size_t)r * polys * N + c] = rotated_coeff(s + (size_t)p * N, i, (r * N / torus_base) & (2 * N - 1), N);
    o[(size_t)torus_base * polys * N + c] = rotated_coeff(s + (size_t)p * N, i, torus_base & (2 * N - 1), N) + v;
  }
}

// coefficient c (< k*N, or == k*N for b) of trlwe_extract_tlwe(tmp, idx) (trlwe.c:540-552)
__device__ __forceinline__ u64 extract_word(const u64 *tmp, int c, int idx, int N, int k) {
  if (c == k * N) return tmp[(size_t)k * N + idx];
  const int p = c / N, j = c - p * N;
  return (j <= idx) ? tmp[p * N + idx - j] : (0ull - tmp[p * N + N + idx - j]);
}

__global__ void mv_phase2_kernel(u64 *out, const int *lut, int lut_count, const u64 *rot, int N, int k,
                                 int torus_base, int log_torus_base) {
  extern __shared__ u64 tmp[];                       // [(k+1)*N]
  const int ct = blockIdx.x, polys = k + 1, W = polys * N;
  const int *in = lut + (size_t)(lut_count > 1 ? ct : 0) * torus_base;
  const u64 *r = rot + (size_t)ct * (torus_base + 1) * W;
  u64 *o = out + (size_t)ct * (k * N + 1);
  // each thread owns output words c = tid, tid + T, ...: accumulate in registers (<= 8 per thread)
  u64 acc[8];
#pragma unroll
  for (int q = 0; q < 8; ++q) acc[q] = 0;
  for (int j = 0; j < log_torus_base; ++j) {
    const int s0 = ((in[0] >> j) & 1) + ((in[torus_base - 1] >> j) & 1);
    __syncthreads();
    for (int c = threadIdx.x; c < W; c += blockDim.x) {
      u64 v = s0 == 2 ? r[(size_t)torus_base * W + c] : (s0 == 1 ? r[c] : 0ull);
      for (int i = 1; i < torus_base; ++i) {
        const int d = ((in[i] >> j) & 1) - ((in[i - 1] >> j) & 1);
        if (d == 1) v += r[(size_t)i * W + c];
        else if (d == -1) v -= r[(size_t)i * W + c];
      }
      tmp[c] = v;
    }
    __syncthreads();
    const int amount = 1 << j;
#pragma unroll
    for (int q = 0; q < 8; ++q) {
      const int c = threadIdx.x + q * blockDim.x;
      if (c > k * N) continue;
      u64 a = acc[q];
      for (int i = amount / 2; i < amount; ++i) a -= extract_word(tmp, c, N - 1 - (i - amount / 2), N, k);
      for (int i = 0; i < amount / 2; ++i) a += extract_word(tmp, c, i, N, k);
      acc[q] = a;
    }
  }
#pragma unroll
  for (int q = 0; q < 8; ++q) {
    const int c = threadIdx.x + q * blockDim.x;
    if (c <= k * N) o[c] = acc[q];
  }
}

// ---- the extraction family of the multi-ciphertext caller (trlwe.c:554-620, integer.c:94-100) --------------------------------
// Every member is a signed sum of sample extractions over two index ranges,
//     out = [acc ? out : 0] + sign * ( sum_{idx in [p_lo, p_hi]} extract(in, idx) - sum_{idx in [q_lo, q_hi]} extract(in, idx) ),
// with extract = trlwe_extract_tlwe (trlwe.c:540-552).  Pure u64 wrap-around arithmetic: bit-exact with the reference.
//   trlwe_extract_tlwe_addto / _subto (idx)        P = [idx, idx], Q empty, sign = +1 / -1, acc
//   trlwe_mv_extract_tlwe_scaling (s)              P = [0, s/2], Q = [N - (s - 1 - s/2), N - 1], no acc
//   trlwe_mv_extract_tlwe_scaling_addto / _subto   P = [0, s/2 - 1], Q = [N - (s - s/2), N - 1], sign = +1 / -1, acc
//   trlwe_mv_extract_tlwe (amount), output i       P = [i, i] (i < amount/2) or Q = [N-1-(i-amount/2)] (above), no acc
// `outs_per_in` > 1 (trlwe_mv_extract_tlwe): output o of input ct uses the o-th range quadruple of `ranges`.
struct MvExtractArgs {
  u64 *out;            // [count * outs_per_in][out_stride]
  const u64 *in;       // [count][(k+1)*N]
  const int *ranges;   // device [outs_per_in][4] = p_lo, p_hi, q_lo, q_hi (empty range: lo > hi)
  int N, k, out_stride, outs_per_in, sign, acc;
};
__global__ void mv_extract_kernel(MvExtractArgs A) {
  const int v = blockIdx.x, ct = v / A.outs_per_in, oi = v - ct * A.outs_per_in;
  const int N = A.N, k = A.k;
  const u64 *tr = A.in + (size_t)ct * (k + 1) * N;
  u64 *o = A.out + (size_t)v * A.out_stride;
  const int p_lo = A.ranges[4 * oi], p_hi = A.ranges[4 * oi + 1], q_lo = A.ranges[4 * oi + 2], q_hi = A.ranges[4 * oi + 3];
  for (int c = threadIdx.x; c <= k * N; c += blockDim.x) {
    u64 sum = 0;
    for (int idx = p_lo; idx <= p_hi; ++idx) sum += extract_word(tr, c, idx, N, k);
    for (int idx = q_lo; idx <= q_hi; ++idx) sum -= extract_word(tr, c, idx, N, k);
    const u64 base = A.acc ? o[c] : 0ull;
    o[c] = A.sign > 0 ? base + sum : base - sum;
  }
}
void launch_mv_extract(u64 *out, int out_stride, const u64 *in, const int *d_ranges, int outs_per_in, int N, int k, int sign,
                       int acc, int count, cudaStream_t st) {
  if (count <= 0 || outs_per_in <= 0) return;
  MvExtractArgs A;
  A.out = out; A.in = in; A.ranges = d_ranges; A.N = N; A.k = k; A.out_stride = out_stride; A.outs_per_in = outs_per_in;
  A.sign = sign; A.acc = acc;
  mv_extract_kernel<<<count * outs_per_in, 256, 0, st>>>(A);
  MB_CHECK(cudaGetLastError());
  count_launch();
}

void launch_mv_phase1_rotations(u64 *out, const u64 *src, int N, int k, int torus_base, int count, cudaStream_t st) {
  mv_phase1_rotations_kernel<<<count, 256, 0, st>>>(out, src, N, k, torus_base);
  MB_CHECK(cudaGetLastError());
  count_launch();
}

void launch_mv_phase2(u64 *out, const int *d_lut, int lut_count, const u64 *rot, int N, int k, int torus_base,
                      int log_torus_base, int count, cudaStream_t st) {
  const int threads = 1024;
  MB_REQUIRE(k * N + 1 <= threads * 8, "multivalue phase 2: k*N = %d too large", k * N);
  const size_t smem = sizeof(u64) * (k + 1) * N;
  static size_t configured_dev[MB_MAX_DEV] = {0};          // function attributes are per device
  size_t &configured = configured_dev[current_device()];
  if (smem > 48 * 1024 && smem > configured) {
    MB_CHECK(cudaFuncSetAttribute(mv_phase2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    configured = smem;
  }
  mv_phase2_kernel<<<count, threads, smem, st>>>(out, d_lut, lut_count, rot, N, k, torus_base, log_torus_base);
  MB_CHECK(cudaGetLastError());
  count_launch();
}

}  // namespace mb
