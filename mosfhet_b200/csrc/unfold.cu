// unfold.cu -- the per-group TRGSW of the unfolded blind rotation (reference bootstrap.c:132-140):
//     xai = su[g*2^u + 0] + sum_{j = 1}^{2^u - 1} X^{round(2N * sum_{bits of j} a[g*u + bit])} * su[g*2^u + j]
// where su[g*2^u + j] = TRGSW(prod_bits s^bit (1-s)^(1-bit)) is the key layout of bootstrap.c:23-48.
// Exact u64 wrap-around arithmetic (rotation = index shift + sign), so bit-exact with the reference whatever
// the summation order.  One CTA per (polynomial, group, ciphertext); HBM bound: 2^u polynomial reads per
// polynomial written.
#include "common.cuh"
#include "device_math.cuh"

namespace mb {

struct UnfoldArgs {
  u64 *out;           // [count][n_groups][npoly][N]
  const u64 *su;      // [groups_total * 2^u][npoly][N]
  const u64 *a;       // [count][a_stride]
  int a_stride;
  int N, npoly, unfolding, group0, n_groups;
};

__global__ void __launch_bounds__(256) unfold_kernel(UnfoldArgs A) {
  __shared__ int e[256];
  const int poly = blockIdx.x, gi = blockIdx.y, ct = blockIdx.z;
  const int g = A.group0 + gi, N = A.N, u = A.unfolding, key_exp = 1 << u;
  const int log_N2 = 31 - __clz(2 * N);
  const size_t T = (size_t)A.npoly * N;
  const u64 *base = A.su + (size_t)g * key_exp * T + (size_t)poly * N;
  const u64 *a = A.a + (size_t)ct * A.a_stride + (size_t)g * u;
  for (int j = threadIdx.x; j < key_exp; j += blockDim.x) {
    u64 sum = 0;
    for (int b = 0; b < u; ++b)
      if ((j >> b) & 1) sum += a[b];
    e[j] = (int)torus2int(sum, log_N2) & (2 * N - 1);          // bootstrap.c:135-139
  }
  __syncthreads();
  u64 *o = A.out + (((size_t)ct * A.n_groups + gi) * A.npoly + poly) * N;
  for (int c = threadIdx.x; c < N; c += blockDim.x) {
    u64 v = base[c];
    for (int j = 1; j < key_exp; ++j) v += rotated_coeff(base + (size_t)j * T, c, e[j], N);   // trgsw_mul_by_xai_addto
    o[c] = v;
  }
}

void launch_unfold(u64 *out, const u64 *su, const u64 *a, int a_stride, int N, int npoly, int unfolding, int group0,
                   int n_groups, int count, cudaStream_t st) {
  MB_REQUIRE(unfolding >= 1 && unfolding <= 8, "unfolding=%d unsupported (1..8)", unfolding);
  MB_REQUIRE(n_groups <= 65535 && count <= 65535, "unfold: grid too large (%d groups x %d ciphertexts)", n_groups, count);
  if (count <= 0 || n_groups <= 0) return;
  UnfoldArgs A;
  A.out = out; A.su = su; A.a = a; A.a_stride = a_stride; A.N = N; A.npoly = npoly; A.unfolding = unfolding;
  A.group0 = group0; A.n_groups = n_groups;
  unfold_kernel<<<dim3(npoly, n_groups, count), 256, 0, st>>>(A);
  MB_CHECK(cudaGetLastError());
  count_launch();
}

// ---- helpers of the batched leveled LUT (vertical_packing.c:36-52 over E independent evaluations) ---------------
// selector of CMUX c of a level: evaluation c % E uses its own TRGSW(bit) = set[(c % E) * size + bit]
__global__ void fill_sel_kernel(int *sel, int E, int size, int bit, int count) {
  for (int c = blockIdx.x * blockDim.x + threadIdx.x; c < count; c += gridDim.x * blockDim.x) sel[c] = (c % E) * size + bit;
}
void launch_fill_sel(int *sel, int E, int size, int bit, int count, cudaStream_t st) {
  fill_sel_kernel<<<(count + 255) / 256 < 1024 ? (count + 255) / 256 : 1024, 256, 0, st>>>(sel, E, size, bit, count);
  MB_CHECK(cudaGetLastError());
  count_launch();
}
// out = in * X^amount for `count` TRLWE samples of `polys` polynomials (trlwe_mul_by_xai, trlwe.c:507-513)
__global__ void rotate_trlwe_kernel(u64 *out, const u64 *in, int amount, int N, size_t total) {
  for (size_t g = (size_t)blockIdx.x * blockDim.x + threadIdx.x; g < total; g += (size_t)gridDim.x * blockDim.x) {
    const size_t poly = g / N;
    const int i = (int)(g - poly * N);
    out[g] = rotated_coeff(in + poly * N, i, amount, N);
  }
}
void launch_rotate_trlwe(u64 *out, const u64 *in, int amount, int N, int polys, int count, cudaStream_t st) {
  const size_t total = (size_t)count * polys * N;
  rotate_trlwe_kernel<<<sm_count() * 8, 256, 0, st>>>(out, in, amount & (2 * N - 1), N, total);
  MB_CHECK(cudaGetLastError());
  count_launch();
}

// programmable_bootstrap's input shaping (bootstrap.c:210-217) on a batch of TLWE words
__global__ void pb_preprocess_kernel(u64 *out, const u64 *in, size_t words, int kappa, int theta, int log_N2) {
  for (size_t g = (size_t)blockIdx.x * blockDim.x + threadIdx.x; g < words; g += (size_t)gridDim.x * blockDim.x)
    out[g] = pb_preprocess(in[g], kappa, theta, log_N2);
}
void launch_pb_preprocess(u64 *out, const u64 *in, size_t words, int kappa, int theta, int log_N2, cudaStream_t st) {
  if (words == 0) return;
  pb_preprocess_kernel<<<sm_count() * 4, 256, 0, st>>>(out, in, words, kappa, theta, log_N2);
  MB_CHECK(cudaGetLastError());
  count_launch();
}

}  // namespace mb
