// blind_rotate_generic.cu -- the any-(k, l, N) blind-rotation / external-product kernel.
//
// One CTA per ciphertext.  The TRLWE accumulator stays in shared memory for all `size` steps of
// the reference loop (bootstrap.c:112-119); each step fuses
//   trlwe_mul_by_xai_minus_1 (polynomial.c:220)  -> polynomial_decompose_i (polynomial.c:74)
//   -> polynomial_torus_to_DFT (polynomial.c:368) -> trlwe_DFT_mul[_addto]_by_polynomial (trlwe.c:491)
//   -> trlwe_from_DFT (trlwe.c:629)               -> trlwe_addto (trlwe.c:437)
// and the epilogue fuses trlwe_extract_tlwe(.., 0) (trlwe.c:540).  The same kernel in `direct`
// mode is the single external product trgsw_mul_trlwe_DFT (trgsw.c:385) with either a torus or
// a host-order Fourier-domain result.
//
// This is the correctness-first path (radix-2, every stage through shared memory); the
// specialised k = 1 kernel in blind_rotate_k1.cu is the fast one.  Shared-memory layout is SoA
// (re[] / im[]) with one pad word every 8 so that both the butterfly strides and the stride-8
// MAC accesses of the tiled key order are conflict-free.
#include "common.cuh"
#include "device_math.cuh"

namespace mb {

__device__ __forceinline__ int padi(int s) { return s + (s >> 3); }

struct GenArgs {
  const double2 *bsk;
  const double2 *tw;
  const u64 *tv;
  int tv_count;
  const u64 *in;
  int in_stride;
  int in_div;
  int size;
  u64 *out;
  int extract, init_rotate;
  u64 prec_offset;
  int preprocess, kappa, theta;
  int N, k, l, Bg_bit;
  int rows_batch;
  int direct;
  const int *sel;
  double *dft_out;
  const int *dft_perm;
  const int *dft_conj;
  const u64 *sub, *add;
  int sel_const;
  int ks_mode;
};

// in-place radix-2 DIF, positive exponent, `nb` polynomials of M complex points; output bit-reversed
__device__ void fft_dif_batch(double *re, double *im, int Mp, int M, int N, int nb, const double2 *__restrict__ tw) {
  for (int h = M >> 1; h >= 1; h >>= 1) {
    const int tw_step = N / h;  // angle pi*j/h
    for (int b = threadIdx.x; b < nb * (M >> 1); b += blockDim.x) {
      const int row = b / (M >> 1), bb = b - row * (M >> 1);
      const int j = bb & (h - 1);
      const int i0 = ((bb - j) << 1) + j, i1 = i0 + h;
      double *r = re + row * Mp, *q = im + row * Mp;
      const int p0 = padi(i0), p1 = padi(i1);
      const double ur = r[p0], ui = q[p0], vr = r[p1], vi = q[p1];
      const double2 w = __ldg(&tw[j * tw_step]);
      const double dr = ur - vr, di = ui - vi;
      r[p0] = ur + vr;
      q[p0] = ui + vi;
      r[p1] = fma(dr, w.x, -di * w.y);
      q[p1] = fma(dr, w.y, di * w.x);
    }
    __syncthreads();
  }
}

// in-place radix-2 DIT with conjugate twiddles: bit-reversed in, natural out (unscaled)
__device__ void fft_dit_inverse_batch(double *re, double *im, int Mp, int M, int N, int nb,
                                      const double2 *__restrict__ tw) {
  for (int h = 1; h < M; h <<= 1) {
    const int tw_step = N / h;
    for (int b = threadIdx.x; b < nb * (M >> 1); b += blockDim.x) {
      const int row = b / (M >> 1), bb = b - row * (M >> 1);
      const int j = bb & (h - 1);
      const int i0 = ((bb - j) << 1) + j, i1 = i0 + h;
      double *r = re + row * Mp, *q = im + row * Mp;
      const int p0 = padi(i0), p1 = padi(i1);
      const double ur = r[p0], ui = q[p0], xr = r[p1], xi = q[p1];
      const double2 w = __ldg(&tw[j * tw_step]);
      const double vr = fma(xr, w.x, xi * w.y);   // x * conj(w)
      const double vi = fma(xi, w.x, -xr * w.y);
      r[p0] = ur + vr;
      q[p0] = ui + vi;
      r[p1] = ur - vr;
      q[p1] = ui - vi;
    }
    __syncthreads();
  }
}

__global__ void __launch_bounds__(512, 1) blind_rotate_generic_kernel(GenArgs A) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int N = A.N, M = N >> 1, k = A.k, l = A.l, Bg_bit = A.Bg_bit;
  const int Mp = M + (M >> 3) + 1;
  const int polys = k + 1, rows = polys * l;
  const int log_N2 = 31 - __clz(2 * N);
  const int ct = blockIdx.x;

  u64 *acc = reinterpret_cast<u64 *>(smem_raw);                       // [polys][N]
  double *fre = reinterpret_cast<double *>(acc + polys * N);         // [rows_batch][Mp]
  double *fim = fre + A.rows_batch * Mp;
  double *are = fim + A.rows_batch * Mp;                             // [polys][Mp]
  double *aim = are + polys * Mp;

  const u64 *in = A.in ? A.in + (size_t)(ct / A.in_div) * A.in_stride : nullptr;
  const u64 *tv = A.tv + (size_t)(A.tv_count > 1 ? ct % A.tv_count : 0) * polys * N;

  // ---- initial accumulator -----------------------------------------------------------------
  int rot0 = 0;
  if (A.init_rotate) {
    u64 b = in[A.size];
    if (A.preprocess) b = pb_preprocess(b, A.kappa, A.theta, log_N2);
    rot0 = (2 * N - (int)torus2int(b + A.prec_offset, log_N2)) & (2 * N - 1);
  }
  const u64 *sub = A.sub ? A.sub + (size_t)ct * polys * N : nullptr;
  for (int c = threadIdx.x; c < polys * N; c += blockDim.x) {
    const int p = c / N, i = c - p * N;
    u64 v = rot0 ? rotated_coeff(tv + (size_t)p * N, i, rot0, N) : tv[c];
    if (sub) v -= sub[c];                            // CMUX operand in2 - in1 (trlwe_sub, vertical_packing.c:27)
    if (A.ks_mode == 2 && p == k) v = 0ull - v;      // trlwe_priv_keyswitch_2 switches -b (keyswitch.c:55)
    acc[c] = v;
  }
  __syncthreads();

  const u64 off = decomp_offset(Bg_bit, l);
  const double inv_M = 1.0 / (double)M;
  const int C8 = M >> 3;

  for (int step = 0; step < A.size; ++step) {
    int a_i = 0;
    if (!A.direct) {
      u64 av = in[step];
      if (A.preprocess) av = pb_preprocess(av, A.kappa, A.theta, log_N2);
      a_i = (int)torus2int(av, log_N2) & (2 * N - 1);
      if (a_i == 0) continue;                       // bootstrap.c:114 (uniform across the CTA)
    }
    const int key_idx = A.sel_const >= 0 ? A.sel_const : (A.sel ? A.sel[ct] : step);
    const double2 *key = A.bsk + (size_t)key_idx * rows * polys * M;

    for (int c = threadIdx.x; c < polys * Mp; c += blockDim.x) { are[c] = 0.0; aim[c] = 0.0; }

    // trlwe_keyswitch decomposes the mask polynomials only (keyswitch.c:176-182): the rows of b are not swept
    const int rows_used = A.ks_mode == 1 ? k * l : rows;
    for (int r0 = 0; r0 < rows_used; r0 += A.rows_batch) {
      const int nb = min(A.rows_batch, rows_used - r0);
      // -- rotate-minus-one, decompose, fold, twist -> FFT buffers
      for (int c = threadIdx.x; c < nb * M; c += blockDim.x) {
        const int rb = c / M, j = c - rb * M;
        const int r = r0 + rb, p = r / l, lev = r - p * l;
        const u64 *ap = acc + p * N;
        u64 v0, v1;
        if (A.direct) { v0 = ap[j]; v1 = ap[j + M]; }
        else {
          v0 = rotated_coeff(ap, j, a_i, N) - ap[j];
          v1 = rotated_coeff(ap, j + M, a_i, N) - ap[j + M];
        }
        const double d0 = (double)decomp_digit(v0 + off, Bg_bit, lev);
        const double d1 = (double)decomp_digit(v1 + off, Bg_bit, lev);
        const double2 w = __ldg(&A.tw[j]);
        fre[rb * Mp + padi(j)] = fma(d0, w.x, -d1 * w.y);
        fim[rb * Mp + padi(j)] = fma(d0, w.y, d1 * w.x);
      }
      __syncthreads();
      fft_dif_batch(fre, fim, Mp, M, N, nb, A.tw);
      // -- Fourier-domain MAC against the key rows (stored index idx <-> position 8*(idx%C8)+idx/C8)
      for (int idx = threadIdx.x; idx < M; idx += blockDim.x) {
        const int s = ((idx % C8) << 3) + idx / C8;
        const int ps = padi(s);
        for (int p = 0; p < polys; ++p) {
          double2 sum = make_double2(are[p * Mp + ps], aim[p * Mp + ps]);
          for (int rb = 0; rb < nb; ++rb) {
            const double2 f = make_double2(fre[rb * Mp + ps], fim[rb * Mp + ps]);
            const double2 kv = __ldg(&key[((size_t)(r0 + rb) * polys + p) * M + idx]);
            cfma(sum, f, kv);
          }
          are[p * Mp + ps] = sum.x;
          aim[p * Mp + ps] = sum.y;
        }
      }
      __syncthreads();
    }

    if (A.dft_out) {
      // trgsw_mul_trlwe_DFT boundary: Fourier-domain result in the HOST slot order
      double *o = A.dft_out + (size_t)ct * polys * N;
      for (int c = threadIdx.x; c < polys * M; c += blockDim.x) {
        const int p = c / M, idx = c - p * M;
        const int s = ((idx % C8) << 3) + idx / C8;
        const int h = A.dft_perm[idx];
        o[(size_t)p * N + h] = are[p * Mp + padi(s)];
        o[(size_t)p * N + M + h] = A.dft_conj[idx] ? -aim[p * Mp + padi(s)] : aim[p * Mp + padi(s)];
      }
      return;
    }

    fft_dit_inverse_batch(are, aim, Mp, M, N, polys, A.tw);
    // -- untwist, scale 2/N, reduce mod 2^64, accumulate
    for (int c = threadIdx.x; c < polys * M; c += blockDim.x) {
      const int p = c / M, j = c - p * M;
      const double2 w = __ldg(&A.tw[j]);
      const double zr = are[p * Mp + padi(j)] * inv_M, zi = aim[p * Mp + padi(j)] * inv_M;
      const double re = fma(zr, w.x, zi * w.y);      // z * conj(w)
      const double im = fma(zi, w.x, -zr * w.y);
      const u64 t0 = f64_to_torus(re), t1 = f64_to_torus(im);
      if (A.direct) { acc[p * N + j] = t0; acc[p * N + j + M] = t1; }
      else { acc[p * N + j] += t0; acc[p * N + j + M] += t1; }
    }
    __syncthreads();
  }

  // ---- epilogue ------------------------------------------------------------------------------
  if (A.extract) {
    u64 *o = A.out + (size_t)ct * (k * N + 1);
    for (int c = threadIdx.x; c < k * N; c += blockDim.x) {
      const int p = c / N, j = c - p * N;
      o[c] = (j == 0) ? acc[p * N] : (0ull - acc[p * N + N - j]);   // trlwe.c:540-552 with idx = 0
    }
    if (threadIdx.x == 0) o[k * N] = acc[k * N];
  } else {
    u64 *o = A.out + (size_t)ct * polys * N;
    const u64 *add = A.add ? A.add + (size_t)ct * polys * N : nullptr;   // CMUX: + in1 (trlwe_add, :30)
    if (A.ks_mode == 1) {          // out = (0, in.b) - sum  (keyswitch.c:184-186)
      for (int c = threadIdx.x; c < polys * N; c += blockDim.x) o[c] = (c >= k * N ? tv[c] : 0ull) - acc[c];
    } else if (A.ks_mode == 2) {   // both halves are (0, 0) - sum, added (keyswitch.c:56-61)
      for (int c = threadIdx.x; c < polys * N; c += blockDim.x) o[c] = 0ull - acc[c];
    } else {
      for (int c = threadIdx.x; c < polys * N; c += blockDim.x) o[c] = add ? acc[c] + add[c] : acc[c];
    }
  }
}

static size_t generic_smem_bytes(int N, int k, int rows_batch) {
  const int M = N / 2, Mp = M + M / 8 + 1, polys = k + 1;
  return (size_t)polys * N * 8 + (size_t)2 * rows_batch * Mp * 8 + (size_t)2 * polys * Mp * 8;
}

void launch_blind_rotate_generic(const BlindRotateLaunch &a, cudaStream_t st) {
  MB_REQUIRE(a.b_index == 0 || a.b_index == a.size, "segmented launches exist for the k1q kernel only");
  const Params &p = a.bsk->p;
  MB_REQUIRE(p.N >= 16 && (p.N & (p.N - 1)) == 0, "generic blind rotate: N=%d must be a power of two >= 16", p.N);
  const int rows = (p.k + 1) * p.l;
  int rb = rows;
  const size_t limit = 227 * 1024;
  while (rb > 1 && generic_smem_bytes(p.N, p.k, rb) > limit) --rb;
  const size_t smem = generic_smem_bytes(p.N, p.k, rb);
  MB_REQUIRE(smem <= limit, "generic blind rotate: (k+1)*N = %d does not fit in shared memory (%zu B needed)",
             (p.k + 1) * p.N, smem);
  static size_t configured_dev[MB_MAX_DEV] = {0};          // function attributes are per device
  size_t &configured = configured_dev[current_device()];
  if (smem > configured) {
    MB_CHECK(cudaFuncSetAttribute(blind_rotate_generic_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    configured = smem;
  }
  GenArgs g;
  g.bsk = a.bsk->d; g.tw = twiddles_for(p.N);
  g.tv = a.tv; g.tv_count = a.tv_count; g.in = a.in; g.in_stride = a.in_stride; g.in_div = a.in_div > 0 ? a.in_div : 1; g.size = a.size;
  g.out = a.out; g.extract = a.extract; g.init_rotate = a.init_rotate; g.prec_offset = a.prec_offset;
  g.preprocess = a.preprocess; g.kappa = a.kappa; g.theta = a.theta;
  g.N = p.N; g.k = p.k; g.l = p.l; g.Bg_bit = p.Bg_bit; g.rows_batch = rb;
  g.direct = a.direct; g.sel = a.sel; g.dft_out = a.dft_out; g.dft_perm = a.dft_perm; g.dft_conj = a.dft_conj;
  g.sub = a.sub; g.add = a.add; g.sel_const = a.direct ? a.sel_const : -1;
  g.ks_mode = a.direct ? a.ks_mode : 0;
  int threads = p.N / 2;
  if (threads > 512) threads = 512;
  if (threads < 64) threads = 64;
  blind_rotate_generic_kernel<<<a.count, threads, smem, st>>>(g);
  MB_CHECK(cudaGetLastError());
  count_launch();
}

// ---------------------------------------------------------------------------------------------
// Stand-alone negacyclic transforms (internal position order, untiled):
//   slot s of the output holds p(w^(1+4*bitrev(s))), Re in [s], Im in [s+M]
// ---------------------------------------------------------------------------------------------
__global__ void torus_to_dft_kernel(double *out, const u64 *in, int N, const double2 *tw) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int M = N >> 1, Mp = M + (M >> 3) + 1;
  double *re = reinterpret_cast<double *>(smem_raw), *im = re + Mp;
  const u64 *p = in + (size_t)blockIdx.x * N;
  for (int j = threadIdx.x; j < M; j += blockDim.x) {
    const double d0 = (double)(i64)p[j], d1 = (double)(i64)p[j + M];   // fft_processor_spqlios.c:81-90
    const double2 w = __ldg(&tw[j]);
    re[padi(j)] = fma(d0, w.x, -d1 * w.y);
    im[padi(j)] = fma(d0, w.y, d1 * w.x);
  }
  __syncthreads();
  fft_dif_batch(re, im, Mp, M, N, 1, tw);
  double *o = out + (size_t)blockIdx.x * N;
  for (int s = threadIdx.x; s < M; s += blockDim.x) { o[s] = re[padi(s)]; o[s + M] = im[padi(s)]; }
}

// perm == nullptr: input in internal position order; else input is in a host slot order and
// perm[s] / conj[s] give the host slot feeding position s.
__global__ void dft_to_torus_kernel(u64 *out, const double *in, int N, const double2 *tw, const int *perm,
                                    const int *conj) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int M = N >> 1, Mp = M + (M >> 3) + 1;
  double *re = reinterpret_cast<double *>(smem_raw), *im = re + Mp;
  const double *p = in + (size_t)blockIdx.x * N;
  for (int s = threadIdx.x; s < M; s += blockDim.x) {
    const int h = perm ? perm[s] : s;
    const double vi = p[h + M];
    re[padi(s)] = p[h];
    im[padi(s)] = (conj && conj[s]) ? -vi : vi;
  }
  __syncthreads();
  fft_dit_inverse_batch(re, im, Mp, M, N, 1, tw);
  const double inv_M = 1.0 / (double)M;
  u64 *o = out + (size_t)blockIdx.x * N;
  for (int j = threadIdx.x; j < M; j += blockDim.x) {
    const double2 w = __ldg(&tw[j]);
    const double zr = re[padi(j)] * inv_M, zi = im[padi(j)] * inv_M;
    o[j] = f64_to_torus(fma(zr, w.x, zi * w.y));
    o[j + M] = f64_to_torus(fma(zi, w.x, -zr * w.y));
  }
}

void launch_torus_to_dft(double *out, const u64 *in, int N, int count, cudaStream_t st) {
  const int M = N / 2, Mp = M + M / 8 + 1;
  const size_t smem = (size_t)2 * Mp * 8;
  MB_REQUIRE(smem <= 48 * 1024 || N <= 8192, "torus_to_dft: N too large");
  if (smem > 48 * 1024) MB_CHECK(cudaFuncSetAttribute(torus_to_dft_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  int threads = M / 2 < 64 ? 64 : (M / 2 > 512 ? 512 : M / 2);
  torus_to_dft_kernel<<<count, threads, smem, st>>>(out, in, N, twiddles_for(N));
  MB_CHECK(cudaGetLastError());
  count_launch();
}

void launch_dft_to_torus(u64 *out, const double *in, int N, int count, const int *perm, const int *conj,
                         cudaStream_t st) {
  const int M = N / 2, Mp = M + M / 8 + 1;
  const size_t smem = (size_t)2 * Mp * 8;
  if (smem > 48 * 1024) MB_CHECK(cudaFuncSetAttribute(dft_to_torus_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  int threads = M / 2 < 64 ? 64 : (M / 2 > 512 ? 512 : M / 2);
  dft_to_torus_kernel<<<count, threads, smem, st>>>(out, in, N, twiddles_for(N), perm, conj);
  MB_CHECK(cudaGetLastError());
  count_launch();
}

// position order (torus_to_dft_kernel output) -> the host FFT backend's slot order; perm[s] / conj[s] give the
// host slot fed by position s (the same maps dft_to_torus_kernel reads the other way)
__global__ void pos_to_host_order_kernel(double *out, const double *in, int N, size_t npolys, const int *perm,
                                         const int *conj) {
  const int M = N >> 1;
  for (size_t g = (size_t)blockIdx.x * blockDim.x + threadIdx.x; g < npolys * M; g += (size_t)gridDim.x * blockDim.x) {
    const size_t poly = g / M;
    const int s = (int)(g - poly * M);
    const int h = perm[s];
    out[poly * N + h] = in[poly * N + s];
    out[poly * N + M + h] = conj[s] ? -in[poly * N + M + s] : in[poly * N + M + s];
  }
}

void launch_pos_to_host_order(double *out, const double *in, int N, size_t npolys, const int *perm, const int *conj,
                              cudaStream_t st) {
  pos_to_host_order_kernel<<<sm_count() * 4, 256, 0, st>>>(out, in, N, npolys, perm, conj);
  MB_CHECK(cudaGetLastError());
  count_launch();
}

// ---------------------------------------------------------------------------------------------
// Sample extraction at arbitrary indices (trlwe.c:540-552), one CTA per (ciphertext, index)
// ---------------------------------------------------------------------------------------------
__global__ void extract_kernel(u64 *out, const u64 *trlwe, const int *idx_list, int idx_count, int N, int k) {
  const int ct = blockIdx.x / idx_count, which = blockIdx.x - ct * idx_count;
  const int idx = idx_list[which];
  const u64 *in = trlwe + (size_t)ct * (k + 1) * N;
  u64 *o = out + (size_t)blockIdx.x * (k * N + 1);
  for (int c = threadIdx.x; c < k * N; c += blockDim.x) {
    const int p = c / N, j = c - p * N;
    o[c] = (j <= idx) ? in[p * N + idx - j] : (0ull - in[p * N + N + idx - j]);
  }
  if (threadIdx.x == 0) o[k * N] = in[k * N + idx];
}

void launch_extract(u64 *out, const u64 *trlwe, const int *d_idx, int idx_count, int N, int k, int count,
                    cudaStream_t st) {
  extract_kernel<<<count * idx_count, 256, 0, st>>>(out, trlwe, d_idx, idx_count, N, k);
  MB_CHECK(cudaGetLastError());
  count_launch();
}

}  // namespace mb
