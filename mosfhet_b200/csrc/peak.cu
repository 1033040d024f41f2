// peak.cu -- measures the FP64 FMA throughput of the device the library is running on.
// MEASURED_PEAKS.json carries HBM and bf16 figures but no FP64 one; the
// blind-rotation kernel is FP64-pipe bound, so bench.py uses this number as its roofline denominator.
#include "../../include/mosfhet_b200.h"
#include "common.cuh"

namespace mb {

__global__ void __launch_bounds__(256) fp64_fma_kernel(double *out, int iters, double a, double b) {
  double x0 = threadIdx.x * 1e-9, x1 = x0 + 1.0, x2 = x0 + 2.0, x3 = x0 + 3.0;
  double x4 = x0 + 4.0, x5 = x0 + 5.0, x6 = x0 + 6.0, x7 = x0 + 7.0;
  for (int i = 0; i < iters; ++i) {
#pragma unroll
    for (int u = 0; u < 8; ++u) {
      x0 = fma(x0, a, b); x1 = fma(x1, a, b); x2 = fma(x2, a, b); x3 = fma(x3, a, b);
      x4 = fma(x4, a, b); x5 = fma(x5, a, b); x6 = fma(x6, a, b); x7 = fma(x7, a, b);
    }
  }
  const double s = ((x0 + x1) + (x2 + x3)) + ((x4 + x5) + (x6 + x7));
  if (s == 123.456) out[0] = s;   // never true; keeps the chain alive
}

}  // namespace mb

extern "C" double mb200_measure_fp64_tflops(int iters) {
  mb::ensure_init();
  cudaStream_t st = mb::default_stream();
  double *d = nullptr;
  MB_CHECK(cudaMalloc(&d, 8));
  const int blocks = mb::sm_count() * 8, threads = 256;
  cudaEvent_t e0, e1;
  MB_CHECK(cudaEventCreate(&e0));
  MB_CHECK(cudaEventCreate(&e1));
  double best = 0.0;
  for (int rep = 0; rep < 4; ++rep) {
    MB_CHECK(cudaEventRecord(e0, st));
    mb::fp64_fma_kernel<<<blocks, threads, 0, st>>>(d, iters, 0.999999, 1e-7);
    MB_CHECK(cudaGetLastError());
    MB_CHECK(cudaEventRecord(e1, st));
    MB_CHECK(cudaEventSynchronize(e1));
    float ms = 0.f;
    MB_CHECK(cudaEventElapsedTime(&ms, e0, e1));
    const double flops = 2.0 * 64.0 * (double)iters * blocks * threads;
    const double tf = flops / (ms * 1e-3) * 1e-12;
    if (rep > 0 && tf > best) best = tf;
  }
  cudaEventDestroy(e0); cudaEventDestroy(e1); cudaFree(d);
  return best;
}
