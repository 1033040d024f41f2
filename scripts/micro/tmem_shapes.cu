// Ground truth for the tensor-memory load shapes: which (lane, column) words reach which thread register when data
// written with tcgen05.st.32x32b (thread t -> lane t) is read back with the 16x256b / 16x128b / 16x64b shapes.
//   nvcc -gencode arch=compute_100a,code=sm_100a -o tmem_shapes tmem_shapes.cu && ./tmem_shapes
#include <cstdio>
#include <cuda_runtime.h>

__global__ void probe(unsigned *out) {
  __shared__ unsigned base_s;
  const int t = threadIdx.x;
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 32;\n\t"
               "tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::"r"((unsigned)__cvta_generic_to_shared(&base_s)) : "memory");
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const unsigned base = base_s;
  unsigned w[16];
  for (int c = 0; c < 16; ++c) w[c] = (unsigned)(t * 256 + c);          // word id = lane*256 + column
  asm volatile("tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16};"
               ::"r"(base), "r"(w[0]), "r"(w[1]), "r"(w[2]), "r"(w[3]), "r"(w[4]), "r"(w[5]), "r"(w[6]), "r"(w[7]), "r"(w[8]),
               "r"(w[9]), "r"(w[10]), "r"(w[11]), "r"(w[12]), "r"(w[13]), "r"(w[14]), "r"(w[15]) : "memory");
  asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
  unsigned r[4];
  // 16x256b.x1: 16 lanes x 8 words, 4 registers per thread; lanes 0..15 then lanes 16..31
  for (int half = 0; half < 2; ++half) {
    asm volatile("tcgen05.ld.sync.aligned.16x256b.x1.b32 {%0,%1,%2,%3}, [%4];\n\ttcgen05.wait::ld.sync.aligned;"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(base + ((unsigned)(16 * half) << 16)) : "memory");
    for (int i = 0; i < 4; ++i) out[(half * 32 + t) * 4 + i] = r[i];
  }
  // 16x128b.x1: 16 lanes x 4 words, 2 registers per thread
  for (int half = 0; half < 2; ++half) {
    asm volatile("tcgen05.ld.sync.aligned.16x128b.x1.b32 {%0,%1}, [%2];\n\ttcgen05.wait::ld.sync.aligned;"
                 : "=r"(r[0]), "=r"(r[1]) : "r"(base + ((unsigned)(16 * half) << 16)) : "memory");
    for (int i = 0; i < 2; ++i) out[256 + (half * 32 + t) * 2 + i] = r[i];
  }
  // 16x64b.x1: 16 lanes x 2 words, 1 register per thread
  for (int half = 0; half < 2; ++half) {
    asm volatile("tcgen05.ld.sync.aligned.16x64b.x1.b32 {%0}, [%1];\n\ttcgen05.wait::ld.sync.aligned;"
                 : "=r"(r[0]) : "r"(base + ((unsigned)(16 * half) << 16)) : "memory");
    out[384 + half * 32 + t] = r[0];
  }
  __syncthreads();
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 32;" ::"r"(base) : "memory");
}

int main() {
  unsigned *d, h[448];
  cudaMalloc(&d, sizeof(h));
  cudaMemset(d, 0xff, sizeof(h));
  probe<<<1, 32>>>(d);
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) { printf("error: %s\n", cudaGetErrorString(e)); return 1; }
  cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost);
  const char *names[3] = {"16x256b.x1", "16x128b.x1", "16x64b.x1"};
  const int regs[3] = {4, 2, 1}, off[3] = {0, 256, 384};
  for (int s = 0; s < 3; ++s)
    for (int half = 0; half < 2; ++half) {
      printf("%s, address lane %d: thread -> (lane,column) per register\n", names[s], 16 * half);
      for (int t = 0; t < 32; ++t) {
        printf("  t%02d:", t);
        for (int i = 0; i < regs[s]; ++i) { unsigned v = h[off[s] + (half * 32 + t) * regs[s] + i]; printf(" (%2u,%2u)", v >> 8, v & 255); }
        printf("%s", (t % 4 == 3) ? "\n" : "");
      }
    }
  return 0;
}
