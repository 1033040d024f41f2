// context.cu -- device selection, per-thread streams, twiddle-table cache, launch accounting.
#include <atomic>
#include <cmath>
#include <map>
#include <mutex>
#include <vector>

#include "common.cuh"

namespace mb {

static std::mutex g_mu;
static int g_device = -1;
static int g_sm_count = 0;
static std::atomic<unsigned long long> g_launches{0};
static std::map<int, double2 *> g_tw;

void count_launch(int n) { g_launches.fetch_add((unsigned long long)n); }
unsigned long long launches() { return g_launches.load(); }
void reset_launches() { g_launches.store(0); }

int device_count_noabort() {
  int n = 0;
  cudaError_t e = cudaGetDeviceCount(&n);
  if (e != cudaSuccess) { cudaGetLastError(); return 0; }
  return n;
}

void init_device(int device) {
  std::lock_guard<std::mutex> lk(g_mu);
  const int n = device_count_noabort();
  MB_REQUIRE(n > 0, "no CUDA device visible: this library has no CPU fallback (B200 / sm_100a required)");
  if (device < 0) device = 0;
  MB_REQUIRE(device < n, "device %d requested but only %d visible", device, n);
  MB_CHECK(cudaSetDevice(device));
  cudaDeviceProp prop;
  MB_CHECK(cudaGetDeviceProperties(&prop, device));
  MB_REQUIRE(prop.major >= 10, "device %d is sm_%d%d; kernels are built for sm_100a only", device, prop.major, prop.minor);
  if (g_device != device) {
    // tables live on one device; switching devices drops them
    for (auto &kv : g_tw) cudaFree(kv.second);
    g_tw.clear();
  }
  g_device = device;
  g_sm_count = prop.multiProcessorCount;
}

void ensure_init() {
  if (g_device >= 0) { MB_CHECK(cudaSetDevice(g_device)); return; }
  init_device(0);
}

int current_device() { return g_device; }
int sm_count() { ensure_init(); return g_sm_count; }

struct ThreadStream {
  cudaStream_t s = nullptr;
  ~ThreadStream() { /* leaked deliberately at thread exit: the context may already be gone */ }
};
static thread_local ThreadStream t_stream;

cudaStream_t default_stream() {
  ensure_init();
  if (!t_stream.s) MB_CHECK(cudaStreamCreateWithFlags(&t_stream.s, cudaStreamNonBlocking));
  return t_stream.s;
}

const double2 *twiddles_for(int N) {
  ensure_init();
  std::lock_guard<std::mutex> lk(g_mu);
  auto it = g_tw.find(N);
  if (it != g_tw.end()) return it->second;
  std::vector<double2> h(N);
  for (int j = 0; j < N; ++j) {
    const long double ang = M_PIl * (long double)j / (long double)N;
    h[j].x = (double)cosl(ang);
    h[j].y = (double)sinl(ang);
  }
  double2 *d = nullptr;
  MB_CHECK(cudaMalloc(&d, sizeof(double2) * N));
  MB_CHECK(cudaMemcpy(d, h.data(), sizeof(double2) * N, cudaMemcpyHostToDevice));
  MB_CHECK(cudaDeviceSynchronize());   // pageable H2D + non-blocking compute streams: fence once
  g_tw[N] = d;
  return d;
}

void drop_tables() {
  std::lock_guard<std::mutex> lk(g_mu);
  for (auto &kv : g_tw) cudaFree(kv.second);
  g_tw.clear();
}

}  // namespace mb
