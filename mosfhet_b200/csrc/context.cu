// context.cu -- device selection, per-thread streams, twiddle-table cache, launch accounting.
#include <atomic>
#include <cmath>
#include <map>
#include <mutex>
#include <vector>

#include "common.cuh"

namespace mb {

static std::mutex g_mu;
// Devices.  mb200_init selects the PRIMARY device; a host thread works on the primary unless it has been bound to another
// one (the per-device workers of the multi-GPU mode, api.cu).  Everything that lives on a device -- streams, twiddle
// tables, kernel attributes, resident keys -- is kept per device.
static int g_primary = -1;
static thread_local int t_device = -1;
static int g_sm_count[MB_MAX_DEV] = {0};
static bool g_dev_ready[MB_MAX_DEV] = {false};
static std::atomic<unsigned long long> g_launches{0};
static std::map<long long, double2 *> g_tw;            // (device << 32 | N) -> table

void count_launch(int n) { g_launches.fetch_add((unsigned long long)n); }
unsigned long long launches() { return g_launches.load(); }
void reset_launches() { g_launches.store(0); }

int device_count_noabort() {
  int n = 0;
  cudaError_t e = cudaGetDeviceCount(&n);
  if (e != cudaSuccess) { cudaGetLastError(); return 0; }
  return n;
}

static void prepare_device(int device) {               // g_mu held
  const int n = device_count_noabort();
  MB_REQUIRE(n > 0, "no CUDA device visible: this library has no CPU fallback (B200 / sm_100a required)");
  MB_REQUIRE(device >= 0 && device < n && device < MB_MAX_DEV, "device %d requested but only %d visible", device, n);
  MB_CHECK(cudaSetDevice(device));
  if (g_dev_ready[device]) return;
  cudaDeviceProp prop;
  MB_CHECK(cudaGetDeviceProperties(&prop, device));
  MB_REQUIRE(prop.major >= 10, "device %d is sm_%d%d; kernels are built for sm_100a only", device, prop.major, prop.minor);
  g_sm_count[device] = prop.multiProcessorCount;
  g_dev_ready[device] = true;
}

void init_device(int device) {
  std::lock_guard<std::mutex> lk(g_mu);
  if (device < 0) device = 0;
  prepare_device(device);
  g_primary = device;
}

// binds the calling host thread to `device` (>= 0) or back to the primary (-1)
void bind_thread_to_device(int device) {
  std::lock_guard<std::mutex> lk(g_mu);
  if (device >= 0) prepare_device(device);
  t_device = device;
}

int current_device() { return t_device >= 0 ? t_device : g_primary; }
int primary_device() { return g_primary; }

void ensure_init() {
  const int d = current_device();
  if (d >= 0) { MB_CHECK(cudaSetDevice(d)); return; }
  init_device(0);
}

int sm_count() { ensure_init(); return g_sm_count[current_device()]; }

// one stream per (host thread, device), leaked deliberately at thread exit: the context may already be gone
static thread_local cudaStream_t t_stream[MB_MAX_DEV] = {nullptr};

cudaStream_t default_stream() {
  ensure_init();
  cudaStream_t &s = t_stream[current_device()];
  if (!s) MB_CHECK(cudaStreamCreateWithFlags(&s, cudaStreamNonBlocking));
  return s;
}

// a second stream of this host thread on its device (overlapping launches of the pipelined entry points)
static thread_local cudaStream_t t_stream2[MB_MAX_DEV] = {nullptr};
cudaStream_t second_stream() {
  ensure_init();
  cudaStream_t &s = t_stream2[current_device()];
  if (!s) MB_CHECK(cudaStreamCreateWithFlags(&s, cudaStreamNonBlocking));
  return s;
}

const double2 *twiddles_for(int N) {
  ensure_init();
  std::lock_guard<std::mutex> lk(g_mu);
  const long long key = ((long long)current_device() << 32) | (long long)N;
  auto it = g_tw.find(key);
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
  g_tw[key] = d;
  return d;
}

void drop_tables() {
  std::lock_guard<std::mutex> lk(g_mu);
  for (auto &kv : g_tw) cudaFree(kv.second);
  g_tw.clear();
}

}  // namespace mb
