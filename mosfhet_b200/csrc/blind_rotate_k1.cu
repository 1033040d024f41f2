// blind_rotate_k1.cu -- specialised k = 1 blind-rotation kernel (placeholder until implemented).
#include "common.cuh"
namespace mb {
bool k1_supported(const Params &) { return false; }
void launch_blind_rotate_k1(const BlindRotateLaunch &, cudaStream_t) { MB_FATAL("k1 kernel not built"); }
const char *k1_variant_name(const Params &) { return "k1"; }
const double2 *k1_tables_for(int) { return nullptr; }
}  // namespace mb
