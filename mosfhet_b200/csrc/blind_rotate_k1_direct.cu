// blind_rotate_k1_direct.cu -- ONE external product per ciphertext on the specialised k = 1 kernel
// (DIRECT instantiations of k1_kernel.cuh): trgsw_mul_trlwe_DFT + trlwe_from_DFT (trgsw.c:385-423,
// trlwe.c:629-634) and the CMUX of the leveled LUT (vertical_packing.c:24-33) as one launch
//     out[c] = in1[c] + TRGSW[sel] (.) (tv[c] - in1[c])        (in1 = null: the plain external product)
// These are what the compositions of SURVEY 8(f) spend their time in (CMUX tree, TRGSW-accumulator phase 2,
// unfolded rotation, circuit-bootstrap consumers).  One step of the bootstrap kernel costs ~9 us per CTA, so a
// batch is bound by moving its operands: 3 x 2N words per ciphertext through HBM.
#include "k1_kernel.cuh"

namespace mb {

bool k1_supported(const Params &p);

bool k1_direct_supported(const BlindRotateLaunch &b) {
  return b.direct && !b.dft_out && b.ks_mode == 0 && b.size == 1 && b.sub == b.add && k1_supported(b.bsk->p);
}

template <int LOGM, int L, int LB, bool PKALL>
static void launch_direct_one(const K1Args &a, int count, cudaStream_t st) {
  constexpr int M = 1 << LOGM;
  const size_t smem = (size_t)2 * 2 * M * 8 + (size_t)2 * LB * M * 16 + 16;
  static bool configured_dev[MB_MAX_DEV] = {false};        // function attributes are per device
  bool &configured = configured_dev[current_device()];
  if (!configured) {
    MB_CHECK(cudaFuncSetAttribute(blind_rotate_k1_kernel<LOGM, L, LB, 1, PKALL, 0, true>,
                                  cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    configured = true;
  }
  blind_rotate_k1_kernel<LOGM, L, LB, 1, PKALL, 0, true><<<count, M / 8, smem, st>>>(a);
  MB_CHECK(cudaGetLastError());
  count_launch();
}

void launch_extprod_k1(const BlindRotateLaunch &b, cudaStream_t st) {
  const Params &p = b.bsk->p;
  upload_w64();
  K1Args a{};
  a.bsk = b.bsk->d; a.tab = k1_tables_for(p.N); a.tv = b.tv; a.tv_count = b.tv_count;
  a.in_div = 1; a.size = 1; a.out = b.out; a.extract = b.extract; a.init_rotate = 0; a.preprocess = 0;
  a.Bg_bit = p.Bg_bit; a.count = b.count; a.sel = b.sel; a.sel_const = b.sel_const; a.in1 = b.add;   // b.sub == b.add
  MB_REQUIRE(a.sel_const >= 0 || a.sel != nullptr, "external product: no TRGSW selector");
  const int logm = ilog2i(p.N) - 1;
  // levels per shared-memory batch: 2 (1 at N = 4096 or when two digits exceed the 32-bit packed word)
  int lb = p.l < 2 ? p.l : 2;
  if (logm == 11 || 2 * p.Bg_bit > 32) lb = 1;
  const bool pkall = p.l * p.Bg_bit <= 32;
#define MB_K1D_CASE(LM, LL, LBB) \
  if (logm == LM && p.l == LL && lb == LBB) { \
    if (pkall) launch_direct_one<LM, LL, LBB, true>(a, b.count, st); else launch_direct_one<LM, LL, LBB, false>(a, b.count, st); \
    return; }
  MB_K1D_CASE(8, 1, 1) MB_K1D_CASE(8, 2, 2) MB_K1D_CASE(8, 2, 1) MB_K1D_CASE(8, 3, 2) MB_K1D_CASE(8, 3, 1) MB_K1D_CASE(8, 4, 2) MB_K1D_CASE(8, 4, 1)
  MB_K1D_CASE(9, 1, 1) MB_K1D_CASE(9, 2, 2) MB_K1D_CASE(9, 2, 1) MB_K1D_CASE(9, 3, 2) MB_K1D_CASE(9, 3, 1) MB_K1D_CASE(9, 4, 2) MB_K1D_CASE(9, 4, 1)
  MB_K1D_CASE(10, 1, 1) MB_K1D_CASE(10, 2, 2) MB_K1D_CASE(10, 2, 1) MB_K1D_CASE(10, 3, 2) MB_K1D_CASE(10, 3, 1) MB_K1D_CASE(10, 4, 2) MB_K1D_CASE(10, 4, 1)
  MB_K1D_CASE(11, 1, 1) MB_K1D_CASE(11, 2, 1) MB_K1D_CASE(11, 3, 1) MB_K1D_CASE(11, 4, 1)
#undef MB_K1D_CASE
  MB_FATAL("k1 external product: no instantiation for N=%d l=%d lb=%d", p.N, p.l, lb);
}

}  // namespace mb
