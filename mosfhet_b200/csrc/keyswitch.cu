// keyswitch.cu -- batched TLWE key switch (reference tlwe_keyswitch, tlwe.c:289-303):
//     out = (0, in.b) - sum_{i < N_in} sum_{j < t, d_ij != 0} KSK[i][j][d_ij - 1]
//     d_ij = digit j (base 2^base_bit, most significant first) of in.a[i] + 2^(63 - t*base_bit)
// Pure u64 wrap-around arithmetic: bit-exact with the reference whatever the summation order.
//
// Mapping: ONE WARP PER CIPHERTEXT, up to 28 ciphertexts per CTA, one CTA per SM.  A lane keeps the
// output words 2*(lane+32q), 2*(lane+32q)+1 (q < NV) of its ciphertext in registers for the whole
// sweep, so the 5 KB accumulator never touches memory.  For every (i, j) the warp computes its
// digit (warp-uniform, no divergence), and if it is non-zero reads that table row with NV coalesced
// 128-bit loads per lane (rows are padded to a multiple of 64 words = 512 B) and subtracts it.
// All warps of a CTA walk (i, j) in the same order and are re-aligned by a block barrier every few
// input coefficients, so the (at most 2^base_bit - 1) rows of the current (i, j) are fetched from L2
// once per CTA and then served to the other warps by L1; across the grid every CTA sweeps the table
// in the same order, so HBM sees each row about once per wave.
//
// Algorithmic work per ciphertext: N_in * t * (1 - 2^-base_bit) row subtractions of (n+1) words.
#include "common.cuh"
#include "device_math.cuh"

namespace mb {

constexpr int KS_MAX_WARPS = 28;     // 28 warps x 72 registers x 32 lanes = the whole register file
constexpr int KS_SYNC_EVERY = 4;     // input coefficients between block barriers (L1 window: 4*t*(2^b-1) rows)

__device__ __forceinline__ ulonglong2 ldg_u128(const u64 *p) {
  ulonglong2 v;
  asm volatile("ld.global.nc.v2.u64 {%0, %1}, [%2];" : "=l"(v.x), "=l"(v.y) : "l"(p));
  return v;
}

// The same kernel serves every table key switch of the reference:
//   tlwe_keyswitch (tlwe.c:289)            rows = TLWE(n) padded to 64 words, b added at word n
//   trlwe_packing1_keyswitch (keyswitch.c:458) rows = TRLWE (k+1)*N words, in.b added at word k*N (b[0])
//   trlwe_priv_keyswitch (keyswitch.c:639)  same rows, n+1 input entries (the last one is in.b), nothing added
// A warp handles one (ciphertext, input slice, column chunk): `chunks` > 1 splits rows wider than 1024
// words over several warps; `slices` > 1 (small batches) splits the sweep over the input coefficients,
// the partial sums being added into the zero-initialised output with 64-bit integer atomics -- exact
// and order independent either way.
struct TableKsArgs {
  u64 *out;
  const u64 *in;
  const u64 *table;
  int count;
  int n_entries;      // input words swept (n_in, or n_in + 1 when the key has an entry for b)
  int in_stride;      // words between consecutive input TLWEs (n_in + 1)
  int b_word;         // index of b in the input TLWE (n_in)
  int out_words;      // valid output words per ciphertext
  int out_stride;
  int b_index;        // output word that receives + in.b, or -1
  int t, base_bit, row_stride;
  int per_cta, slices, i_per, chunks;
};

// CHUNKED = false keeps the table pointer warp-uniform (a uniform register): the kernel sits exactly at the
// 72-register cap that 28 warps per SM allow, and two more live registers cost the fifth load in flight
// (measured: 4.2 ms instead of 3.4 ms per 4096 at Level 1).
template <int NV, bool CHUNKED>
__global__ void __launch_bounds__(KS_MAX_WARPS * 32, 1) keyswitch_warp_kernel(TableKsArgs A) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int vct = blockIdx.x * A.per_cta + warp;             // (ciphertext, chunk); blockIdx.y = slice of the sweep, so that
  const int ch = CHUNKED ? vct % A.chunks : 0, ct = CHUNKED ? vct / A.chunks : vct;   // the warps of a CTA read the SAME rows (L1)
  const int sl = blockIdx.y;
  const bool live = warp < A.per_cta && ct < A.count;
  const int t = A.t, base_bit = A.base_bit, row_stride = A.row_stride;
  const int bm1 = (1 << base_bit) - 1;
  const u64 prec_offset = 1ull << (64 - (1 + base_bit * t));
  const u64 *a = A.in + (size_t)(live ? ct : 0) * A.in_stride;
  const int i_begin = sl * A.i_per, i_end = min(A.n_entries, i_begin + A.i_per);
  const u64 *__restrict__ ksk = CHUNKED ? A.table + (size_t)ch * (64 * NV) : A.table;

  u64 acc[2 * NV];
#pragma unroll
  for (int q = 0; q < 2 * NV; ++q) acc[q] = 0ull;

  for (int i0 = i_begin; i0 < i_begin + A.i_per; i0 += 32) { // same trip count for every warp (block barriers inside)
    const int i_mine = i0 + lane;
    const u64 a_mine = (live && i_mine < i_end) ? a[i_mine] + prec_offset : 0ull;   // tlwe.c:297
    const int ni = min(32, i_begin + A.i_per - i0);
    for (int il = 0; il < ni; ++il) {
      const u64 ai = __shfl_sync(0xffffffffu, a_mine, il);
      const bool valid = live && (i0 + il) < i_end;
      const u64 *base_row = ksk + (size_t)(i0 + il) * t * bm1 * row_stride + 2 * lane;
      for (int j = 0; j < t; ++j) {
        const unsigned d = (unsigned)(ai >> (64 - (j + 1) * base_bit)) & (unsigned)bm1;
        if (d != 0 && valid) {                      // warp-uniform
          const u64 *row = base_row + (size_t)(j * bm1 + (int)d - 1) * row_stride;
#pragma unroll
          for (int q0 = 0; q0 < NV; q0 += 5) {      // 5 x 128-bit loads in flight keeps the kernel under 72 registers
            ulonglong2 v[5];
#pragma unroll
            for (int q = 0; q < 5; ++q)
              if (q0 + q < NV) v[q] = ldg_u128(row + 64 * (q0 + q));
#pragma unroll
            for (int q = 0; q < 5; ++q)
              if (q0 + q < NV) { acc[2 * (q0 + q)] -= v[q].x; acc[2 * (q0 + q) + 1] -= v[q].y; }
          }
        }
      }
      if (((i0 + il + 1) % KS_SYNC_EVERY) == 0) __syncthreads();
    }
  }
  if (live) {
    u64 *o = A.out + (size_t)ct * A.out_stride;
    const u64 b = (sl == 0 && A.b_index >= 0) ? a[A.b_word] : 0ull;
#pragma unroll
    for (int q = 0; q < NV; ++q) {
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        const int cc = ch * (64 * NV) + 2 * (lane + 32 * q) + h;
        if (cc >= A.out_words) continue;
        const u64 v = acc[2 * q + h] + (cc == A.b_index ? b : 0ull);
        if (A.slices == 1) o[cc] = v;
        else atomicAdd(&o[cc], v);
      }
    }
  }
}

template <int NV, bool CHUNKED>
static void launch_ks_nv(TableKsArgs A, cudaStream_t st) {
  const int sms = sm_count();
  const int work = A.count * A.chunks;
  // A warp's sweep is one serial chain (3.4 ms at Level 1 whatever the batch), so a batch below one full wave of
  // 28 warps per SM splits every ciphertext's sweep over `slices` warps: up to the full wave from two ciphertexts per SM,
  // about 16 warps per SM below (measured, profiles/r2m_ks_small.log: 256 ciphertexts 0.76 ms at 8 warps per SM, 0.45 at
  // 16, 0.48 at 28 -- more slices only add atomics there)
  int slices = 1;
  if (work < 2 * sms) slices = (16 * sms + work - 1) / work;
  else if (2 * work <= KS_MAX_WARPS * sms) slices = KS_MAX_WARPS * sms / work;
  if (slices > 1) {
    const int max_slices = (A.n_entries + 31) / 32;
    if (slices > max_slices) slices = max_slices;
  }
  A.i_per = (((A.n_entries + slices - 1) / slices) + 31) & ~31;
  A.slices = (A.n_entries + A.i_per - 1) / A.i_per;
  const int vcount = work * A.slices;
  // one CTA per SM when the batch allows; as few waves as possible otherwise.  grid = (CTAs per slice, slices)
  const int waves = (vcount + sms * KS_MAX_WARPS - 1) / (sms * KS_MAX_WARPS);
  int gx = sms * waves / A.slices;
  if (gx < 1) gx = 1;
  int per = (work + gx - 1) / gx;
  if (per < 1) per = 1;
  if (per > KS_MAX_WARPS) per = KS_MAX_WARPS;
  A.per_cta = per;
  const dim3 grid((work + per - 1) / per, A.slices);
  static bool configured_dev[MB_MAX_DEV] = {false};        // function attributes are per device
  bool &configured = configured_dev[current_device()];
  if (!configured) {
    // no shared memory needed: give the whole unified array to L1, which is what shares rows between warps
    MB_CHECK(cudaFuncSetAttribute(keyswitch_warp_kernel<NV, CHUNKED>, cudaFuncAttributePreferredSharedMemoryCarveout, 0));
    configured = true;
  }
  if (A.slices > 1) {
    MB_CHECK(cudaMemset2DAsync(A.out, sizeof(u64) * A.out_stride, 0, sizeof(u64) * A.out_words, A.count, st));
  }
  keyswitch_warp_kernel<NV, CHUNKED><<<grid, per * 32, 0, st>>>(A);
  MB_CHECK(cudaGetLastError());
  count_launch();
}

// table: [n_entries][t][2^base_bit-1][row_stride]; row_stride a multiple of 64 words
void launch_table_keyswitch(const u64 *table, int row_stride, int n_entries, int t, int base_bit, u64 *out,
                            int out_words, int out_stride, int b_index, const u64 *in, int in_stride, int b_word,
                            int count, cudaStream_t st) {
  MB_REQUIRE(base_bit >= 1 && base_bit <= 8, "keyswitch: base_bit=%d unsupported (1..8)", base_bit);
  MB_REQUIRE(t >= 1 && t * base_bit < 64, "keyswitch: t*base_bit must be < 64");
  MB_REQUIRE(row_stride % 64 == 0, "keyswitch: resident rows must be padded to 64 words");
  if (count <= 0) return;
  TableKsArgs A;
  A.out = out; A.in = in; A.table = table; A.count = count; A.n_entries = n_entries; A.in_stride = in_stride;
  A.b_word = b_word; A.out_words = out_words; A.out_stride = out_stride; A.b_index = b_index;
  A.t = t; A.base_bit = base_bit; A.row_stride = row_stride;
  int nv = row_stride / 64;
  A.chunks = 1;
  if (nv > 10 && nv % 8 == 0) {                    // wide (TRLWE) rows: 512-word column chunks, 16 accumulators per lane
    A.chunks = nv / 8;
    nv = 8;
  }
  MB_REQUIRE(nv <= 16, "keyswitch: row stride %d unsupported (above 1024 words it must be a multiple of 512)", row_stride);
  // latency mode: a batch that cannot occupy the GPU even with its sweep sliced over 32-coefficient groups is also
  // split by 64-word output columns (one 128-bit load per lane and row): nv times more warps in flight
  const int max_slices = (n_entries + 31) / 32;
  if ((long long)count * A.chunks * max_slices < 2LL * sm_count()) {
    A.chunks = row_stride / 64;
    launch_ks_nv<1, true>(A, st);
    return;
  }
  if (A.chunks > 1) { launch_ks_nv<8, true>(A, st); return; }
  switch (nv) {
#define MB_KS_CASE(NV_) case NV_: launch_ks_nv<NV_, false>(A, st); break;
    MB_KS_CASE(1) MB_KS_CASE(2) MB_KS_CASE(3) MB_KS_CASE(4) MB_KS_CASE(5) MB_KS_CASE(6) MB_KS_CASE(7) MB_KS_CASE(8)
    MB_KS_CASE(9) MB_KS_CASE(10) MB_KS_CASE(11) MB_KS_CASE(12) MB_KS_CASE(13) MB_KS_CASE(14) MB_KS_CASE(15) MB_KS_CASE(16)
#undef MB_KS_CASE
    default: MB_FATAL("keyswitch: row stride %d unsupported", row_stride);
  }
}

void launch_keyswitch(const KskDev *ksk, u64 *out, const u64 *in, int count, cudaStream_t st) {
  const Params &p = ksk->p;
  MB_REQUIRE(p.n + 1 <= 1024, "keyswitch: output dimension n=%d too large (max 1023)", p.n);
  const int n_in = p.k * p.N;
  launch_table_keyswitch(ksk->d, ksk->row_stride, n_in, p.t, p.base_bit, out, p.n + 1, p.n + 1, p.n, in, n_in + 1, n_in,
                         count, st);
}

}  // namespace mb
