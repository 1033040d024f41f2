// common.cuh -- shared declarations for libmosfhet_b200 (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>

typedef unsigned long long u64;
typedef long long i64;

#define MB_CHECK(expr)                                                                         \
  do {                                                                                         \
    cudaError_t _e = (expr);                                                                   \
    if (_e != cudaSuccess) {                                                                   \
      fprintf(stderr, "mosfhet_b200: CUDA error '%s' at %s:%d (%s) -- no CPU fallback, aborting\n", \
              cudaGetErrorString(_e), __FILE__, __LINE__, #expr);                              \
      abort();                                                                                 \
    }                                                                                          \
  } while (0)

#define MB_FATAL(...)                                                                          \
  do {                                                                                         \
    fprintf(stderr, "mosfhet_b200: " __VA_ARGS__);                                             \
    fprintf(stderr, "\n");                                                                     \
    abort();                                                                                   \
  } while (0)

#define MB_REQUIRE(cond, ...)                                                                  \
  do { if (!(cond)) MB_FATAL(__VA_ARGS__); } while (0)

namespace mb {

// ---- parameters of one resident key / one launch ------------------------------------------
struct Params {
  int n, N, k, l, Bg_bit, t, base_bit;
};

// Resident bootstrapping key: double2 [n][(k+1)*l][(k+1)][M] in the tiled internal slot order
// (see keys.cu: stored index m*(M/8)+c  <->  FFT position 8c+m  <->  frequency bitrev(position)).
struct BskDev {
  Params p;
  double2 *d;
  bool owned;
};

// Resident key-switching table: u64 [N_in][t][2^base_bit-1][row_stride], row_stride = round_up(n+1, 64)
struct KskDev {
  Params p;
  u64 *d;
  int row_stride;
  bool owned;
};

inline int ilog2i(int x) { int r = 0; while ((1 << r) < x) ++r; return r; }
inline int ksk_row_stride(int n_out) { return (n_out + 1 + 63) & ~63; }  // 512-byte rows: 128-bit lane loads, no tail guards

// ---- twiddle tables (tables.cu) -------------------------------------------------------------
// tw[j] = exp(i*pi*j/N), j in [0, N): covers the twist (j < N/2) and every FFT twiddle of the
// N/2-point transform (W_M^t = tw[4t], t < M/2).
const double2 *twiddles_for(int N);
// Per-thread pass tables of the specialised k=1 kernel (blind_rotate_k1.cu).
const double2 *k1_tables_for(int N);

// ---- launch accounting ------------------------------------------------------------------------
void count_launch(int n = 1);

// ---- context ----------------------------------------------------------------------------------
constexpr int MB_MAX_DEV = 16;
void ensure_init();                 // cudaSetDevice(device of this host thread)
cudaStream_t default_stream();      // this host thread's stream on its device
int sm_count();
int current_device();               // device of this host thread (the primary unless bound elsewhere)
int primary_device();
void bind_thread_to_device(int device);
// key of per-device caches: (device, n)
inline long long dev_key(int n) { return ((long long)current_device() << 32) | (long long)(unsigned)n; }

// ---- kernels' host-side launchers ---------------------------------------------------------------
struct BlindRotateLaunch {
  const BskDev *bsk;
  const u64 *tv;        // [(tv_count)][(k+1)*N] initial accumulator(s)
  int tv_count;
  const u64 *in;        // [count][in_stride]: a[0..size) (and b at index `size` when init_rotate)
  int in_stride;
  int in_div;           // 0/1: one input per ciphertext; r > 1: ciphertext ct reads input ct / r (and tv ct % tv_count)
  int size;             // number of blind-rotation steps (= n for a bootstrap)
  int b_index;          // index of b in an input row when init_rotate (0: `size`); set when a launch covers a SEGMENT of the steps
  u64 *out;             // mode 0: [count][(k+1)*N] accumulator; mode 1: [count][k*N+1] TLWE (extract idx 0)
  int extract;          // 0 / 1
  int init_rotate;      // 1: acc = tv * X^(2N - round((b+prec_offset)*2N)) (bootstrap.c:194-195); 0: acc = tv
  u64 prec_offset;
  int preprocess;       // programmable_bootstrap input shaping (bootstrap.c:210-217)
  int kappa, theta;
  int count;
  // single external product mode (trgsw.c:385): out = TRGSW[sel[ct]] (.) in, no rotation, no accumulate
  int direct;
  const int *sel;       // device [count] or nullptr (-> step index)
  double *dft_out;      // when non-null: write the Fourier-domain result [count][(k+1)][N] (Re|Im) in
  const int *dft_perm;  //   host slot order via dft_perm/dft_conj (device [M]) instead of inverting
  const int *dft_conj;
  // CMUX (vertical_packing.c:24-33), direct mode only: operand = tv - sub, result = product + add
  const u64 *sub;       // [count][(k+1)*N] or nullptr
  const u64 *add;       // [count][(k+1)*N] or nullptr (may alias `out`)
  int sel_const;        // >= 0: every ciphertext uses TRGSW number sel_const (no `sel` array)
  // FFT-based TRLWE key switches, direct mode only (the key is held as a TRGSW-shaped row set, l = t, Bg_bit = base_bit):
  //   1: trlwe_keyswitch (keyswitch.c:162-193)        out = (0, in.b) - sum over the mask rows only
  //   2: trlwe_priv_keyswitch_2 (keyswitch.c:52-63)   out = -(rows (.) (in.a, -in.b))
  int ks_mode;
};

void launch_blind_rotate_generic(const BlindRotateLaunch &a, cudaStream_t st);
void launch_unfold(u64 *out, const u64 *su, const u64 *a, int a_stride, int N, int npoly, int unfolding, int group0,
                   int n_groups, int count, cudaStream_t st);
void launch_pb_preprocess(u64 *out, const u64 *in, size_t words, int kappa, int theta, int log_N2, cudaStream_t st);
void launch_fill_sel(int *sel, int E, int size, int bit, int count, cudaStream_t st);
void launch_rotate_trlwe(u64 *out, const u64 *in, int amount, int N, int polys, int count, cudaStream_t st);
void launch_pos_to_host_order(double *out, const double *in, int N, size_t npolys, const int *perm, const int *conj,
                              cudaStream_t st);
bool k1_supported(const Params &p);
bool k1_direct_supported(const BlindRotateLaunch &a);     // one external product / CMUX on the k = 1 kernel
void launch_extprod_k1(const BlindRotateLaunch &a, cudaStream_t st);
void launch_blind_rotate_k1(const BlindRotateLaunch &a, cudaStream_t st);
void k1_variant_name(const Params &p, char *dst, size_t cap);   // variant names are written into the caller's buffer
bool k1h_supported(const Params &p);
bool k1c_supported(const Params &p);                      // one ciphertext per 2-CTA cluster (latency)
void launch_blind_rotate_k1c(const BlindRotateLaunch &a, cudaStream_t st);
void k1c_variant_name(const Params &p, char *dst, size_t cap);
void launch_blind_rotate_k1h(const BlindRotateLaunch &a, cudaStream_t st);
void k1h_variant_name(const Params &p, char *dst, size_t cap);
bool k1q_supported(const Params &p);                      // T = M/4 threads, radix RA x 16 x 4, four warps per scheduler
void launch_blind_rotate_k1q(const BlindRotateLaunch &a, cudaStream_t st);
void k1q_variant_name(const Params &p, char *dst, size_t cap);

void launch_keyswitch(const KskDev *ksk, u64 *out, const u64 *in, int count, cudaStream_t st);
void launch_table_keyswitch(const u64 *table, int row_stride, int n_entries, int t, int base_bit, u64 *out,
                            int out_words, int out_stride, int b_index, const u64 *in, int in_stride, int b_word,
                            int count, cudaStream_t st);
void launch_extract(u64 *out, const u64 *trlwe, const int *d_idx, int idx_count, int N, int k, int count,
                    cudaStream_t st);
void launch_torus_to_dft(double *out, const u64 *in, int N, int count, cudaStream_t st);
void launch_dft_to_torus(u64 *out, const double *in, int N, int count, const int *perm, const int *conj,
                         cudaStream_t st);

// multivalue.cu
void launch_mv_phase1_rotations(u64 *out, const u64 *src, int N, int k, int torus_base, int count, cudaStream_t st);
void launch_mv_phase2(u64 *out, const int *d_lut, int lut_count, const u64 *rot, int N, int k, int torus_base,
                      int log_torus_base, int count, cudaStream_t st);

void launch_mv_extract(u64 *out, int out_stride, const u64 *in, const int *d_ranges, int outs_per_in, int N, int k, int sign,
                       int acc, int count, cudaStream_t st);

// keys.cu
void import_bsk(BskDev *dst, const double *d_host_layout /* device copy of the host-form key */, const int32_t *h_exponents,
                cudaStream_t st);
void synth_bsk(BskDev *dst, const u64 *h_lwe_key, const u64 *h_rlwe_key, double sigma, u64 seed, cudaStream_t st);
void synth_ksk(KskDev *dst, const u64 *h_in_key, const u64 *h_out_key, double sigma, u64 seed, cudaStream_t st);
void bsk_from_torus(BskDev *dst, const u64 *d_torus, double *d_dft, cudaStream_t st);
void synth_gksk(u64 *d_out, const u64 *h_in_key, const u64 *h_out_key, int n_in, int include_b, int N, int t,
                int base_bit, double sigma, u64 seed, cudaStream_t st);
void host_slot_exponents(int layout, int N, int32_t *e);
// host slot h -> (internal stored index, conj flag); used by import and by the DFT-boundary ops
void slot_maps(int N, const int32_t *e, int *stored_to_host /* M */, int *stored_conj /* M */);
int stored_index_of_position(int s, int M);
int position_of_stored_index(int idx, int M);

}  // namespace mb
