/*
 * mosfhet_b200.h -- C ABI of libmosfhet_b200.so
 *
 * B200 (sm_100a) implementation of MOSFHET's programmable-bootstrap hot path, exported
 *   (1) under the reference's own symbol names and signatures, so that the library is a drop-in
 *       for that path when it is linked / LD_PRELOADed ahead of libmosfhet.so, and
 *   (2) as new batched entry points over arrays of ciphertext handles, and
 *   (3) as a flat (plain pointers + sizes) API for device-resident pipelines and other-language
 *       bindings (ctypes, cgo, JNI).
 *
 * There is NO CPU fallback: every compute entry point aborts with a message on stderr if no
 * CUDA device / kernel image is available (the reference's own error style: assert()/exit(),
 * mosfhet.h has no return codes -- see /root/reference/src/misc.c:104-128).
 *
 * Reference interface citations are "mosfhet.h:<line>" = /root/reference/include/mosfhet.h.
 */
#ifndef MOSFHET_B200_H
#define MOSFHET_B200_H

#include <stdint.h>
#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

/* ------------------------------------------------------------------------------------------
 * Ciphertext / key handle types.  A caller that already includes the reference's <mosfhet.h>
 * defines MOSFHET_B200_USE_REFERENCE_TYPES before including this file; otherwise the
 * ABI-identical definitions below are used (same member order, widths and pointer depth as
 * mosfhet.h:22-133; 64-bit torus only -- the TORUS32 build switch of mosfhet.h:23-28 is not
 * supported by the GPU path).
 * ---------------------------------------------------------------------------------------- */
#ifndef MOSFHET_B200_USE_REFERENCE_TYPES

typedef uint64_t Torus;                                                      /* mosfhet.h:27  */

typedef struct _TorusPolynomial { Torus  *coeffs; int N; } *TorusPolynomial; /* mosfhet.h:32  */
typedef struct _DFT_Polynomial  { double *coeffs; int N; } *DFT_Polynomial;  /* mosfhet.h:37  */

typedef struct _TLWE { Torus *a, b; int n; } *TLWE;                          /* mosfhet.h:51  */

typedef struct _TLWE_KS_Key {                                                /* mosfhet.h:62  */
  TLWE ***s;              /* s[i][j][d-1], i < n (input dim), j < t, d in 1..2^base_bit-1      */
  int base_bit, t, n;
} *TLWE_KS_Key;

typedef struct _TRLWE     { TorusPolynomial *a, b; int k; } *TRLWE;          /* mosfhet.h:73  */
typedef struct _TRLWE_DFT { DFT_Polynomial  *a, b; int k; } *TRLWE_DFT;      /* mosfhet.h:78  */

typedef struct _TRGSW     { TRLWE     *samples; int l, Bg_bit; } *TRGSW;     /* mosfhet.h:106 */
typedef struct _TRGSW_DFT { TRLWE_DFT *samples; int l, Bg_bit; } *TRGSW_DFT; /* mosfhet.h:111 */

typedef struct _Generic_KS_Key {                                             /* mosfhet.h:100 */
  TRLWE ***s;             /* s[i][j][d-1], i < n + include_b: TRLWE rows (possibly seed-compressed) */
  int base_bit, t, n, include_b;
} *Generic_KS_Key;

typedef struct _TRLWE_KS_Key {                                               /* mosfhet.h:90 */
  TRLWE_DFT **s;          /* s[i][j], i < k (input mask polynomials), j < t: Fourier-domain TRLWE rows */
  int base_bit, t, k;
} *TRLWE_KS_Key;

typedef struct _Bootstrap_Key {                                              /* mosfhet.h:129 */
  TRGSW_DFT *s;           /* s[i], i < n: Fourier-domain TRGSW of LWE key bit i (unfolding==1) */
  TRGSW     *su;          /* su[g*2^u + j]: torus-domain keys for unfolding u > 1 (bootstrap.c:23-48) */
  int n, k, N, Bg_bit, l, unfolding;
} *Bootstrap_Key;

#endif /* MOSFHET_B200_USE_REFERENCE_TYPES */

/* ------------------------------------------------------------------------------------------
 * (1) Drop-in entry points: same names, argument meaning, ownership and error behaviour as
 *     the reference.  Outputs are caller-allocated; keys are read-only; blind_rotate mutates
 *     `tv` in place (bootstrap.c:118).  Keys are uploaded to HBM on first use and cached by
 *     host pointer (see mb200_register_* below to do it eagerly).
 * ---------------------------------------------------------------------------------------- */
void functional_bootstrap(TLWE out, TRLWE tv, TLWE in, Bootstrap_Key key, int torus_base);            /* mosfhet.h:412, bootstrap.c:200 */
void functional_bootstrap_wo_extract(TRLWE out, TRLWE tv, TLWE in, Bootstrap_Key key, int torus_base); /* mosfhet.h:411, bootstrap.c:192 */
void programmable_bootstrap(TLWE out, TRLWE tv, TLWE in, Bootstrap_Key key,
                            int precision, int kappa, int theta);                                     /* mosfhet.h:425, bootstrap.c:208 */
void blind_rotate(TRLWE tv, Torus *a, TRGSW_DFT *s, int size);                                        /* mosfhet.h:409, bootstrap.c:107 */
void trgsw_mul_trlwe_DFT(TRLWE_DFT out, TRLWE in1, TRGSW_DFT in2);                                    /* mosfhet.h:344, trgsw.c:385     */
void trlwe_from_DFT(TRLWE out, TRLWE_DFT in);                                                         /* mosfhet.h:263, trlwe.c:629     */
void trlwe_extract_tlwe(TLWE out, TRLWE in, int idx);                                                 /* mosfhet.h:277, trlwe.c:540     */
void tlwe_keyswitch(TLWE out, TLWE in, TLWE_KS_Key ks_key);                                           /* mosfhet.h:227, tlwe.c:289      */
void multivalue_bootstrap_CLOT21(TLWE *out, TRLWE tv, TLWE in, Bootstrap_Key key,
                                 int torus_base, int n_luts);                                         /* mosfhet.h:424, bootstrap.c:222 */
/* circuit bootstrap: TLWE -> TRGSW through one packed-LUT blind rotation, l
 * extractions and 2*l table key switches over TRLWE rows */
void trlwe_packing1_keyswitch(TRLWE out, TLWE in, Generic_KS_Key ks_key);                             /* mosfhet.h:384, keyswitch.c:458 */
void trlwe_priv_keyswitch(TRLWE out, TLWE in, Generic_KS_Key ks_key);                                 /* mosfhet.h:386, keyswitch.c:639 */
void circuit_bootstrap_2(TRGSW out, TLWE in, Bootstrap_Key key, Generic_KS_Key kska, Generic_KS_Key kskb); /* mosfhet.h:421, bootstrap.c:324 */
void circuit_bootstrap(TRGSW out, TLWE in, Bootstrap_Key key, Generic_KS_Key kska, Generic_KS_Key kskb);   /* mosfhet.h:420, bootstrap.c:309 */
void circuit_bootstrap_3(TRGSW out, TLWE in, Bootstrap_Key key, TRLWE_KS_Key *kska, Generic_KS_Key kskb);  /* mosfhet.h:422, bootstrap.c:347 */
void trlwe_keyswitch(TRLWE out, TRLWE in, TRLWE_KS_Key ks_key);                                           /* mosfhet.h:375, keyswitch.c:162 (out may alias in) */
void trlwe_priv_keyswitch_2(TRLWE out, TRLWE in, TRLWE_KS_Key *ks_key);                                   /* mosfhet.h:399, keyswitch.c:52 */
/* TRGSW-accumulator bootstrap and unfolded blind rotation.  functional_bootstrap and
 * functional_bootstrap_wo_extract also accept unfolding > 1 keys (bootstrap.c:197). */
void functional_bootstrap_trgsw_phase1(TRGSW_DFT out, TLWE in, Bootstrap_Key key, int torus_base);        /* mosfhet.h:415, bootstrap.c:286 */
void functional_bootstrap_trgsw_phase2(TLWE out, TRGSW_DFT in, TRLWE tv);                                 /* mosfhet.h:416, bootstrap.c:298 */
void blind_rotate_unfolded(TRLWE tv, Torus *a, TRGSW *s, int size, int unfolding);                        /* mosfhet.h:410, bootstrap.c:124 */
void multivalue_bootstrap_UBR_phase1(TRGSW_DFT *out, TLWE in, Bootstrap_Key key);                         /* mosfhet.h:431, bootstrap.c:151 */
void multivalue_bootstrap_UBR_phase2(TLWE out, TRLWE tv, TLWE in, TRGSW_DFT *sa, Bootstrap_Key key,
                                     int torus_base);                                                     /* mosfhet.h:432, bootstrap.c:174 */
/* extraction family of the multi-ciphertext caller (applications/multi-ciphertext-arith/src/integer.c:94-100): integer, bit-exact */
void trlwe_extract_tlwe_addto(TLWE out, TRLWE in, int idx);                                               /* mosfhet.h:297, trlwe.c:554 */
void trlwe_extract_tlwe_subto(TLWE out, TRLWE in, int idx);                                               /* mosfhet.h:298, trlwe.c:567 */
void trlwe_mv_extract_tlwe(TLWE *out, TRLWE in, int amount);                                              /* mosfhet.h:279, trlwe.c:580 */
void trlwe_mv_extract_tlwe_scaling(TLWE out, TRLWE in, int scale);                                        /* mosfhet.h:280, trlwe.c:591 */
void trlwe_mv_extract_tlwe_scaling_addto(TLWE out, TRLWE in, int scale);                                  /* mosfhet.h:281, trlwe.c:602 */
void trlwe_mv_extract_tlwe_scaling_subto(TLWE out, TRLWE in, int scale);                                  /* mosfhet.h:282, trlwe.c:612 */
void multivalue_bootstrap_phase1(TRLWE *out, TLWE in, Bootstrap_Key key, int torus_base);             /* mosfhet.h:413, bootstrap.c:232 */
void multivalue_bootstrap_phase2(TLWE out, int *in, TRLWE *rotated_tv, int torus_base,
                                 int log_torus_base);                                                 /* mosfhet.h:414, bootstrap.c:245 */

/* The reference's key destructors, interposed: they drop the resident copy of the key (malloc hands freed addresses to
 * the next key) and then run the reference's own definition, found with dlsym(RTLD_NEXT).  Resident copies are also
 * guarded by a fingerprint of the host tree, so a key regenerated in place is uploaded again. */
void free_bootstrap_key(Bootstrap_Key key);                                                           /* mosfhet.h:417, bootstrap.c:51 */
void free_tlwe_ks_key(TLWE_KS_Key key);                                                               /* mosfhet.h:231, tlwe.c:232 */
void free_trlwe_generic_ks_key(Generic_KS_Key key);                                                   /* mosfhet.h:388, keyswitch.c:393 */
void free_trlwe_ks_key(TRLWE_KS_Key key);                                                             /* mosfhet.h:376, keyswitch.c:109 */

/* ------------------------------------------------------------------------------------------
 * (2) Batched variants over arrays of handles (new).  `count` independent ciphertexts per call.
 *     tv_count is 1 (one test vector shared by the whole batch) or `count` (one per input).
 * ---------------------------------------------------------------------------------------- */
void functional_bootstrap_batch(TLWE *out, TRLWE *tv, int tv_count, TLWE *in,
                                Bootstrap_Key key, int torus_base, int count);
void functional_bootstrap_wo_extract_batch(TRLWE *out, TRLWE *tv, int tv_count, TLWE *in,
                                           Bootstrap_Key key, int torus_base, int count);
void programmable_bootstrap_batch(TLWE *out, TRLWE *tv, int tv_count, TLWE *in, Bootstrap_Key key,
                                  int precision, int kappa, int theta, int count);
void blind_rotate_batch(TRLWE *tv, Torus **a, TRGSW_DFT *s, int size, int count);
void trgsw_mul_trlwe_DFT_batch(TRLWE_DFT *out, TRLWE *in1, TRGSW_DFT *in2, int in2_count, int count);
void trlwe_from_DFT_batch(TRLWE *out, TRLWE_DFT *in, int count);
/* CMUX over arrays with one selector: out[i] = in1[i] + selector (.) (in2[i] - in1[i]); out[i] may be in1[i]
 * (applications/leveled_lut/vertical_packing.c:24-33) */
void trgsw_cmux_batch(TRLWE *out, TRLWE *in1, TRLWE *in2, TRGSW_DFT selector, int count);
void trlwe_extract_tlwe_batch(TLWE *out, TRLWE *in, const int *idx, int idx_count, int count);
void tlwe_keyswitch_batch(TLWE *out, TLWE *in, TLWE_KS_Key ks_key, int count);
/* The BASELINE metric's unit of work: functional_bootstrap followed by tlwe_keyswitch
 * (caller pattern of applications/multi-ciphertext-arith/src/integer.c:94-96); the
 * intermediate dimension-k*N TLWE never leaves HBM. out[i]->n == ks_key out dimension. */
void functional_bootstrap_keyswitch_batch(TLWE *out, TRLWE *tv, int tv_count, TLWE *in,
                                          Bootstrap_Key key, TLWE_KS_Key ks_key,
                                          int torus_base, int count);
/* The extraction family over arrays: out[i] (op)= f(in[i]).  `mode`: 0 = plain (out is overwritten; scaling only),
 * +1 = addto, -1 = subto.  trlwe_mv_extract_tlwe_batch: out[i] points to `amount` TLWEs. */
void trlwe_extract_tlwe_acc_batch(TLWE *out, TRLWE *in, const int *idx, int idx_count /* 1 or count */, int mode, int count);
void trlwe_mv_extract_tlwe_batch(TLWE **out, TRLWE *in, int amount, int count);
void trlwe_mv_extract_tlwe_scaling_batch(TLWE *out, TRLWE *in, int scale, int mode, int count);
/* One digit step of the multi-ciphertext arithmetic (integer.c:94-100) for `count` independent digits, nothing but the
 * digits crossing PCIe:   tmp = tlwe_keyswitch(digit[i]);  acc = functional_bootstrap_wo_extract(tv, tmp, torus_base);
 *   digit[i] -= trlwe_mv_extract_tlwe_scaling(acc, scale_digit)        (trlwe_mv_extract_tlwe_scaling_subto)
 *   carry[i] += trlwe_mv_extract_tlwe_scaling(acc, scale_carry)        (trlwe_mv_extract_tlwe_scaling_addto; carry may be NULL)
 * digit[i] and carry[i] have dimension k*N; the constant adjustments of b (integer.c:97, 99) stay with the caller. */
void tlwe_keyswitch_bootstrap_mv_extract_batch(TLWE *digit, TLWE *carry, TRLWE *tv, int tv_count, TLWE_KS_Key ks_key,
                                               Bootstrap_Key key, int torus_base, int scale_digit, int scale_carry, int count);
void multivalue_bootstrap_CLOT21_batch(TLWE **out, TRLWE *tv, int tv_count, TLWE *in,
                                       Bootstrap_Key key, int torus_base, int n_luts, int count);
/* out[c] points to torus_base+1 TRLWEs; lut[c] (lut_count == count) or lut[0] (lut_count == 1) holds
 * torus_base integers (the cleartext LUT of bootstrap.c:245). */
void trlwe_packing1_keyswitch_batch(TRLWE *out, TLWE *in, Generic_KS_Key ks_key, int count);
void trlwe_priv_keyswitch_batch(TRLWE *out, TLWE *in, Generic_KS_Key ks_key, int count);
void circuit_bootstrap_2_batch(TRGSW *out, TLWE *in, Bootstrap_Key key, Generic_KS_Key kska, Generic_KS_Key kskb, int count);
void circuit_bootstrap_batch(TRGSW *out, TLWE *in, Bootstrap_Key key, Generic_KS_Key kska, Generic_KS_Key kskb, int count);
void circuit_bootstrap_3_batch(TRGSW *out, TLWE *in, Bootstrap_Key key, TRLWE_KS_Key *kska, Generic_KS_Key kskb, int count);
void trlwe_keyswitch_batch(TRLWE *out, TRLWE *in, TRLWE_KS_Key ks_key, int count);
void trlwe_priv_keyswitch_2_batch(TRLWE *out, TRLWE *in, TRLWE_KS_Key *ks_key, int count);
void functional_bootstrap_trgsw_phase1_batch(TRGSW_DFT *out, TLWE *in, Bootstrap_Key key, int torus_base, int count);
void functional_bootstrap_trgsw_phase2_batch(TLWE *out, TRGSW_DFT *in, TRLWE *tv, int tv_count, int count);
void blind_rotate_unfolded_batch(TRLWE *tv, Torus **a, TRGSW *s, int size, int unfolding, int count);
void multivalue_bootstrap_phase1_batch(TRLWE **out, TLWE *in, Bootstrap_Key key, int torus_base, int count);
void multivalue_bootstrap_phase2_batch(TLWE *out, int **lut, int lut_count, TRLWE **rotated_tv,
                                       int torus_base, int log_torus_base, int count);

/* ------------------------------------------------------------------------------------------
 * Runtime, key residency and the host Fourier slot order.
 * ---------------------------------------------------------------------------------------- */
int  mb200_init(int device);            /* select device, create context; <0 only if device==-2 probe and none present */
void mb200_shutdown(void);              /* free cached keys, staging buffers, streams */
int  mb200_device_count(void);          /* number of visible CUDA devices (0 without a GPU; no abort) */
/* Multi-GPU inside the library: after mb200_init_multi(ndev) the batched drop-in entry points
 * (functional_bootstrap[_wo_extract|_keyswitch]_batch, programmable_bootstrap_batch, tlwe_keyswitch_batch) cut their batch
 * into contiguous shards over devices 0..ndev-1 (ndev <= 0: all), one host worker thread, stream and staging area per
 * device; registered keys are replicated to every device over NVLink (cudaMemcpyPeer).  Returns the devices in use. */
int  mb200_init_multi(int ndev);
int  mb200_multi_device_count(void);
const char *mb200_version(void);
void mb200_device_synchronize(void);

/* The reference stores Fourier-domain polynomials in the slot order of whichever CPU FFT backend
 * it was compiled with (polynomial.c:359-375): slot h of a DFT_Polynomial holds p(w^e_h),
 * w = exp(i*pi/N).  The GPU library needs e_h to import keys and to return TRLWE_DFT results.
 *   MB200_FFT_AUTO    : probe the host's polynomial_torus_to_DFT (dlsym) with the monomial X
 *   MB200_FFT_SPQLIOS : e_h = 1 + 4*bitrev(h)   (src/fft/spqlios)
 *   MB200_FFT_FFNT    : e_h = 1 - 4*bitrev(h)   (src/fft/ffnt)
 *   MB200_FFT_NATURAL : e_h = 1 + 4*h
 * Slots hold Re in coeffs[h] and Im in coeffs[h + N/2] (polynomial.c:396-399). */
enum { MB200_FFT_AUTO = 0, MB200_FFT_SPQLIOS = 1, MB200_FFT_FFNT = 2, MB200_FFT_NATURAL = 3 };
void mb200_set_host_fft_layout(int layout);
int  mb200_get_host_fft_layout(void);
void mb200_host_slot_exponents(int layout, int N, int32_t *e_out /* N/2 */);

void mb200_register_bootstrap_key(Bootstrap_Key key);   /* upload now (otherwise lazily on first use) */
void mb200_release_bootstrap_key(Bootstrap_Key key);
void mb200_register_ks_key(TLWE_KS_Key key);
void mb200_release_ks_key(TLWE_KS_Key key);
/* Seed-compressed rows (the reference's default A_PRNG=vaes build, keyswitch.c:231-241) are expanded at
 * upload by calling the host's own trlwe_compressed_subto through dlsym; the AES key is process-global in
 * the reference (rnd/aes_rng.c:88-93), so this must happen in the process that generated the key. */
void mb200_register_generic_ks_key(Generic_KS_Key key);
void mb200_release_generic_ks_key(Generic_KS_Key key);
void mb200_register_trlwe_ks_key(TRLWE_KS_Key key);              /* keyed by key->s */
void mb200_release_trlwe_ks_key(TRLWE_KS_Key key);
void mb200_register_trlwe_priv_ks_key(TRLWE_KS_Key *keys);       /* the TRLWE_KS_Key[2] of trlwe_new_priv_KS_key, keyed by the array */
void mb200_release_trlwe_priv_ks_key(TRLWE_KS_Key *keys);

/* ------------------------------------------------------------------------------------------
 * (3) Flat API.  Layouts (all little-endian u64 / f64, contiguous):
 *     TLWE   of dimension n : n+1 words   = a[0..n) then b
 *     TRLWE  (k, N)         : (k+1)*N     = a[0] .. a[k-1] then b, N coefficients each
 *     BSK, host form        : [n][(k+1)*l rows][(k+1) polys a[0..k),b][N doubles = Re(0..N/2) | Im(0..N/2)]
 *                             in the slot order given by `layout`
 *     KSK, host form        : [N_in][t][2^base_bit - 1][n_out + 1]
 *     Pointers named h_* are host, d_* are device (cudaMalloc'd / torch CUDA tensors).
 *     `stream` is a cudaStream_t passed as void* (NULL = the library's per-thread stream).
 * ---------------------------------------------------------------------------------------- */
typedef struct mb200_params {
  int n;         /* LWE dimension (blind-rotation length)                       */
  int N, k;      /* ring degree and TRLWE mask polynomials                      */
  int l, Bg_bit; /* gadget levels / log2 base of the bootstrapping key          */
  int t, base_bit; /* key-switch levels / log2 base                             */
} mb200_params;

typedef struct mb200_bsk *mb200_bsk_t;   /* resident Fourier-domain bootstrapping key */
typedef struct mb200_ksk *mb200_ksk_t;   /* resident key-switching table              */

size_t      mb200_bsk_device_bytes(const mb200_params *p);
size_t      mb200_ksk_device_bytes(const mb200_params *p);
mb200_bsk_t mb200_bsk_from_host(const mb200_params *p, const double *h_bsk, int layout);
mb200_ksk_t mb200_ksk_from_host(const mb200_params *p, const uint64_t *h_ksk);
/* Adopt caller-owned device memory already in the resident layout (e.g. the receive buffer of an
 * NCCL broadcast from rank 0, see mosfhet_b200/sharding.py); the library does not free it. */
mb200_bsk_t mb200_bsk_adopt_device(const mb200_params *p, void *d_bsk);
mb200_ksk_t mb200_ksk_adopt_device(const mb200_params *p, void *d_ksk);
void       *mb200_bsk_device_ptr(mb200_bsk_t bsk);
void       *mb200_ksk_device_ptr(mb200_ksk_t ksk);
void        mb200_bsk_free(mb200_bsk_t bsk);
void        mb200_ksk_free(mb200_ksk_t ksk);

/* Synthetic keys generated on the device from caller-supplied binary secrets (bench / test
 * support; replaces the host-side key generators bootstrap.c:3-21 and tlwe.c:193-212, which stay
 * with the reference CPU build).  lwe_key: n words in {0,1}; rlwe_key: k*N words in {0,1}. */
mb200_bsk_t mb200_bsk_synthesize(const mb200_params *p, const uint64_t *h_lwe_key,
                                 const uint64_t *h_rlwe_key, double rlwe_sigma, uint64_t seed);
mb200_ksk_t mb200_ksk_synthesize(const mb200_params *p, const uint64_t *h_rlwe_key /* k*N, input key  */,
                                 const uint64_t *h_lwe_key  /* n, output key */,
                                 double lwe_sigma, uint64_t seed);

/* trgsw_to_DFT (trgsw.c:345) for a set of p->n torus-domain TRGSW samples already on the device
 * ([n][(k+1)l][(k+1)][N] words, e.g. the output of mb200_circuit_bootstrap_dev): returns a resident set usable
 * with mb200_extprod_dev / mb200_cmux_dev / mb200_vertical_packing_dev / mb200_blind_rotate_dev. */
mb200_bsk_t mb200_bsk_from_torus_dev(const mb200_params *p, const uint64_t *d_trgsw, void *stream);

/* TRLWE-row key-switching keys (k = 1): host form [n_in + include_b][t][2^base_bit-1][2][N]; synthetic keys
 * are generated on the device from binary secrets (bench / test support, replaces keyswitch.c:368-390, 611-637) */
typedef struct mb200_gksk *mb200_gksk_t;
mb200_gksk_t mb200_gksk_from_host(const uint64_t *h_rows, int n_in, int include_b, int N, int t, int base_bit);
mb200_gksk_t mb200_gksk_synthesize(const uint64_t *h_in_key /* n_in */, const uint64_t *h_out_rlwe_key /* N */,
                                   int n_in, int include_b, int N, int t, int base_bit, double sigma, uint64_t seed);
void        *mb200_gksk_device_ptr(mb200_gksk_t k);   /* [n_in + include_b][t][2^base_bit-1][2][N] words */
void         mb200_gksk_free(mb200_gksk_t k);
/* trlwe_packing1_keyswitch (include_b = 0 keys) / trlwe_priv_keyswitch (include_b = 1 keys) on device buffers */
void mb200_trlwe_ks_dev(mb200_gksk_t ksk, uint64_t *d_out_trlwe /* [count][2N] */, const uint64_t *d_in_tlwe /* [count][n_in+1] */,
                        int count, void *stream);
/* circuit_bootstrap_2 on device buffers: d_out_trgsw [count][2l][2][N] (rows 0..l-1 private, l..2l-1 packing) */
void mb200_circuit_bootstrap_dev(mb200_bsk_t bsk, mb200_gksk_t kska, mb200_gksk_t kskb, uint64_t *d_out_trgsw,
                                 const uint64_t *d_in /* [count][n+1] */, int Bg_bit_out, int count, void *stream);

/* FFT-based TRLWE key switches on device buffers.  `row_set` is a TRGSW-shaped resident set with n = 1, l = t,
 * Bg_bit = base_bit (mb200_bsk_from_host): rows [i*t + j] = TRLWE_KS_Key->s[i][j]; the k*t.. rows are zero for
 * mode 1 (trlwe_keyswitch, keyswitch.c:162) or the second key of the pair for mode 2 (trlwe_priv_keyswitch_2, :52). */
void mb200_trlwe_fft_ks_dev(mb200_bsk_t row_set, int mode, uint64_t *d_out /* [count][(k+1)N], may alias d_in */,
                            const uint64_t *d_in, int count, void *stream);
/* circuit_bootstrap (variant 1, bootstrap.c:309), _2 (variant 2, :324) or _3 (variant 3, :347; private key switch =
 * `kska_fft` row set of the TRLWE_KS_Key pair, `kska` unused) on device buffers; d_out_trgsw [count][2*l_out][2][N] */
void mb200_circuit_bootstrap_variant_dev(int variant, mb200_bsk_t bsk, mb200_gksk_t kska, mb200_bsk_t kska_fft,
                                         mb200_gksk_t kskb, uint64_t *d_out_trgsw, const uint64_t *d_in, int l_out,
                                         int Bg_bit_out, int count, void *stream);

/* functional_bootstrap_trgsw_phase1 without the final trgsw_to_DFT: d_out_trgsw [count][(k+1)*l_out][(k+1)N] torus
 * rows of TRGSW(X^-phase); feed mb200_bsk_from_torus_dev + mb200_extprod_dev + mb200_extract_dev for phase 2. */
void mb200_bootstrap_trgsw_phase1_dev(mb200_bsk_t bsk, uint64_t *d_out_trgsw, const uint64_t *d_in /* [count][n+1] */,
                                      int l_out, int Bg_bit_out, int torus_base, int count, void *stream);

/* Device-resident batch ops (inputs/outputs already in HBM; asynchronous on `stream`). */
void mb200_pbs_dev(mb200_bsk_t bsk, uint64_t *d_out_tlwe /* [count][k*N+1] */,
                   const uint64_t *d_tv /* [tv_count][(k+1)*N] */, int tv_count,
                   const uint64_t *d_in /* [count][n+1] */, int torus_base, int count, void *stream);
void mb200_pbs_wo_extract_dev(mb200_bsk_t bsk, uint64_t *d_out_trlwe /* [count][(k+1)*N] */,
                              const uint64_t *d_tv, int tv_count, const uint64_t *d_in,
                              int torus_base, int count, void *stream);
void mb200_blind_rotate_dev(mb200_bsk_t bsk, uint64_t *d_acc /* [count][(k+1)*N], in place */,
                            const uint64_t *d_a /* [count][a_stride] */, int a_stride, int size,
                            int count, void *stream);
void mb200_extract_dev(uint64_t *d_out_tlwe /* [count*idx_count][k*N+1] */, const uint64_t *d_trlwe,
                       const int *h_idx, int idx_count, int N, int k, int count, void *stream);
void mb200_ks_dev(mb200_ksk_t ksk, uint64_t *d_out /* [count][n+1] */,
                  const uint64_t *d_in /* [count][k*N+1] */, int count, void *stream);
void mb200_pbs_ks_dev(mb200_bsk_t bsk, mb200_ksk_t ksk, uint64_t *d_out /* [count][n+1] */,
                      const uint64_t *d_tv, int tv_count, const uint64_t *d_in,
                      uint64_t *d_scratch /* [count][k*N+1] */, int torus_base, int count, void *stream);
/* One external product + inverse transform per ciphertext: out = TRGSW_i (.) in (trgsw.c:385 + trlwe.c:629) */
void mb200_extprod_dev(mb200_bsk_t trgsw_set, const int *h_sel /* [count] index into the set */,
                       uint64_t *d_out_trlwe, const uint64_t *d_in_trlwe, int count, void *stream);
/* CMUX on device-resident TRLWEs, selector = sample `sel` of a resident TRGSW set; d_out may alias d_in1 */
void mb200_cmux_dev(mb200_bsk_t trgsw_set, int sel, uint64_t *d_out, const uint64_t *d_in1, const uint64_t *d_in2,
                    int count, void *stream);
/* CGGI vertical packing (vertical_packing.c:36-52): bits = TRGSW(bit i), i < size; d_luts holds
 * 2^(size - log2 N) TRLWE LUTs (consumed); d_out_tlwe receives the TLWE (dimension k*N) of LUT[input]. */
void mb200_vertical_packing_dev(mb200_bsk_t bits, uint64_t *d_luts, uint64_t *d_out_tlwe, int size, void *stream);
/* E independent evaluations at once (BASELINE config 5: leveled LUT over batched ciphertexts): evaluation e's TRGSW(bit i)
 * is sample e*size + i of `bits`; d_luts is LUT-major [2^(size - log2 N)][E][(k+1)N] (consumed); d_out_tlwe [E][k*N+1] */
void mb200_vertical_packing_batch_dev(mb200_bsk_t bits, uint64_t *d_luts, uint64_t *d_out_tlwe, int size, int E, void *stream);
/* Negacyclic transforms in the library's internal slot order (bit-reversed, e = 1+4*bitrev(s)). */
void mb200_torus_to_dft_dev(double *d_out /* [count][N] Re|Im */, const uint64_t *d_in, int N, int count, void *stream);
void mb200_dft_to_torus_dev(uint64_t *d_out, const double *d_in, int N, int count, void *stream);

/* Host-buffer batch ops: H2D of inputs, kernels, D2H of results, synchronous on return.  From three GPU waves of ciphertexts
 * up, mb200_pbs_ks_host overlaps the copies with the kernels on two streams (first wave launched at once, key switch in four
 * slices whose results leave while the next is switched) -- effective when the host buffers are pinned (cudaHostAlloc /
 * cudaHostRegister); with pageable memory the copies are staged by the driver and the call is merely correct. */
void mb200_pbs_ks_host(mb200_bsk_t bsk, mb200_ksk_t ksk, uint64_t *h_out /* [count][n+1] */,
                       const uint64_t *h_tv, int tv_count, const uint64_t *h_in /* [count][n+1] */,
                       int torus_base, int count);
void mb200_pbs_host(mb200_bsk_t bsk, uint64_t *h_out /* [count][k*N+1] */, const uint64_t *h_tv,
                    int tv_count, const uint64_t *h_in, int torus_base, int count);
void mb200_ks_host(mb200_ksk_t ksk, uint64_t *h_out, const uint64_t *h_in, int count);

/* Introspection for tests / bench: kernels launched by this library since the last reset, and
 * which blind-rotation kernel variant the last PBS call dispatched to. */
uint64_t    mb200_launch_count(void);
void        mb200_reset_launch_count(void);
const char *mb200_last_blind_rotate_kernel(void);
/* Kernel choice for the blind rotation: 0 = auto (by batch size and shape), 1 = generic (any k, l, N), 2 = k1 (one CTA
 * per ciphertext, throughput), 3 = k1h (T = M/4 threads), 4 = k1c (one ciphertext per 2-CTA cluster, latency) */
void        mb200_set_kernel_policy(int policy);
/* Measured FP64 FMA throughput (TFLOP/s) of the current device: the roofline denominator of the
 * FP64-bound blind-rotation kernel (MEASURED_PEAKS.json has no FP64 entry). */
double      mb200_measure_fp64_tflops(int iters);

#ifdef __cplusplus
}
#endif
#endif /* MOSFHET_B200_H */
