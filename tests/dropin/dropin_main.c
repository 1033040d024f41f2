/*
 * dropin_main.c -- the reference as a caller would use it, with libmosfhet_b200.so interposed.
 *
 * Linked against the UNMODIFIED reference library (oracle/_ref/libmosfhet_<variant>.so) for key
 * generation, encryption and decryption, and run with LD_PRELOAD=libmosfhet_b200.so so that the
 * hot-path symbols (functional_bootstrap, tlwe_keyswitch, ...) resolve to the CUDA implementation.
 * A second, private copy of the reference is loaded with dlmopen() into its own namespace (where
 * nothing is interposed) and serves as the oracle on the same in-memory keys and inputs
 * (SURVEY.md 7 step 1).  Prints "DROPIN OK" and exits 0 on parity.
 */
#define _GNU_SOURCE
#include <dlfcn.h>
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "mosfhet_b200.h"

/* reference functions that stay on the CPU (declarations as in mosfhet.h:185-230, 236-262, 316-318, 408) */
typedef struct _TLWE_Key *TLWE_Key;
typedef struct _TRLWE_Key *TRLWE_Key;
typedef struct _TRGSW_Key *TRGSW_Key;
TLWE_Key tlwe_new_binary_key(int n, double sigma);
TRLWE_Key trlwe_new_binary_key(int N, int k, double sigma);
void trlwe_extract_tlwe_key(TLWE_Key out, TRLWE_Key in);
TRGSW_Key trgsw_new_key(TRLWE_Key trlwe_key, int l, int Bg_bit);
Bootstrap_Key new_bootstrap_key(TRGSW_Key out_key, TLWE_Key in_key, int unfolding);
TLWE_KS_Key tlwe_new_KS_key(TLWE_Key out_key, TLWE_Key in_key, int t, int base_bit);
TLWE tlwe_new_sample(Torus m, TLWE_Key key);
TLWE tlwe_alloc_sample(int n);
Torus tlwe_phase(TLWE c, TLWE_Key key);
TRLWE trlwe_alloc_new_sample(int k, int N);
void trlwe_torus_packing(TRLWE out, Torus *in, int size);
Generic_KS_Key trlwe_new_priv_SK_KS_key_N2(TRLWE_Key out_key, TLWE_Key in_key, int t, int base_bit);   /* keyswitch.c:611 */
Generic_KS_Key trlwe_new_packing1_KS_key(TRLWE_Key out_key, TLWE_Key in_key, int t, int base_bit);     /* keyswitch.c:368 */
TRGSW trgsw_alloc_new_sample(int l, int Bg_bit, int k, int N);

typedef void (*fb_t)(TLWE, TRLWE, TLWE, Bootstrap_Key, int);
typedef void (*ks_t)(TLWE, TLWE, TLWE_KS_Key);
typedef void (*gks_t)(TRLWE, TLWE, Generic_KS_Key);
typedef void (*cb_t)(TRGSW, TLWE, Bootstrap_Key, Generic_KS_Key, Generic_KS_Key);

static int trlwe_equal(TRLWE a, TRLWE b, int k, int N) {
  for (int q = 0; q < k; q++)
    if (memcmp(a->a[q]->coeffs, b->a[q]->coeffs, sizeof(Torus) * N)) return 0;
  return !memcmp(a->b->coeffs, b->b->coeffs, sizeof(Torus) * N);
}

static long long sdist(Torus a, Torus b) { long long d = (long long)(a - b); return d < 0 ? -d : d; }

int main(int argc, char **argv) {
  if (argc < 2) { fprintf(stderr, "usage: dropin_main <path to reference .so>\n"); return 2; }
  const int n = 96, N = 1024, k = 1, l = 3, Bg_bit = 6, t = 7, base_bit = 2, torus_base = 4, count = 6;
  void *ref = dlmopen(LM_ID_NEWLM, argv[1], RTLD_NOW | RTLD_LOCAL);
  if (!ref) { fprintf(stderr, "dlmopen: %s\n", dlerror()); return 2; }
  fb_t ref_fb = (fb_t)dlsym(ref, "functional_bootstrap");
  ks_t ref_ks = (ks_t)dlsym(ref, "tlwe_keyswitch");
  /* the interposed symbols must be the CUDA library's */
  Dl_info info;
  dladdr((void *)functional_bootstrap, &info);
  if (!strstr(info.dli_fname, "libmosfhet_b200")) { fprintf(stderr, "functional_bootstrap resolves to %s\n", info.dli_fname); return 3; }

  TLWE_Key key_lwe = tlwe_new_binary_key(n, pow(2, -20));
  TLWE_Key key_ext = tlwe_new_binary_key(k * N, pow(2, -30));
  TRLWE_Key key_rlwe = trlwe_new_binary_key(N, k, pow(2, -30));
  trlwe_extract_tlwe_key(key_ext, key_rlwe);
  TRGSW_Key key_gsw = trgsw_new_key(key_rlwe, l, Bg_bit);
  Bootstrap_Key bk = new_bootstrap_key(key_gsw, key_lwe, 1);
  TLWE_KS_Key ksk = tlwe_new_KS_key(key_lwe, key_ext, t, base_bit);
  Torus lut_vals[4];
  for (int m = 0; m < 4; m++) lut_vals[m] = (Torus)((3 * m + 1) % 4) << 61;
  TRLWE lut = trlwe_alloc_new_sample(k, N);
  trlwe_torus_packing(lut, lut_vals, torus_base);

  int bad = 0;
  for (int i = 0; i < count; i++) {
    const int m = i % torus_base;
    TLWE c = tlwe_new_sample((Torus)m << 61, key_lwe);
    TLWE out = tlwe_alloc_sample(k * N), out_ref = tlwe_alloc_sample(k * N);
    functional_bootstrap(out, lut, c, bk, torus_base);          /* CUDA (interposed) */
    ref_fb(out_ref, lut, c, bk, torus_base);                     /* reference CPU, private namespace */
    const Torus ph = tlwe_phase(out, key_ext), ph_ref = tlwe_phase(out_ref, key_ext);
    /* two FFT implementations agree in phase up to the gadget's rounding step times ~2^7
       (tests/test_gpu_parity.py:phase_tol): 2^(64 - l*Bg_bit + 7) = 2^53 here; messages must be identical */
    if (sdist(ph, ph_ref) > (1LL << (64 - l * Bg_bit + 7))) { printf("PBS %d: phase differs by %lld\n", i, sdist(ph, ph_ref)); bad++; }
    if (((ph + (1ULL << 60)) >> 61) != ((ph_ref + (1ULL << 60)) >> 61)) { printf("PBS %d: message differs\n", i); bad++; }
    TLWE ko = tlwe_alloc_sample(n), ko_ref = tlwe_alloc_sample(n);
    tlwe_keyswitch(ko, out_ref, ksk);                            /* CUDA (interposed) */
    ref_ks(ko_ref, out_ref, ksk);                                /* reference CPU */
    if (ko->b != ko_ref->b || memcmp(ko->a, ko_ref->a, sizeof(Torus) * n)) { printf("KS %d: not bit-exact\n", i); bad++; }
    const Torus dec = (tlwe_phase(ko, key_lwe) + (1ULL << 60)) >> 61;
    if ((int)(dec & 7) != (3 * m + 1) % 4) { printf("KS %d: decrypts to %d, want %d\n", i, (int)(dec & 7), (3 * m + 1) % 4); bad++; }
  }
  /* ---- circuit bootstrap keys as the DEFAULT reference build makes them: seed-compressed TRLWE rows under a
   * process-global AES key (keyswitch.c:231-241, rnd/aes_rng.c:88-93).  The CUDA library expands them at upload
   * through this process's own trlwe_compressed_subto.  The oracle here is the reference function in THIS namespace
   * (dlsym RTLD_NEXT: the definition after the preloaded library) -- the private dlmopen copy has another AES key. */
  {
    const int t2 = 4, bb2 = 2;
    Generic_KS_Key kska = trlwe_new_priv_SK_KS_key_N2(key_rlwe, key_ext, t2, bb2);
    Generic_KS_Key kskb = trlwe_new_packing1_KS_key(key_rlwe, key_ext, t2, bb2);
    gks_t next_priv = (gks_t)dlsym(RTLD_NEXT, "trlwe_priv_keyswitch");
    gks_t next_pack = (gks_t)dlsym(RTLD_NEXT, "trlwe_packing1_keyswitch");
    cb_t next_cb2 = (cb_t)dlsym(RTLD_NEXT, "circuit_bootstrap_2");
    if (!next_priv || !next_pack || !next_cb2) { fprintf(stderr, "RTLD_NEXT lookup failed\n"); return 4; }
    for (int i = 0; i < 3; i++) {
      TLWE c = tlwe_new_sample((Torus)(i % 2) << 62, key_lwe);
      TLWE tl = tlwe_alloc_sample(k * N);
      ref_fb(tl, lut, c, bk, torus_base);                       /* some TLWE of dimension k*N */
      TRLWE a = trlwe_alloc_new_sample(k, N), a_ref = trlwe_alloc_new_sample(k, N);
      trlwe_priv_keyswitch(a, tl, kska);                         /* CUDA (interposed) */
      next_priv(a_ref, tl, kska);                                /* reference CPU, same AES state */
      if (!trlwe_equal(a, a_ref, k, N)) { printf("priv KS %d: not bit-exact\n", i); bad++; }
      trlwe_packing1_keyswitch(a, tl, kskb);
      next_pack(a_ref, tl, kskb);
      if (!trlwe_equal(a, a_ref, k, N)) { printf("packing KS %d: not bit-exact\n", i); bad++; }
      /* the reference's composed caller running on the interposed primitives == the fused CUDA composition */
      TRGSW g = trgsw_alloc_new_sample(l, Bg_bit, k, N), g_ref = trgsw_alloc_new_sample(l, Bg_bit, k, N);
      circuit_bootstrap_2(g, c, bk, kska, kskb);
      next_cb2(g_ref, c, bk, kska, kskb);
      for (int r = 0; r < 2 * l; r++)
        if (!trlwe_equal(g->samples[r], g_ref->samples[r], k, N)) { printf("circuit bootstrap %d row %d differs\n", i, r); bad++; }
    }
  }
  printf(bad ? "DROPIN FAILED (%d)\n" : "DROPIN OK\n", bad);
  return bad ? 1 : 0;
}
