/*
 * dropin_multi.c -- the multi-GPU mode of libmosfhet_b200.so under a C caller of the reference API.
 *
 * Linked against the UNMODIFIED reference library (key generation, encryption, decryption) and run with
 * LD_PRELOAD=libmosfhet_b200.so.  The same batch goes through functional_bootstrap_keyswitch_batch on ONE device and,
 * after mb200_init_multi(ndev), sharded over ndev devices: the two results must be identical word for word (same kernels,
 * same keys, contiguous shards) and decrypt to LUT[m]; the rates of both runs are printed as one JSON line.
 *
 *   dropin_multi <ndev> <batch> [n N l Bg_bit t base_bit]
 */
#define _GNU_SOURCE
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>
#include <unistd.h>

#include "mosfhet_b200.h"

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

static double now_s(void) {
  struct timespec ts;
  clock_gettime(CLOCK_MONOTONIC, &ts);
  return ts.tv_sec + ts.tv_nsec * 1e-9;
}

int main(int argc, char **argv) {
  if (argc < 3) { fprintf(stderr, "usage: dropin_multi <ndev> <batch> [n N l Bg_bit t base_bit]\n"); return 2; }
  const int ndev = atoi(argv[1]), count = atoi(argv[2]);
  const int n = argc > 3 ? atoi(argv[3]) : 632, N = argc > 4 ? atoi(argv[4]) : 1024, k = 1, l = argc > 5 ? atoi(argv[5]) : 3;
  const int Bg_bit = argc > 6 ? atoi(argv[6]) : 6, t = argc > 7 ? atoi(argv[7]) : 7, base_bit = argc > 8 ? atoi(argv[8]) : 2;
  const int torus_base = 4, reps = 3;

  TLWE_Key key_lwe = tlwe_new_binary_key(n, pow(2, -15));
  TLWE_Key key_ext = tlwe_new_binary_key(k * N, pow(2, -25));
  TRLWE_Key key_rlwe = trlwe_new_binary_key(N, k, pow(2, -25));
  trlwe_extract_tlwe_key(key_ext, key_rlwe);
  TRGSW_Key key_gsw = trgsw_new_key(key_rlwe, l, Bg_bit);
  Bootstrap_Key bk = new_bootstrap_key(key_gsw, key_lwe, 1);
  TLWE_KS_Key ksk = tlwe_new_KS_key(key_lwe, key_ext, t, base_bit);
  Torus lut_vals[4];
  for (int m = 0; m < 4; m++) lut_vals[m] = (Torus)((3 * m + 1) % 4) << 61;
  TRLWE lut = trlwe_alloc_new_sample(k, N);
  trlwe_torus_packing(lut, lut_vals, torus_base);

  TLWE *in = malloc(sizeof(TLWE) * count), *out1 = malloc(sizeof(TLWE) * count), *outN = malloc(sizeof(TLWE) * count);
  for (int i = 0; i < count; i++) {
    in[i] = tlwe_new_sample((Torus)(i % torus_base) << 61, key_lwe);
    out1[i] = tlwe_alloc_sample(n);
    outN[i] = tlwe_alloc_sample(n);
  }
  /* one device */
  mb200_register_bootstrap_key(bk);
  mb200_register_ks_key(ksk);
  functional_bootstrap_keyswitch_batch(out1, &lut, 1, in, bk, ksk, torus_base, count);   /* warm-up */
  double t0 = now_s();
  for (int r = 0; r < reps; r++) functional_bootstrap_keyswitch_batch(out1, &lut, 1, in, bk, ksk, torus_base, count);
  const double rate1 = (double)count * reps / (now_s() - t0);
  /* ndev devices */
  const int used = mb200_init_multi(ndev);
  t0 = now_s();
  mb200_register_bootstrap_key(bk);                                                      /* replicates over NVLink */
  mb200_register_ks_key(ksk);
  const double repl_s = now_s() - t0;
  functional_bootstrap_keyswitch_batch(outN, &lut, 1, in, bk, ksk, torus_base, count);   /* warm-up */
  t0 = now_s();
  for (int r = 0; r < reps; r++) functional_bootstrap_keyswitch_batch(outN, &lut, 1, in, bk, ksk, torus_base, count);
  const double rateN = (double)count * reps / (now_s() - t0);

  int differ = 0, wrong = 0;
  for (int i = 0; i < count; i++) {
    if (out1[i]->b != outN[i]->b || memcmp(out1[i]->a, outN[i]->a, sizeof(Torus) * n)) differ++;
    const Torus dec = (tlwe_phase(outN[i], key_lwe) + (1ULL << 60)) >> 61;
    if ((int)(dec & 7) != (3 * (i % torus_base) + 1) % 4) wrong++;
  }
  printf("{\"devices\": %d, \"batch\": %d, \"params\": \"n=%d N=%d l=%d Bg_bit=%d t=%d base_bit=%d\", \"rate_1gpu\": %.1f, "
         "\"rate_multi\": %.1f, \"speedup\": %.3f, \"key_replication_s\": %.3f, \"host_cores\": %ld, \"differ_from_1gpu\": %d, \"wrong\": %d}\n",
         used, count, n, N, l, Bg_bit, t, base_bit, rate1, rateN, rateN / rate1, repl_s, sysconf(_SC_NPROCESSORS_ONLN), differ, wrong);
  printf((differ || wrong) ? "DROPIN MULTI FAILED\n" : "DROPIN MULTI OK\n");
  return (differ || wrong) ? 1 : 0;
}
