/*
 * ref_bench.c -- TEST / BASELINE INFRASTRUCTURE ONLY.  Times the UNMODIFIED reference CPU
 * implementation (linked from oracle/_ref/libmosfhet_<variant>.so) on the metric's unit of work:
 * one functional_bootstrap followed by one tlwe_keyswitch per ciphertext, on `threads` host
 * threads sharing one read-only key set (SURVEY.md 8(d) "CPU baseline, same run").
 * Keys and inputs are generated exactly as test/benchmark.c:97-114 does.
 *
 *   ref_bench n N k l Bg_bit t base_bit lwe_sigma rlwe_sigma threads ops_per_thread steps warmup
 *
 * Prints one JSON object on stdout.
 */
#define _GNU_SOURCE
#include <mosfhet.h>
#include <pthread.h>
#include <time.h>

static int n, N, k, l, Bg_bit, t, base_bit, threads, ops, steps, warmup;
static Bootstrap_Key bk;
static TLWE_KS_Key ksk;
static TRLWE lut;
static TLWE *c_in, *c_mid, *c_out;
static pthread_barrier_t bar;

static double now_ms(void) {
  struct timespec ts;
  clock_gettime(CLOCK_MONOTONIC, &ts);
  return ts.tv_sec * 1e3 + ts.tv_nsec * 1e-6;
}

static void *worker(void *arg) {
  const long id = (long)arg;
  init_fft(N);
  for (int s = 0; s < warmup + steps; s++) {
    pthread_barrier_wait(&bar);
    for (int i = 0; i < ops; i++) {
      const long idx = id * ops + i;
      functional_bootstrap(c_mid[idx], lut, c_in[idx], bk, 4);
      tlwe_keyswitch(c_out[idx], c_mid[idx], ksk);
    }
    pthread_barrier_wait(&bar);
  }
  return NULL;
}

int main(int argc, char **argv) {
  if (argc < 14) { fprintf(stderr, "usage: see header\n"); return 2; }
  n = atoi(argv[1]); N = atoi(argv[2]); k = atoi(argv[3]); l = atoi(argv[4]); Bg_bit = atoi(argv[5]);
  t = atoi(argv[6]); base_bit = atoi(argv[7]);
  const double lwe_sigma = atof(argv[8]), rlwe_sigma = atof(argv[9]);
  threads = atoi(argv[10]); ops = atoi(argv[11]); steps = atoi(argv[12]); warmup = atoi(argv[13]);

  init_fft(N);
  TLWE_Key key_tlwe = tlwe_new_binary_key(n, lwe_sigma);
  TLWE_Key key_tlwe_out = tlwe_new_binary_key(k * N, rlwe_sigma);
  TRLWE_Key key_trlwe = trlwe_new_binary_key(N, k, rlwe_sigma);
  trlwe_extract_tlwe_key(key_tlwe_out, key_trlwe);
  TRGSW_Key trgsw_key = trgsw_new_key(key_trlwe, l, Bg_bit);
  ksk = tlwe_new_KS_key(key_tlwe, key_tlwe_out, t, base_bit);
  bk = new_bootstrap_key(trgsw_key, key_tlwe, 1);

  /* LUT m -> (3m+1) mod 4 on torus_base 4 */
  Torus lut_vals[4];
  for (int m = 0; m < 4; m++) lut_vals[m] = int2torus((3 * m + 1) % 4, 3);
  lut = trlwe_alloc_new_sample(k, N);
  trlwe_torus_packing(lut, lut_vals, 4);

  const long total = (long)threads * ops;
  c_in = tlwe_alloc_sample_array(total, n);
  c_mid = tlwe_alloc_sample_array(total, k * N);
  c_out = tlwe_alloc_sample_array(total, n);
  uint64_t sm = 1;   /* splitmix64, seed 1: plaintext messages */
  int *msg = (int *)malloc(sizeof(int) * total);
  for (long i = 0; i < total; i++) {
    sm += 0x9E3779B97F4A7C15ull;
    uint64_t z = sm;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    z ^= z >> 31;
    msg[i] = (int)(z & 3);
    tlwe_sample(c_in[i], int2torus(msg[i], 3), key_tlwe);   /* encrypt on the main thread (RNG statics) */
  }

  pthread_barrier_init(&bar, NULL, threads + 1);
  pthread_t *th = (pthread_t *)malloc(sizeof(pthread_t) * threads);
  for (long i = 0; i < threads; i++) pthread_create(&th[i], NULL, worker, (void *)i);
  double *ms = (double *)malloc(sizeof(double) * (warmup + steps));
  for (int s = 0; s < warmup + steps; s++) {
    pthread_barrier_wait(&bar);
    const double t0 = now_ms();
    pthread_barrier_wait(&bar);
    ms[s] = now_ms() - t0;
  }
  for (long i = 0; i < threads; i++) pthread_join(th[i], NULL);

  long wrong = 0;
  for (long i = 0; i < total; i++) {
    const uint64_t dec = torus2int(tlwe_phase(c_out[i], key_tlwe), 3);
    if ((int)(dec & 7) != (3 * msg[i] + 1) % 4) wrong++;
  }
  double sum = 0;
  for (int s = warmup; s < warmup + steps; s++) sum += ms[s];
  printf("{\"threads\": %d, \"ops_per_step\": %ld, \"steps\": %d, \"warmup\": %d, \"ms_per_step\": %.3f, "
         "\"pbs_ks_per_s\": %.3f, \"wrong\": %ld}\n",
         threads, total, steps, warmup, sum / steps, total * steps / (sum * 1e-3), wrong);
  return wrong ? 1 : 0;
}
