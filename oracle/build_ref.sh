#!/usr/bin/env bash
# TEST INFRASTRUCTURE ONLY -- builds the UNMODIFIED reference (antoniocgj/MOSFHET) CPU library
# from the sources where they lie under /root/reference into oracle/_ref/ (git-ignored; the
# built .so files travel to the GPU box with the gpurun snapshot).
#
# The reference's own Makefile is not run (it hard-codes -march=native, which would tie the
# binary to THIS container's CPU); this script restates its source lists and -D flags
# (Makefile.def:7-56) with an explicit ISA level per variant so the right one can be picked at
# run time from /proc/cpuinfo on the GPU box:
#
#   avx512   FFT_LIB=spqlios_avx512 A_PRNG=vaes   (Makefile.def defaults)       needs avx512f/dq/vl + vaes
#   fma      FFT_LIB=spqlios A_PRNG=none ENABLE_VAES=false                      needs avx2 + fma + aes + rdrand
#   portable PORTABLE_BUILD=1 A_PRNG=none ENABLE_VAES=false (scalar FFNT)       plain x86-64
#
# Nothing is copied out of /root/reference; only object code lands in oracle/_ref/.
set -euo pipefail
REF=${MOSFHET_REF:-/root/reference}
HERE="$(cd "$(dirname "${BASH_SOURCE[0]}")" && pwd)"
OUT="$HERE/_ref"
if [ ! -d "$REF/src" ]; then
  echo "build_ref: $REF not present (GPU box?) -- keeping prebuilt files in $OUT" >&2
  exit 0
fi
mkdir -p "$OUT"
S="$REF/src"
COMMON_SRC="$S/keyswitch.c $S/bootstrap.c $S/bootstrap_ga.c $S/tlwe.c $S/trlwe.c $S/trgsw.c $S/misc.c $S/polynomial.c $S/register.c $S/sha3/fips202.c $S/fft/karatsuba.c"
COMMON_FLAGS="-O3 -g0 -fPIC -shared -w -funroll-all-loops -I$REF/include"
SPQ="$S/fft/spqlios"

build() { # name, march flags, -D flags, extra sources
  local name=$1 march=$2 defs=$3 extra=$4
  local so="$OUT/libmosfhet_${name}.so"
  if [ -f "$so" ] && [ "$so" -nt "$0" ]; then return; fi
  echo "build_ref: $so"
  gcc $COMMON_FLAGS $march $defs $COMMON_SRC $extra -lm -o "$so"
}

build avx512 "-march=x86-64-v4 -mvaes -maes -mrdrnd -mpclmul" \
  "-DUSE_SPQLIOS -DAVX512_OPT -DUSE_COMPRESSED_TRLWE -DVAES_OPT" \
  "$S/trlwe_compressed_vaes.c $S/rnd/aes_rng.c $SPQ/spqlios-fft-avx512.s $SPQ/spqlios-ifft-avx512.s $SPQ/spqlios-fft-impl-avx512.c $SPQ/fft_processor_spqlios.c"

build fma "-march=x86-64-v3 -maes -mrdrnd -mpclmul" \
  "-DUSE_SPQLIOS" \
  "$S/rnd/aes_rng.c $SPQ/spqlios-fft-fma.s $SPQ/spqlios-ifft-fma.s $SPQ/spqlios-fft-impl.c $SPQ/fft_processor_spqlios.c"

build portable "-march=x86-64" \
  "-DPORTABLE_BUILD -DUSE_SHAKE" \
  "$S/fft/ffnt/ffnt.c"

# Multi-threaded timing driver for the CPU baseline (bench.py cpu_baseline / --impl reference).
for v in avx512 fma portable; do
  exe="$OUT/ref_bench_$v"
  if [ ! -f "$exe" ] || [ "$HERE/ref_bench.c" -nt "$exe" ] || [ "$OUT/libmosfhet_$v.so" -nt "$exe" ]; then
    if [ -f "$HERE/ref_bench.c" ]; then
      echo "build_ref: $exe"
      pd=""; [ "$v" = portable ] && pd="-DPORTABLE_BUILD"
      gcc -O2 -w $pd -I"$REF/include" "$HERE/ref_bench.c" -o "$exe" \
        -L"$OUT" -l:libmosfhet_$v.so -Wl,-rpath,'$ORIGIN' -lpthread -lm
    fi
  fi
done
# The reference's own benchmark build (Makefile.def:2-6: -O3 -fwhole-program -flto, all sources in one program) as a second
# CPU variant of the timing driver, with the same explicit ISA level instead of -march=native.
exe="$OUT/ref_bench_avx512_lto"
if [ -f "$HERE/ref_bench.c" ] && { [ ! -f "$exe" ] || [ "$HERE/ref_bench.c" -nt "$exe" ] || [ "$0" -nt "$exe" ]; }; then
  echo "build_ref: $exe"
  gcc -O3 -fwhole-program -flto -g0 -w -funroll-all-loops -I"$REF/include" \
    -march=x86-64-v4 -mvaes -maes -mrdrnd -mpclmul -DUSE_SPQLIOS -DAVX512_OPT -DUSE_COMPRESSED_TRLWE -DVAES_OPT \
    $COMMON_SRC $S/trlwe_compressed_vaes.c $S/rnd/aes_rng.c $SPQ/spqlios-fft-avx512.s $SPQ/spqlios-ifft-avx512.s \
    $SPQ/spqlios-fft-impl-avx512.c $SPQ/fft_processor_spqlios.c "$HERE/ref_bench.c" -lpthread -lm -o "$exe"
fi
echo "build_ref: done"
