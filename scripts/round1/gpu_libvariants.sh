#!/usr/bin/env bash
# times the default k1 variant of each alternative build of the library (MB200_LIB_NAME)
mkdir -p gpurun_out
for lib in libmosfhet_b200.so $(cd mosfhet_b200 && ls libmb_*.so 2>/dev/null); do
  echo "== $lib"
  MB200_LIB_NAME=$lib MB200_ONLY_DEFAULT=1 timeout 600 python scripts/k1_variants.py 2>&1 | tail -3
done | tee gpurun_out/k1_libvariants.log
