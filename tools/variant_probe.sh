#!/bin/bash
# A/B of the node / triangle fetch variants (Makefile `variant` target) on one B200: frame hashes must agree.
# usage: tools/variant_probe.sh "lib lib_n256 ..." ; writes gpurun_out/variants.txt
mkdir -p gpurun_out
OUT=gpurun_out/variants.txt
: > $OUT
for v in $1; do
  export RT_B200_LIBDIR=$PWD/raytrace_b200/$v
  echo "=== $v" | tee -a $OUT
  timeout 300 python tools/perf_probe.py c3 c2 2>&1 | cut -c1-420 | tee -a $OUT
  RT_PIPE_SHARE="4" RT_PIPE_M="3" timeout 200 python tools/pipe_probe.py c3 1 2>&1 | tail -1 | tee -a $OUT
  RT_PIPE_SHARE="1" RT_PIPE_M="8" timeout 200 python tools/pipe_probe.py c3 8 2>&1 | tail -1 | tee -a $OUT
done
unset RT_B200_LIBDIR
