#!/bin/bash
# tc3 bottleneck experiments (timing only; exp builds give wrong results by design)
mkdir -p gpurun_out; : > gpurun_out/tc3c.log
for v in "" _s16 _s100 _s101 _s102; do
  lib=$PWD/conditioned_nerf_gan_b200/libcng_b200$v.so
  [ -f $lib ] || continue
  echo -n "lib${v:-_default}: " >> gpurun_out/tc3c.log
  CNG_LIB=$lib CNG_TC_V=3 timeout 120 python tools/bench_mlp.py TALLSIREN_FG 30 2>&1 | tail -1 >> gpurun_out/tc3c.log
  CNG_LIB=$lib CNG_TC_V=3 timeout 120 python tools/trace_tc.py 1 > gpurun_out/trace_tc3${v}.log 2>&1
done
cat gpurun_out/tc3c.log; for v in "" _s16 _s100 _s101 _s102; do echo "== $v"; head -19 gpurun_out/trace_tc3${v}.log | awk 'NR==1 || (NR>=6 && NR%2==0 && NR<=14)'; done
