#!/bin/bash
# tc3 tuning: sub-block width variants (CNG_LIB) x sine split, timeline of the default build
mkdir -p gpurun_out; : > gpurun_out/tc3b.log
for pass in 1 2; do
  for v in "" _cw4 _cw16; do
    lib=$PWD/conditioned_nerf_gan_b200/libcng_b200$v.so
    [ -f $lib ] || continue
    echo -n "pass $pass lib${v:-_cw8}: " >> gpurun_out/tc3b.log
    CNG_LIB=$lib CNG_TC_V=3 timeout 120 python tools/bench_mlp.py TALLSIREN_FG 30 2>&1 | tail -1 >> gpurun_out/tc3b.log
  done
done
for v in "" _cw4 _cw16; do
  lib=$PWD/conditioned_nerf_gan_b200/libcng_b200$v.so
  [ -f $lib ] || continue
  for pl in 8 4; do echo -n "lib${v:-_cw8}: " >> gpurun_out/tc3b.log; CNG_LIB=$lib CNG_TC_POLY=$pl CNG_TC_V=3 timeout 120 python tools/bench_mlp.py TALLSIREN_FG 30 2>&1 | tail -1 >> gpurun_out/tc3b.log; done
  CNG_LIB=$lib CNG_TC_V=3 timeout 120 python tools/trace_tc.py 1 > gpurun_out/trace_tc3${v}.log 2>&1
done
CNG_TC_V=1 timeout 120 python tools/bench_mlp.py TALLSIREN_FG 30 2>&1 | tail -1 >> gpurun_out/tc3b.log
cat gpurun_out/tc3b.log; for v in "" _cw4 _cw16; do head -19 gpurun_out/trace_tc3${v}.log | awk 'NR==1 || NR%2==0'; done
