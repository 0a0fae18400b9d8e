#!/bin/bash
mkdir -p gpurun_out; : > gpurun_out/tc1b.log
for pass in 1 2; do
for v in "" _pipe; do
  lib=$PWD/conditioned_nerf_gan_b200/libcng_b200$v.so
  echo -n "lib${v:-_default}: " >> gpurun_out/tc1b.log
  CNG_LIB=$lib CNG_TC_V=1 timeout 120 python tools/bench_mlp.py TALLSIREN_FG 30 2>&1 | tail -1 >> gpurun_out/tc1b.log
done; done
CNG_LIB=$PWD/conditioned_nerf_gan_b200/libcng_b200_pipe.so CNG_TC_V=1 timeout 120 python tools/trace_tc.py 1 > gpurun_out/trace_tc1_pipe.log 2>&1
cat gpurun_out/tc1b.log; head -13 gpurun_out/trace_tc1_pipe.log
