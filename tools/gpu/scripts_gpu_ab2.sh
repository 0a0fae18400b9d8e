#!/bin/bash
mkdir -p gpurun_out; : > gpurun_out/ab2.log
for pass in 1 2 3; do for v in "" _grouped _hint2000; do
  echo -n "pass $pass lib${v:-_default}: " >> gpurun_out/ab2.log
  CNG_LIB=$PWD/conditioned_nerf_gan_b200/libcng_b200$v.so timeout 120 python tools/bench_mlp.py TALLSIREN_FG 30 2>&1 | tail -1 >> gpurun_out/ab2.log
done; done
cat gpurun_out/ab2.log
