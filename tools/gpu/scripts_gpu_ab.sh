#!/bin/bash
# same-box A/B of K2 build variants (CNG_LIB): two passes over all variants so box drift shows up
mkdir -p gpurun_out; : > gpurun_out/ab.log
for pass in 1 2; do
  for v in "" _grouped _hint1000 _hint100 _grouped_hint1000; do
    lib=conditioned_nerf_gan_b200/libcng_b200$v.so
    echo -n "pass $pass ${v:-default}: " >> gpurun_out/ab.log
    CNG_LIB=$PWD/$lib timeout 120 python tools/bench_mlp.py TALLSIREN_FG 30 2>&1 | tail -1 >> gpurun_out/ab.log
  done
done
cat gpurun_out/ab.log
