#!/bin/bash
mkdir -p gpurun_out; : > gpurun_out/mlp.log
for p in 0 8 4 3 2; do CNG_TC_POLY=$p timeout 300 python tools/bench_mlp.py TALLSIREN_FG 20 >> gpurun_out/mlp.log 2>&1; done
CNG_TC_POLY=4 timeout 300 python tools/bench_mlp.py SHORTSIREN_FG 20 >> gpurun_out/mlp.log 2>&1
CNG_TC_POLY=4 timeout 300 python tools/bench_mlp.py DOUBLESIREN_FG 20 >> gpurun_out/mlp.log 2>&1
cat gpurun_out/mlp.log
