#!/bin/bash
mkdir -p gpurun_out; : > gpurun_out/fp16.log
for s in TALLSIREN_FG SHORTSIREN_FG DOUBLESIREN_FG; do for p in bf16 fp16; do timeout 120 python tools/bench_mlp.py $s 20 $p >> gpurun_out/fp16.log 2>&1; done; done
timeout 600 python -m pytest tests/test_gpu_parity.py -q -m gpu -p no:cacheprovider -x -s -k "tensor_core or forward or psnr" 2>&1 | grep -E "passed|failed|fp16.*N=38405|fp16: coarse|Error" >> gpurun_out/fp16.log
cat gpurun_out/fp16.log
