#!/bin/bash
mkdir -p gpurun_out; : > gpurun_out/tc2.log
CNG_TC_CG=2 timeout 120 python tools/bench_mlp.py TALLSIREN_FG 20 >> gpurun_out/tc2.log 2>&1; echo "cg2 exit $?" >> gpurun_out/tc2.log
CNG_TC_CG=1 timeout 120 python tools/bench_mlp.py TALLSIREN_FG 20 >> gpurun_out/tc2.log 2>&1; echo "cg1 exit $?" >> gpurun_out/tc2.log
CNG_TC_CG=2 timeout 300 python -m pytest tests/test_gpu_parity.py -q -m gpu -p no:cacheprovider -x -k "film_siren or forward or psnr" >> gpurun_out/tc2.log 2>&1; echo "pytest cg2 exit $?" >> gpurun_out/tc2.log
tail -25 gpurun_out/tc2.log
