#!/bin/bash
mkdir -p gpurun_out; : > gpurun_out/tc1.log
CNG_TC_CG=1 timeout 120 python tools/bench_mlp.py TALLSIREN_FG 20 >> gpurun_out/tc1.log 2>&1; echo "cg1 exit $?" >> gpurun_out/tc1.log
CNG_TC_CG=1 timeout 120 python tools/bench_mlp.py SHORTSIREN_FG 20 >> gpurun_out/tc1.log 2>&1
CNG_TC_CG=1 timeout 120 python tools/bench_mlp.py DOUBLESIREN_FG 20 >> gpurun_out/tc1.log 2>&1
CNG_TC_CG=1 timeout 300 python -m pytest tests/test_gpu_parity.py -q -m gpu -p no:cacheprovider -x -k "film_siren or forward or psnr" >> gpurun_out/tc1.log 2>&1; echo "pytest cg1 exit $?" >> gpurun_out/tc1.log
timeout 200 python tools/trace_tc.py 1 > gpurun_out/trace1.log 2>&1
tail -12 gpurun_out/tc1.log; sed -n 2,12p gpurun_out/trace1.log
