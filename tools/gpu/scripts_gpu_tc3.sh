#!/bin/bash
mkdir -p gpurun_out; : > gpurun_out/tc3.log
for p in 0 8 4 3 2; do CNG_TC_CG=1 CNG_TC_POLY=$p timeout 120 python tools/bench_mlp.py TALLSIREN_FG 20 >> gpurun_out/tc3.log 2>&1; done
CNG_TC_CG=2 timeout 120 python tools/bench_mlp.py TALLSIREN_FG 20 >> gpurun_out/tc3.log 2>&1; echo "cg2 exit $?" >> gpurun_out/tc3.log
CNG_TC_CG=2 timeout 300 python -m pytest tests/test_gpu_parity.py -q -m gpu -p no:cacheprovider -x -k "film_siren or forward or psnr" >> gpurun_out/tc3.log 2>&1; echo "pytest cg2 exit $?" >> gpurun_out/tc3.log
timeout 200 python tools/trace_tc.py 2 > gpurun_out/trace2.log 2>&1
cat gpurun_out/tc3.log | tail -12; sed -n 2,10p gpurun_out/trace2.log
