#!/bin/bash
# ping-pong K2 (CNG_TC_V=1) after the shift-row rework: parity, timing per variant, timeline, train-mode tests
mkdir -p gpurun_out; : > gpurun_out/tc1.log
timeout 900 python -m pytest tests/test_gpu_parity.py -q -m gpu -x -p no:cacheprovider -k "film_siren" > gpurun_out/pytest_tc1.log 2>&1; echo "pytest film_siren exit $?" >> gpurun_out/tc1.log
tail -3 gpurun_out/pytest_tc1.log >> gpurun_out/tc1.log
for s in TALLSIREN_FG SHORTSIREN_FG DOUBLESIREN_FG SingleSIREN_dg; do
  for v in 1 3; do CNG_TC_V=$v timeout 300 python tools/bench_mlp.py $s 30 2>&1 | tail -1 >> gpurun_out/tc1.log; done
done
for v in 1 3; do CNG_TC_V=$v timeout 300 python tools/bench_mlp.py TALLSIREN_FG 30 fp16 2>&1 | tail -1 >> gpurun_out/tc1.log; done
CNG_TC_V=1 timeout 300 python tools/trace_tc.py 1 > gpurun_out/trace_tc1.log 2>&1
timeout 900 python -m pytest tests/test_gpu_backward.py -q -m gpu -x -p no:cacheprovider > gpurun_out/pytest_bwd.log 2>&1; echo "pytest backward exit $?" >> gpurun_out/tc1.log
tail -3 gpurun_out/pytest_bwd.log >> gpurun_out/tc1.log
cat gpurun_out/tc1.log; head -19 gpurun_out/trace_tc1.log
