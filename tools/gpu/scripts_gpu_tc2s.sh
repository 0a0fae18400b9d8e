#!/bin/bash
# K2 kernel versions (CNG_TC_V): 1 slot-bound epilogue, 2 shared epilogue, 3 layer-pipelined -- bit-identity test, timings, timeline
mkdir -p gpurun_out; : > gpurun_out/tc2s.log
timeout 900 python -m pytest tests/test_gpu_parity.py -q -m gpu -x -p no:cacheprovider -k "film_siren" > gpurun_out/pytest_tc2s.log 2>&1; echo "pytest film_siren exit $?" >> gpurun_out/tc2s.log
tail -3 gpurun_out/pytest_tc2s.log >> gpurun_out/tc2s.log
for pass in 1 2; do for v in 1 2 3; do CNG_TC_V=$v timeout 300 python tools/bench_mlp.py TALLSIREN_FG 30 2>&1 | tail -1 >> gpurun_out/tc2s.log; done; done
for s in SHORTSIREN_FG DOUBLESIREN_FG SingleSIREN_dg; do for v in 1 2; do CNG_TC_V=$v timeout 300 python tools/bench_mlp.py $s 30 2>&1 | tail -1 >> gpurun_out/tc2s.log; done; done
for pl in 0 4; do CNG_TC_POLY=$pl CNG_TC_V=2 timeout 300 python tools/bench_mlp.py TALLSIREN_FG 30 2>&1 | tail -1 >> gpurun_out/tc2s.log; done
CNG_TC_V=2 timeout 300 python tools/trace_tc.py 1 > gpurun_out/trace_tc2s.log 2>&1
cat gpurun_out/tc2s.log; head -19 gpurun_out/trace_tc2s.log
