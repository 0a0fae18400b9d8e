#!/bin/bash
# layer-pipelined K2 (film_siren_tc3.cu): parity, A/B against the ping-pong kernel (CNG_TC_V=1), timeline, merge fast path, bench
mkdir -p gpurun_out; : > gpurun_out/tc3.log
timeout 900 python -m pytest tests/test_gpu_parity.py -q -m gpu -x -p no:cacheprovider -k "film_siren" > gpurun_out/pytest_tc3.log 2>&1; echo "pytest film_siren exit $?" >> gpurun_out/tc3.log
tail -3 gpurun_out/pytest_tc3.log >> gpurun_out/tc3.log
for s in TALLSIREN_FG SHORTSIREN_FG DOUBLESIREN_FG SingleSIREN_dg; do
  for v in 1 3; do CNG_TC_V=$v timeout 300 python tools/bench_mlp.py $s 20 2>&1 | tail -1 >> gpurun_out/tc3.log; done
done
for v in 1 3; do CNG_TC_V=$v timeout 300 python tools/bench_mlp.py TALLSIREN_FG 20 fp16 2>&1 | tail -1 >> gpurun_out/tc3.log; done
for pl in 8 4; do CNG_TC_POLY=$pl CNG_TC_V=3 timeout 300 python tools/bench_mlp.py TALLSIREN_FG 20 2>&1 | tail -1 >> gpurun_out/tc3.log; done
CNG_TC_V=3 timeout 300 python tools/trace_tc.py 1 > gpurun_out/trace_tc3.log 2>&1
timeout 900 python -m pytest tests/ -q -m gpu -x -p no:cacheprovider -k "merge or composite or backward or forward" > gpurun_out/pytest_merge.log 2>&1; echo "pytest merge/composite/backward/forward exit $?" >> gpurun_out/tc3.log
tail -3 gpurun_out/pytest_merge.log >> gpurun_out/tc3.log
timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/bench_tc3.log 2>&1; echo "bench exit $?" >> gpurun_out/tc3.log
grep '^{' gpurun_out/bench_tc3.log | python tools/show_bench.py >> gpurun_out/tc3.log 2>&1
cat gpurun_out/tc3.log; head -25 gpurun_out/trace_tc3.log
