#!/bin/bash
# last check of the round: whole GPU suite, smoke, default bench line, c5 sweep after the merge-path change
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -q -m gpu -p no:cacheprovider > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?" > gpurun_out/final.log
tail -3 gpurun_out/pytest_gpu.log >> gpurun_out/final.log
timeout 300 python __graft_entry__.py --smoke > gpurun_out/smoke.log 2>&1; echo "smoke exit $?" >> gpurun_out/final.log
timeout 600 python bench.py > gpurun_out/bench_default.log 2>&1; echo "bench (no flags) exit $?" >> gpurun_out/final.log
timeout 600 python tools/bench_c5.py --rays-max 4 --json gpurun_out/c5.json > gpurun_out/c5.log 2>&1; echo "c5 exit $?" >> gpurun_out/final.log
cat gpurun_out/final.log; python tools/show_bench.py gpurun_out/bench_default.log; grep merge_composite gpurun_out/c5.log
