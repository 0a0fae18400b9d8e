#!/bin/bash
# parity suite + smoke + bench, then the two ncu passes of B200_PROFILING.md on a short bench command
mkdir -p gpurun_out
timeout 900 python -m pytest tests -q -m gpu -s -p no:cacheprovider > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?" > gpurun_out/summary.txt
timeout 300 python __graft_entry__.py --smoke > gpurun_out/smoke.log 2>&1; echo "smoke exit $?" >> gpurun_out/summary.txt
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/bench.log 2>&1; echo "bench exit $?" >> gpurun_out/summary.txt
CMD="python bench.py --steps 2 --warmup 3 --no-cpu-baseline"
$CMD > gpurun_out/plain.log 2>&1 && \
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv $CMD > gpurun_out/ncu_launches.log 2>&1
echo "ncu launches exit $?" >> gpurun_out/summary.txt
$CMD > gpurun_out/plain2.log 2>&1 && \
timeout 900 ncu --set full --clock-control none --import-source on -k regex:film_siren_tc_kernel -s 4 -c 2 -o gpurun_out/prof_tc $CMD > gpurun_out/ncu_full.log 2>&1
echo "ncu full exit $?" >> gpurun_out/summary.txt
cat gpurun_out/summary.txt; tail -5 gpurun_out/pytest_gpu.log; tail -2 gpurun_out/smoke.log
