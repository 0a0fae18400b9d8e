#!/bin/bash
mkdir -p gpurun_out
CMD="python bench.py --steps 2 --warmup 3 --no-cpu-baseline"
$CMD > gpurun_out/plain2.log 2>&1 && \
timeout 900 ncu --set full --clock-control none --import-source on -k regex:film_siren_tc_kernel -s 4 -c 1 -o gpurun_out/prof_tc $CMD > gpurun_out/ncu_full.log 2>&1
echo "ncu full exit $?"
