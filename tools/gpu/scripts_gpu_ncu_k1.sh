#!/bin/bash
mkdir -p gpurun_out
CMD="python bench.py --steps 2 --warmup 3 --no-cpu-baseline"
$CMD > gpurun_out/plain3.log 2>&1 && \
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"raymarch_gather" -s 4 -c 2 -o gpurun_out/prof_k1 $CMD > gpurun_out/ncu_k1.log 2>&1
echo "ncu exit $?"
