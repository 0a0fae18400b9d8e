#!/bin/bash
mkdir -p gpurun_out
CMD="python bench.py --steps 2 --warmup 3 --no-cpu-baseline"
$CMD > gpurun_out/plain3.log 2>&1 && \
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"raymarch|composite_kernel|channels_last|sample_pdf" -s 12 -c 6 -o gpurun_out/prof_small $CMD > gpurun_out/ncu_small.log 2>&1
echo "ncu exit $?"
