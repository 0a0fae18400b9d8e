#!/bin/bash
# round 2, final build: ncu launch list of a short render bench and --set full of K2 (B200_PROFILING.md recipe)
mkdir -p gpurun_out
CMD="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-train --no-extras"
$CMD > gpurun_out/r2l_plain.log 2>&1 || exit 1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2l_launches.csv $CMD > gpurun_out/r2l_ncu_launches.log 2>&1; echo "ncu launches exit $?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:film_siren_tc_kernel -s 4 -c 1 -o gpurun_out/r2l_prof_tc -f $CMD > gpurun_out/r2l_ncu_tc.log 2>&1; echo "ncu K2 exit $?"
tail -1 gpurun_out/r2l_plain.log | cut -c1-200
