#!/bin/bash
# the two ncu passes of B200_PROFILING.md on the final build (launch list of a short bench command, --set full of K2)
mkdir -p gpurun_out
CMD="python bench.py --steps 2 --warmup 3 --no-cpu-baseline"
$CMD > gpurun_out/plain.log 2>&1 || exit 1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv $CMD > gpurun_out/ncu_launches.log 2>&1; echo "ncu launches exit $?"
$CMD > gpurun_out/plain2.log 2>&1 && timeout 600 ncu --set full --clock-control none --import-source on -k regex:film_siren_tc_kernel -s 4 -c 1 -o gpurun_out/prof_tc -f $CMD > gpurun_out/ncu_full.log 2>&1; echo "ncu full exit $?"
timeout 300 python bench.py --steps 20 --warmup 5 > gpurun_out/bench.log 2>&1; echo "bench exit $?"; python tools/show_bench.py gpurun_out/bench.log | head -4
