#!/bin/bash
# round 2: the two ncu passes of B200_PROFILING.md -- launch list of a short render bench and of one MLP-backward chunk,
# --set full of the dominant kernels (K2 inference, training-mode K2, dgrad chain, split-K weight gradient)
mkdir -p gpurun_out
CMD="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-train --no-extras"
BWD="python tools/bench_bwd.py --points 524288 --reps 1"
$CMD > gpurun_out/r2_plain.log 2>&1 || exit 1
$BWD > gpurun_out/r2_plain_bwd.log 2>&1 || exit 1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2_launches.csv $CMD > gpurun_out/r2_ncu_launches.log 2>&1; echo "ncu launches exit $?"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2_launches_bwd.csv $BWD > gpurun_out/r2_ncu_launches_bwd.log 2>&1; echo "ncu launches (bwd) exit $?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:film_siren_tc_kernel -s 4 -c 1 -o gpurun_out/r2_prof_tc -f $CMD > gpurun_out/r2_ncu_tc.log 2>&1; echo "ncu K2 exit $?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:film_siren_dgrad -s 2 -c 1 -o gpurun_out/r2_prof_dgrad -f $BWD > gpurun_out/r2_ncu_dgrad.log 2>&1; echo "ncu dgrad exit $?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:film_siren_wgrad -s 2 -c 1 -o gpurun_out/r2_prof_wgrad -f $BWD > gpurun_out/r2_ncu_wgrad.log 2>&1; echo "ncu wgrad exit $?"
# the training-mode forward is the film_siren_tc_kernel<4, true, true, ...> instantiation: the 6th and later launches of the tool
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"film_siren_tc_kernel<4" -s 2 -c 1 -o gpurun_out/r2_prof_tctrain -f $BWD > gpurun_out/r2_ncu_tctrain.log 2>&1; echo "ncu train fwd exit $?"
ls -la gpurun_out/r2_prof_*.ncu-rep
