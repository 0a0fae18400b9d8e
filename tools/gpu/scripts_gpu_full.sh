#!/bin/bash
# full refresh: GPU suite, smoke, bench (render/train/video), ncu launch list + full capture of K2, K2 timeline
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -q -m gpu -p no:cacheprovider > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?" > gpurun_out/summary.txt
timeout 300 python __graft_entry__.py --smoke > gpurun_out/smoke.log 2>&1; echo "smoke exit $?" >> gpurun_out/summary.txt
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/bench.log 2>&1; echo "bench exit $?" >> gpurun_out/summary.txt
timeout 600 python bench.py --steps 20 --warmup 5 --precision fp16 --no-cpu-baseline > gpurun_out/bench_fp16.log 2>&1; echo "bench fp16 exit $?" >> gpurun_out/summary.txt
timeout 600 python bench.py --workload train --steps 4 --warmup 3 > gpurun_out/bench_train.log 2>&1; echo "train exit $?" >> gpurun_out/summary.txt
timeout 600 python bench.py --workload train_generator --steps 4 --warmup 3 > gpurun_out/bench_train_generator.log 2>&1; echo "train_generator exit $?" >> gpurun_out/summary.txt
timeout 600 python bench.py --workload video --steps 4 --warmup 3 > gpurun_out/bench_video.log 2>&1; echo "video exit $?" >> gpurun_out/summary.txt
timeout 300 python bench.py --impl reference --steps 1 --warmup 0 > gpurun_out/bench_reference.log 2>&1; echo "reference arm exit $?" >> gpurun_out/summary.txt
timeout 300 python tools/profile_gan_step.py 8 > gpurun_out/profile_gan_step.log 2>&1
timeout 200 python tools/trace_tc.py 1 > gpurun_out/trace1.log 2>&1
: > gpurun_out/mlp_ab.log
for v in 1 2 3; do for pl in 0 8; do CNG_TC_V=$v CNG_TC_POLY=$pl timeout 200 python tools/bench_mlp.py TALLSIREN_FG 30 2>&1 | tail -1 >> gpurun_out/mlp_ab.log; done; done
for s in SHORTSIREN_FG DOUBLESIREN_FG SingleSIREN_dg; do timeout 200 python tools/bench_mlp.py $s 30 2>&1 | tail -1 >> gpurun_out/mlp_ab.log; done
timeout 200 python tools/bench_mlp.py SHORTSIREN_FG 30 fp16 2>&1 | tail -1 >> gpurun_out/mlp_ab.log
timeout 600 python tools/bench_c5.py --json gpurun_out/c5.json > gpurun_out/c5.log 2>&1; echo "c5 exit $?" >> gpurun_out/summary.txt
timeout 300 python tools/profile_train.py 4 > gpurun_out/profile_train.log 2>&1; echo "profile_train exit $?" >> gpurun_out/summary.txt
timeout 300 python tools/bench_graph.py > gpurun_out/graph.log 2>&1
CMD="python bench.py --steps 2 --warmup 3 --no-cpu-baseline"
$CMD > gpurun_out/plain.log 2>&1 && \
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv $CMD > gpurun_out/ncu_launches.log 2>&1
echo "ncu launches exit $?" >> gpurun_out/summary.txt
$CMD > gpurun_out/plain2.log 2>&1 && \
timeout 900 ncu --set full --clock-control none --import-source on -k regex:film_siren_tc_kernel -s 4 -c 1 -o gpurun_out/prof_tc $CMD > gpurun_out/ncu_full.log 2>&1
echo "ncu full exit $?" >> gpurun_out/summary.txt
cat gpurun_out/summary.txt; tail -3 gpurun_out/pytest_gpu.log; tail -2 gpurun_out/smoke.log
for f in bench bench_fp16 bench_train bench_train_generator bench_video; do python tools/show_bench.py gpurun_out/$f.log; done
cat gpurun_out/mlp_ab.log
