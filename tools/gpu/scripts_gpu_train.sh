#!/bin/bash
mkdir -p gpurun_out
timeout 900 python bench.py --workload train --steps 3 --warmup 1 > gpurun_out/bench_train.log 2>&1; echo "train exit $?"; tail -3 gpurun_out/bench_train.log | cut -c1-1800
