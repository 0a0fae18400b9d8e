#!/bin/bash
# last check of the round: whole GPU suite, smoke, default bench line, c5 sweep after the merge-path change
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -q -m gpu -p no:cacheprovider > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?" > gpurun_out/final.log
tail -3 gpurun_out/pytest_gpu.log >> gpurun_out/final.log
timeout 300 python __graft_entry__.py --smoke > gpurun_out/smoke.log 2>&1; echo "smoke exit $?" >> gpurun_out/final.log
timeout 600 python bench.py > gpurun_out/bench_default.log 2>&1; echo "bench (no flags) exit $?" >> gpurun_out/final.log
timeout 600 python bench.py --workload train --steps 4 --warmup 3 > gpurun_out/bench_train.log 2>&1; echo "train exit $?" >> gpurun_out/final.log
timeout 300 python tools/profile_gan_step.py 4 > gpurun_out/profile_gan_step_b4.log 2>&1
cat gpurun_out/final.log; python tools/show_bench.py gpurun_out/bench_default.log; python tools/show_bench.py gpurun_out/bench_train.log | head -2; head -1 gpurun_out/profile_gan_step_b4.log; grep -E "Self C(PU|UDA) time total" gpurun_out/profile_gan_step_b4.log
