#!/bin/bash
# full GAN train step (U-Net + generator + discriminator): GPU parity tests, then the config-3 bench and the generator-only share
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_train_step.py tests/test_gpu_backward.py -q -m gpu -x -p no:cacheprovider > gpurun_out/pytest_trainstep.log 2>&1; echo "pytest exit $?" > gpurun_out/trainstep.log
tail -15 gpurun_out/pytest_trainstep.log >> gpurun_out/trainstep.log
timeout 900 python bench.py --workload train --steps 4 --warmup 3 > gpurun_out/bench_train_full.log 2>&1; echo "train full exit $?" >> gpurun_out/trainstep.log
timeout 300 python tools/bench_disc.py 2>&1 | grep -v Warning | head -2 >> gpurun_out/trainstep.log
cat gpurun_out/trainstep.log
python tools/show_bench.py gpurun_out/bench_train_full.log
