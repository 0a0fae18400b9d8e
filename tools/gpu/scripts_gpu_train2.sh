#!/bin/bash
# train step: launch-overhead check at 4 images per GPU (profile), then the full step at N=2
mkdir -p gpurun_out
timeout 300 python tools/profile_gan_step.py 4 > gpurun_out/profile_gan_step_b4.log 2>&1; head -2 gpurun_out/profile_gan_step_b4.log; grep -E "Self C(PU|UDA) time total" gpurun_out/profile_gan_step_b4.log
timeout 600 python -m pytest tests/test_train_step.py -q -m gpu -p no:cacheprovider 2>&1 | tail -2
RUN="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2"
timeout 600 $RUN --workload train --steps 4 --warmup 3 > gpurun_out/n2_train.log 2>&1; echo "train N=2 exit $?"; python tools/show_bench.py gpurun_out/n2_train.log | head -2
timeout 600 $RUN --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/n2_render.log 2>&1; echo "render N=2 exit $?"; python tools/show_bench.py gpurun_out/n2_render.log | head -2
