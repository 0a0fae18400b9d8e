#!/bin/bash
mkdir -p gpurun_out
RUN="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2"
timeout 600 $RUN --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/n2_render.log 2>&1; echo "render exit $?"; grep '^{' gpurun_out/n2_render.log | cut -c1-400
timeout 600 $RUN --workload train --steps 3 --warmup 1 > gpurun_out/n2_train.log 2>&1; echo "train exit $?"; grep '^{' gpurun_out/n2_train.log | cut -c1-300
timeout 600 $RUN --workload video --steps 3 --warmup 1 > gpurun_out/n2_video.log 2>&1; echo "video exit $?"; grep '^{' gpurun_out/n2_video.log | cut -c1-300
timeout 300 $RUN --impl reference --steps 1 --warmup 0 > gpurun_out/n2_ref.log 2>&1; echo "ref exit $?"; grep '^{' gpurun_out/n2_ref.log | cut -c1-200
