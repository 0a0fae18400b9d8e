#!/bin/bash
mkdir -p gpurun_out
N=4
RUN="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29519 bench.py --gpus $N"
timeout 300 $RUN --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/n4_render.log 2>&1; echo "render exit $?"; python tools/show_bench.py gpurun_out/n4_render.log 2>/dev/null | head -2
timeout 400 $RUN --workload train --steps 4 --warmup 3 > gpurun_out/n4_train.log 2>&1; echo "train exit $?"; python tools/show_bench.py gpurun_out/n4_train.log 2>/dev/null | head -1
