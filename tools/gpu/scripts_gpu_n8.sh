#!/bin/bash
mkdir -p gpurun_out
N=${1:-8}
RUN="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus $N"
timeout 300 $RUN --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/n${N}_render.log 2>&1; echo "render exit $?"; python tools/show_bench.py gpurun_out/n${N}_render.log | head -3
timeout 300 $RUN --workload train --steps 4 --warmup 3 > gpurun_out/n${N}_train.log 2>&1; echo "train exit $?"; python tools/show_bench.py gpurun_out/n${N}_train.log | head -2
timeout 300 $RUN --workload train_generator --steps 3 --warmup 2 > gpurun_out/n${N}_train_generator.log 2>&1; echo "train_generator exit $?"; python tools/show_bench.py gpurun_out/n${N}_train_generator.log | head -2
timeout 300 $RUN --workload video --steps 3 --warmup 2 > gpurun_out/n${N}_video.log 2>&1; echo "video exit $?"; python tools/show_bench.py gpurun_out/n${N}_video.log | head -2
