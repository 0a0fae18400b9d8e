#!/bin/bash
# first GPU pass: per-group pytest (separate processes so a sticky CUDA error does not poison the rest), smoke, bench
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,driver_version,clocks.max.sm --format=csv > gpurun_out/gpu.txt 2>&1
for grp in channels_last gather raymarch "film_siren_fp32" "film_siren_bf16" composite "sample_pdf or resample" merge "forward or siren_secondary or staged or errors"; do
  name=$(echo "$grp" | tr ' ' '_')
  timeout 600 python -m pytest tests/test_gpu_parity.py -q -m gpu -k "$grp" -s -p no:cacheprovider > gpurun_out/pytest_$name.log 2>&1
  echo "group [$grp] exit $?" >> gpurun_out/summary.txt
done
timeout 300 python __graft_entry__.py --smoke > gpurun_out/smoke.log 2>&1; echo "smoke exit $?" >> gpurun_out/summary.txt
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/bench.log 2>&1; echo "bench exit $?" >> gpurun_out/summary.txt
cat gpurun_out/summary.txt
tail -3 gpurun_out/bench.log
