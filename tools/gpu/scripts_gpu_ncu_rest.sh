#!/bin/bash
# ncu --set full of the non-MLP kernels of the render step (K1 gather, K3 composite, K3' merge+composite, K4 resample, layout)
mkdir -p gpurun_out
CMD="python bench.py --steps 2 --warmup 3 --no-cpu-baseline"
$CMD > gpurun_out/plain3.log 2>&1 || exit 1
for k in raymarch_gather_kernel composite_kernel sample_pdf_kernel channels_last_kernel; do
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:$k -s 6 -c 2 -o gpurun_out/prof_$k -f $CMD > gpurun_out/ncu_$k.log 2>&1
  echo "ncu $k exit $?"
done
ls -la gpurun_out/prof_*.ncu-rep
