#!/bin/bash
# quick iteration: MLP + forward parity, then bench
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -q -m gpu -s -p no:cacheprovider -k "film_siren or forward or psnr or staged" > gpurun_out/pytest_quick.log 2>&1; echo "pytest exit $?" > gpurun_out/summary.txt
timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/bench.log 2>&1; echo "bench exit $?" >> gpurun_out/summary.txt
cat gpurun_out/summary.txt; tail -4 gpurun_out/pytest_quick.log; python - <<'PY'
import json
try:
    d=json.loads([x for x in open("gpurun_out/bench.log") if x.startswith("{")][-1])
    r=d["roofline"]
    print("rays/s %.3e  ms/step %.3f  e2e %.3e  mlp ms %.3f  TFLOPs %.1f frac %.3f  clocks %s" % (d["value"], d["ms_per_step"], d["e2e"]["value"], r["ms_per_launch"], r["achieved"], r["frac"], d["clocks"]))
    print(r["step_ms_by_entry_point"])
except Exception as e:
    print("bench parse failed", e); print(open("gpurun_out/bench.log").read()[-2000:])
PY
