#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -q -m gpu -p no:cacheprovider -x -k "sample_pdf or resample or merge or composite or forward or backward" > gpurun_out/pytest_c5.log 2>&1; echo "pytest exit $?"; tail -3 gpurun_out/pytest_c5.log
timeout 900 python tools/bench_c5.py --rays-max 4 --json gpurun_out/c5.json > gpurun_out/c5.log 2>&1; cat gpurun_out/c5.log
