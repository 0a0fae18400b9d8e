#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_backward.py -q -m gpu -x -s -p no:cacheprovider -k "generator_backward" > gpurun_out/pytest_dres.log 2>&1; echo "pytest exit $?"
grep -E "worst|passed|failed|Error|error|dRes siren" gpurun_out/pytest_dres.log | tail -30
