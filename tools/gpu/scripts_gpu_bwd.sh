#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_backward.py -q -m gpu -s -p no:cacheprovider -x > gpurun_out/pytest_bwd.log 2>&1; echo "pytest exit $?"
grep -E "passed|failed|rel-L2|worst|Error|assert" gpurun_out/pytest_bwd.log | tail -70
