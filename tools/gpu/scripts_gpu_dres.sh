#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -q -m gpu -x -s -p no:cacheprovider -k "residual or dRes or film_siren" > gpurun_out/pytest_dres.log 2>&1; echo "pytest exit $?"
grep -E "dRes|passed|failed|Error|error" gpurun_out/pytest_dres.log | tail -30
