#!/bin/bash
# residual SIREN variants: forward against the reference goldens (all precisions), backward against oracle autograd
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_backward.py -q -m gpu -s -p no:cacheprovider -k "Res or residual or generator_backward" > gpurun_out/pytest_dres.log 2>&1; echo "pytest exit $?"
grep -E "worst|passed|failed|Error|error|^\.?fwd_.*Res.*(fp32|bf16|fp16):" gpurun_out/pytest_dres.log | tail -30
