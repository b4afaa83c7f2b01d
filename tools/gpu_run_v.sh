#!/bin/bash
# 1-GPU: whole GPU suite after moving the index construction to the device, then the default bench (e2e phases)
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q > gpurun_out/pytest_v.log 2>&1; echo "pytest exit $?"; grep -E "passed|failed|error" gpurun_out/pytest_v.log | tail -3
grep -E "FAILED|Error" gpurun_out/pytest_v.log | cut -c1-240 | head
bash tools/gpu_run_scale.sh 1
