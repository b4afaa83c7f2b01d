#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_dlt.py tests/test_propagation.py -m gpu -q -s > gpurun_out/pytest_u.log 2>&1; echo "pytest exit $?"; grep -E "passed|failed|error" gpurun_out/pytest_u.log | tail -3
grep -E "^DLT|\.DLT|FAILED|Error|assert" gpurun_out/pytest_u.log | cut -c1-260 | head -30
