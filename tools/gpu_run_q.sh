#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q -s > gpurun_out/pytest_q.log 2>&1; echo "pytest exit $?"; grep -E "passed|failed|error" gpurun_out/pytest_q.log | tail -3
grep -E "\[structured\]|mixed|FAILED|Error" gpurun_out/pytest_q.log | cut -c1-240
