#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q -s > gpurun_out/pytest_w.log 2>&1; echo "pytest exit $?"; grep -E "passed|failed|error" gpurun_out/pytest_w.log | tail -3
grep -E "ragged|FAILED|Error" gpurun_out/pytest_w.log | cut -c1-260 | head -20
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
