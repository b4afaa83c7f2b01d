#!/bin/bash
# 1-GPU: whole GPU suite (incl. propagation), then the ncu launch list of the structured route at config 5
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q -s > gpurun_out/pytest_s.log 2>&1; echo "pytest exit $?"; grep -E "passed|failed|error" gpurun_out/pytest_s.log | tail -3
grep -E "propagation|FAILED|Error" gpurun_out/pytest_s.log | cut -c1-240 | head
JAICOV_SOLVER=structured python bench.py --config 5 --steps 1 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/s_plain.log 2>&1 && \
JAICOV_SOLVER=structured ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none \
    --launch-skip 777 -c 300 --csv --log-file gpurun_out/launches_c5_structured.csv \
    python bench.py --config 5 --steps 1 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/s_ncu.log 2>&1
echo "ncu exit $?"
