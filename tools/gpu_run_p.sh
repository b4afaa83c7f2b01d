#!/bin/bash
# 1-GPU: parity of the structured route, then config 2/4/5 timings of both routes
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q -s > gpurun_out/pytest_p.log 2>&1; echo "pytest exit $?"; grep -E "passed|failed|error" gpurun_out/pytest_p.log | tail -3
grep -E "\[structured\]|\[dense\]|mixed|fixed datum" gpurun_out/pytest_p.log | cut -c1-220
for c in 2 4 5; do
 for sv in dense structured; do
  JAICOV_SOLVER=$sv python bench.py --config $c --steps 3 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/p_c${c}_$sv.log 2>&1
  echo "c$c $sv: $(grep -o '"ms_per_step": [0-9.]*' gpurun_out/p_c${c}_$sv.log) $(grep -o '"phase_ms": {[^}]*}' gpurun_out/p_c${c}_$sv.log)"
  tail -2 gpurun_out/p_c${c}_$sv.log | grep -v '^{' | cut -c1-300
 done
done
