#!/bin/bash
# 1-GPU: GPU suite after the structured-route changes, timings, then ncu --set full of the dominant GEMM launches
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q -s > gpurun_out/pytest_t.log 2>&1; echo "pytest exit $?"; grep -E "passed|failed|error" gpurun_out/pytest_t.log | tail -3
grep -E "FAILED|Error" gpurun_out/pytest_t.log | cut -c1-240 | head
for c in 2 4 5; do
  JAICOV_SOLVER=structured python bench.py --config $c --steps 3 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/t_c${c}_structured.log 2>&1
  echo "c$c structured: $(grep -o '"ms_per_step": [0-9.]*' gpurun_out/t_c${c}_structured.log) $(grep -o '"stage_ms": {[^}]*}' gpurun_out/t_c${c}_structured.log)"
done
# dense route, config 4: the single-launch LAUUM (k_gemm<1,1,65>) -- first such launch with >= 148 tiles is the big one
JAICOV_SOLVER=dense timeout 600 ncu --set full --clock-control none --import-source on --kernel-name-base mangled -k regex:k_gemmILi1ELi1ELi65E -c 1 \
    -o gpurun_out/ncu_full_lauum_c4_shape65 -f python bench.py --config 4 --steps 1 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/t_ncu1.log 2>&1
echo "ncu1 exit $?"
# structured route, config 5: the Y (Q'Y') product (second k_gemm<1,1,65> launch of a pass; the first is the small LAUUM)
JAICOV_SOLVER=structured timeout 900 ncu --set full --clock-control none --import-source on --kernel-name-base mangled -k regex:k_gemmILi1ELi1ELi65E --launch-skip 1 -c 1 \
    -o gpurun_out/ncu_full_qpp_c5_structured -f python bench.py --config 5 --steps 1 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/t_ncu2.log 2>&1
echo "ncu2 exit $?"
ls -la gpurun_out/*.ncu-rep
