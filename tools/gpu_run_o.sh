#!/bin/bash
# 1-GPU: A/B of the 64x128 / 4-warp / 2-CTA-per-SM GEMM shape against the 8-warp default, then memcheck of smoke()
mkdir -p gpurun_out
JAICOV_GEMM_WARPS=4 timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_w4.log 2>&1; echo "pytest(w4) exit $?"; tail -3 gpurun_out/pytest_w4.log
for w in 8 4 8 4; do
  JAICOV_GEMM_WARPS=$w python bench.py --config 5 --steps 2 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/ab_c5_w$w.log 2>&1
  echo "c5 w=$w: $(grep -o '"ms_per_step": [0-9.]*' gpurun_out/ab_c5_w$w.log) $(grep -o '"frac": [0-9.]*' gpurun_out/ab_c5_w$w.log)"
done
for w in 8 4; do
  JAICOV_GEMM_WARPS=$w python bench.py --config 4 --steps 10 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/ab_c4_w$w.log 2>&1
  echo "c4 w=$w: $(grep -o '"ms_per_step": [0-9.]*' gpurun_out/ab_c4_w$w.log) $(grep -o '"frac": [0-9.]*' gpurun_out/ab_c4_w$w.log)"
done
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 9 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/memcheck_smoke.log 2>&1
echo "memcheck exit $?"; tail -5 gpurun_out/memcheck_smoke.log
