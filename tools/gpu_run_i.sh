#!/bin/bash
# 1-GPU: tests, 8 vs 16 warp A/B at config 4 and 5, ncu source-level captures of the latency-bound kernels
mkdir -p gpurun_out
timeout 900 python -m pytest tests -q -m gpu -s --maxfail=40 -p no:cacheprovider > gpurun_out/pytest_gpu.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_gpu.log; tail -3 gpurun_out/pytest_gpu.log
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 1 --master-addr 127.0.0.1 --master-port 29533 tests/multi_worker.py cfg2 > gpurun_out/world1_cfg2.log 2>&1
echo "world1 exit $?"; grep '^{' gpurun_out/world1_cfg2.log | tail -1 | cut -c1-500
for W in 16 8; do
  JAICOV_GEMM_WARPS=$W timeout 900 python bench.py --config 4 --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/bench_c4_w$W.json 2> gpurun_out/bench_c4_w$W.err
  echo "bench c4 warps $W exit $?"; cut -c1-1000 gpurun_out/bench_c4_w$W.json | grep -o '"ms_per_step[^}]*}' | head -2
done
for W in 16 8; do
  JAICOV_GEMM_WARPS=$W timeout 900 python bench.py --config 5 --steps 2 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/bench_c5_w$W.json 2> gpurun_out/bench_c5_w$W.err
  echo "bench c5 warps $W exit $?"; cut -c1-1000 gpurun_out/bench_c5_w$W.json | grep -o '"ms_per_step[^}]*}' | head -2
done
python bench.py --config 2 --steps 1 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/plain_c2.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:'k_potrf_diag|k_solve_bwd_step|k_solve_fwd_step' --launch-skip 24 -c 3 -o gpurun_out/prof_small \
    python bench.py --config 2 --steps 1 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/ncu_small.log 2>&1
echo "ncu small exit $?"; tail -2 gpurun_out/ncu_small.log
ls gpurun_out
