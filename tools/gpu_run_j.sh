#!/bin/bash
# 1-GPU: validate the register-level diagonal kernel and the DMMA skinny solves; launch list; bench 2/4/5
mkdir -p gpurun_out
timeout 900 python -m pytest tests -q -m gpu -s --maxfail=40 -p no:cacheprovider > gpurun_out/pytest_gpu.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_gpu.log; tail -3 gpurun_out/pytest_gpu.log
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 1 --master-addr 127.0.0.1 --master-port 29533 tests/multi_worker.py cfg2 > gpurun_out/world1_cfg2.log 2>&1
echo "world1 exit $?"; grep '^{' gpurun_out/world1_cfg2.log | tail -1 | cut -c1-500
for C in 2 4; do
  timeout 900 python bench.py --config $C --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_c$C.json 2> gpurun_out/bench_c$C.err
  echo "bench c$C exit $?"; grep -o '"ms_per_step[^}]*}' gpurun_out/bench_c$C.json | head -1
done
python bench.py --config 2 --steps 1 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/plain_c2.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 1200 --csv --log-file gpurun_out/launches_c2.csv \
    python bench.py --config 2 --steps 1 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/ncu_c2.log 2>&1
echo "ncu c2 exit $?"
timeout 900 python bench.py --steps 3 --warmup 3 > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err
echo "bench default exit $?"; cut -c1-2500 gpurun_out/bench_default.json
ls gpurun_out
