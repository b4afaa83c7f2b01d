#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -q -m gpu -s --maxfail=40 -p no:cacheprovider > gpurun_out/pytest_gpu.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_gpu.log; tail -3 gpurun_out/pytest_gpu.log | cut -c1-300
for C in 2 4; do
  timeout 900 python bench.py --config $C --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_c$C.json 2> gpurun_out/bench_c$C.err
  echo "bench c$C exit $?"; grep -o '"ms_per_step[^}]*}' gpurun_out/bench_c$C.json | head -1
done
python bench.py --config 2 --steps 1 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/plain_c2.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 1200 --csv --log-file gpurun_out/launches_c2.csv \
    python bench.py --config 2 --steps 1 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/ncu_c2.log 2>&1
echo "ncu c2 exit $?"
