#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -q -m gpu -s --maxfail=40 -p no:cacheprovider > gpurun_out/pytest_gpu.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_gpu.log; tail -4 gpurun_out/pytest_gpu.log
timeout 900 python bench.py --config 4 --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/bench_c4.json 2> gpurun_out/bench_c4.err
echo "bench c4 exit $?"; cut -c1-1000 gpurun_out/bench_c4.json; tail -3 gpurun_out/bench_c4.err
timeout 600 python bench.py --config 2 --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_c2.json 2> gpurun_out/bench_c2.err
echo "bench c2 exit $?"; cut -c1-1000 gpurun_out/bench_c2.json
python bench.py --config 4 --steps 1 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/plain_c4.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -s 4700 -c 1700 --csv --log-file gpurun_out/launches_c4.csv \
    python bench.py --config 4 --steps 1 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/ncu_c4.log 2>&1
echo "ncu c4 list exit $?"
ls gpurun_out
