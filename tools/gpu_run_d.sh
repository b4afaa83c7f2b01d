#!/bin/bash
# 1-GPU run: tests (incl. distributed path with world = 1), benches, launch list for the latency-bound kernels
mkdir -p gpurun_out
timeout 900 python -m pytest tests -q -m gpu -s --maxfail=40 -p no:cacheprovider > gpurun_out/pytest_gpu.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_gpu.log; tail -4 gpurun_out/pytest_gpu.log
for sc in example cfg2 cfg4; do
  timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 1 --master-addr 127.0.0.1 --master-port 29533 tests/multi_worker.py $sc > gpurun_out/world1_$sc.log 2>&1
  echo "world1 $sc exit $?"; grep '^{' gpurun_out/world1_$sc.log | tail -1; tail -3 gpurun_out/world1_$sc.log | grep -v '^{' | cut -c1-300
done
timeout 900 python bench.py --config 4 --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/bench_c4.json 2> gpurun_out/bench_c4.err
echo "bench c4 exit $?"; cut -c1-1200 gpurun_out/bench_c4.json; tail -3 gpurun_out/bench_c4.err
timeout 600 python bench.py --config 2 --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_c2.json 2> gpurun_out/bench_c2.err
echo "bench c2 exit $?"; cut -c1-1200 gpurun_out/bench_c2.json
python bench.py --config 2 --steps 1 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/plain_c2.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 1600 --csv --log-file gpurun_out/launches_c2.csv \
    python bench.py --config 2 --steps 1 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/ncu_c2.log 2>&1
echo "ncu c2 exit $?"
ls gpurun_out
