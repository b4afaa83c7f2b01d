#!/bin/bash
# 2-GPU run: multi-GPU parity tests and strong-scaling bench
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/gpus.txt
timeout 1200 python -m pytest tests/test_gpu_multi.py tests/test_gpu_parity.py::test_spd_solve_invert tests/test_gpu_parity.py::test_adjustment_config2 -q -m gpu -s -p no:cacheprovider > gpurun_out/pytest_multi.log 2>&1
echo "pytest multi exit $?" >> gpurun_out/pytest_multi.log; tail -25 gpurun_out/pytest_multi.log | cut -c1-600
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29541"
timeout 900 $TR bench.py --gpus 2 --config 4 --steps 3 --warmup 3 > gpurun_out/bench_c4_n2.json 2> gpurun_out/bench_c4_n2.err
echo "bench c4 n2 exit $?"; grep '^{' gpurun_out/bench_c4_n2.json | cut -c1-1500; tail -5 gpurun_out/bench_c4_n2.err | cut -c1-400
timeout 1500 $TR bench.py --gpus 2 --config 5 --steps 2 --warmup 3 > gpurun_out/bench_c5_n2.json 2> gpurun_out/bench_c5_n2.err
echo "bench c5 n2 exit $?"; grep '^{' gpurun_out/bench_c5_n2.json | cut -c1-1500; tail -5 gpurun_out/bench_c5_n2.err | cut -c1-400
timeout 600 python bench.py --config 4 --steps 3 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/bench_c4_n1.json 2> gpurun_out/bench_c4_n1.err
echo "bench c4 n1 exit $?"; cut -c1-900 gpurun_out/bench_c4_n1.json
ls gpurun_out
