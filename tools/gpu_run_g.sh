#!/bin/bash
# 2-GPU strong-scaling bench
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29541"
timeout 900 $TR bench.py --gpus 2 --config 4 --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/bench_c4_n2.json 2> gpurun_out/bench_c4_n2.err
echo "bench c4 n2 exit $?"; grep '^{' gpurun_out/bench_c4_n2.json | cut -c1-1700; grep -E "Error|error" gpurun_out/bench_c4_n2.err | tail -3 | cut -c1-400
timeout 1500 $TR bench.py --gpus 2 --config 5 --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/bench_c5_n2.json 2> gpurun_out/bench_c5_n2.err
echo "bench c5 n2 exit $?"; grep '^{' gpurun_out/bench_c5_n2.json | cut -c1-1700; grep -E "Error|error" gpurun_out/bench_c5_n2.err | tail -3 | cut -c1-400
ls gpurun_out
