#!/bin/bash
# 8-GPU run: parity at N=8 (small scene), strong-scaling bench at N=8 and N=4 (config 5)
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/gpus8.txt
TR8="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29551"
TR4="python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29552"
timeout 300 $TR8 tests/multi_worker.py cfg2 > gpurun_out/world8_cfg2.log 2>&1
echo "world8 cfg2 exit $?"; grep '^{' gpurun_out/world8_cfg2.log | tail -1 | cut -c1-600
timeout 300 $TR8 tests/multi_worker.py example > gpurun_out/world8_example.log 2>&1
echo "world8 example exit $?"; grep '^{' gpurun_out/world8_example.log | tail -1 | cut -c1-600
timeout 900 $TR8 bench.py --gpus 8 --config 5 --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/bench_c5_n8.json 2> gpurun_out/bench_c5_n8.err
echo "bench c5 n8 exit $?"; grep '^{' gpurun_out/bench_c5_n8.json | cut -c1-1500; grep -E "Error|error" gpurun_out/bench_c5_n8.err | tail -3 | cut -c1-400
timeout 900 $TR4 bench.py --gpus 4 --config 5 --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/bench_c5_n4.json 2> gpurun_out/bench_c5_n4.err
echo "bench c5 n4 exit $?"; grep '^{' gpurun_out/bench_c5_n4.json | cut -c1-1500; grep -E "Error|error" gpurun_out/bench_c5_n4.err | tail -3 | cut -c1-400
ls gpurun_out
