#!/bin/bash
# 2-GPU: multi tests + benches with the merged update
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_multi.py -q -m gpu -s -p no:cacheprovider > gpurun_out/pytest_multi.log 2>&1
echo "pytest multi exit $?" >> gpurun_out/pytest_multi.log; tail -4 gpurun_out/pytest_multi.log | cut -c1-300
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29541"
timeout 900 $TR bench.py --gpus 2 --config 4 --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/bench_c4_n2.json 2> gpurun_out/bench_c4_n2.err
echo "bench c4 n2 exit $?"; grep '^{' gpurun_out/bench_c4_n2.json | cut -c1-900; grep -E "Error|error" gpurun_out/bench_c4_n2.err | tail -3 | cut -c1-400
timeout 1500 $TR bench.py --gpus 2 --config 5 --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/bench_c5_n2.json 2> gpurun_out/bench_c5_n2.err
echo "bench c5 n2 exit $?"; grep '^{' gpurun_out/bench_c5_n2.json | cut -c1-900; grep -E "Error|error" gpurun_out/bench_c5_n2.err | tail -3 | cut -c1-400
JAICOV_PANEL_TILES=8 timeout 1500 $TR bench.py --gpus 2 --config 5 --steps 2 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/bench_c5_n2_pw8.json 2> gpurun_out/bench_c5_n2_pw8.err
echo "bench c5 n2 pw8 exit $?"; grep '^{' gpurun_out/bench_c5_n2_pw8.json | cut -c1-900
ls gpurun_out
