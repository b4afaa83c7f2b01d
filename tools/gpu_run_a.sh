#!/bin/bash
# first GPU contact: microbenchmarks, GPU parity tests, smoke
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw,memory.total --format=csv > gpurun_out/smi.txt 2>&1
nproc > gpurun_out/host.txt; lscpu | grep -E "Model name|^CPU\(s\)" >> gpurun_out/host.txt; free -g >> gpurun_out/host.txt
timeout 600 python tools/microbench.py 2048 4096 8192 > gpurun_out/microbench.log 2>&1
echo "microbench exit $?" >> gpurun_out/microbench.log
timeout 1500 python -m pytest tests -q -m gpu -s --maxfail=40 -p no:cacheprovider > gpurun_out/pytest_gpu.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1
echo "smoke exit $?" >> gpurun_out/smoke.log
tail -5 gpurun_out/microbench.log; tail -30 gpurun_out/pytest_gpu.log; tail -3 gpurun_out/smoke.log
