#!/bin/bash
# 2-GPU: full gpu suite (incl. multi + writers), N=2 bench config 5, ncu source capture of the diagonal kernel
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -q -m gpu -s --maxfail=40 -p no:cacheprovider > gpurun_out/pytest_gpu.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_gpu.log; tail -4 gpurun_out/pytest_gpu.log | cut -c1-300
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29541"
timeout 1500 $TR bench.py --gpus 2 --config 5 --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/bench_c5_n2.json 2> gpurun_out/bench_c5_n2.err
echo "bench c5 n2 exit $?"; grep '^{' gpurun_out/bench_c5_n2.json | cut -c1-1000; grep -E "Error|error" gpurun_out/bench_c5_n2.err | tail -3 | cut -c1-400
python bench.py --config 2 --steps 1 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/plain_c2.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:'k_potrf_diag' --launch-skip 20 -c 1 -o gpurun_out/prof_diag \
    python bench.py --config 2 --steps 1 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/ncu_diag.log 2>&1
echo "ncu diag exit $?"; tail -2 gpurun_out/ncu_diag.log
ls gpurun_out
