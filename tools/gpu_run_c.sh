#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -q -m gpu -s --maxfail=40 -p no:cacheprovider > gpurun_out/pytest_gpu.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
tail -4 gpurun_out/pytest_gpu.log
timeout 600 python tools/microbench.py 2048 8192 16384 > gpurun_out/microbench.log 2>&1; tail -4 gpurun_out/microbench.log
timeout 900 python bench.py --config 4 --steps 3 --warmup 3 > gpurun_out/bench_c4.json 2> gpurun_out/bench_c4.err
echo "bench c4 exit $?"; cat gpurun_out/bench_c4.json; tail -3 gpurun_out/bench_c4.err
timeout 600 python bench.py --config 2 --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_c2.json 2> gpurun_out/bench_c2.err
echo "bench c2 exit $?"; cat gpurun_out/bench_c2.json
timeout 1500 python bench.py --config 5 --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/bench_c5.json 2> gpurun_out/bench_c5.err
echo "bench c5 exit $?"; cat gpurun_out/bench_c5.json; tail -5 gpurun_out/bench_c5.err
# full capture of the single-launch contraction (LAUUM, k_gemm<1,1>) at config 4
python bench.py --config 4 --steps 1 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/plain_c4.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:'k_gemm' --launch-skip 1132 -c 3 -o gpurun_out/prof_gemm \
    python bench.py --config 4 --steps 1 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/ncu_full.log 2>&1
echo "ncu full exit $?"; tail -3 gpurun_out/ncu_full.log
ls -la gpurun_out
