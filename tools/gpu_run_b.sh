#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -q -m gpu -s --maxfail=40 -p no:cacheprovider > gpurun_out/pytest_gpu.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
tail -4 gpurun_out/pytest_gpu.log
timeout 900 python bench.py --config 4 --steps 3 --warmup 3 > gpurun_out/bench_c4.json 2> gpurun_out/bench_c4.err
echo "bench c4 exit $?"; cat gpurun_out/bench_c4.json; tail -3 gpurun_out/bench_c4.err
timeout 600 python bench.py --config 2 --steps 5 --warmup 3 > gpurun_out/bench_c2.json 2> gpurun_out/bench_c2.err
echo "bench c2 exit $?"; cat gpurun_out/bench_c2.json
# launch list (config 2: ~230 launches per pass)
python bench.py --config 2 --steps 1 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/plain_c2.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 1400 --csv --log-file gpurun_out/launches_c2.csv \
    python bench.py --config 2 --steps 1 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/ncu_c2.log 2>&1
echo "ncu c2 exit $?"
# launch list config 4, last pass only (skip ~3 passes x ~2100 launches)
python bench.py --config 4 --steps 1 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/plain_c4.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -s 6400 -c 2300 --csv --log-file gpurun_out/launches_c4.csv \
    python bench.py --config 4 --steps 1 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/ncu_c4.log 2>&1
echo "ncu c4 list exit $?"
# full capture of the big single-launch contraction (LAUUM, k_gemm<1,1>) at config 4
ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k regex:'k_gemm<1, 1>' -c 1 -o gpurun_out/prof_gemm_lauum \
    python bench.py --config 4 --steps 1 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/ncu_full.log 2>&1
echo "ncu full exit $?"; tail -3 gpurun_out/ncu_full.log
ls -la gpurun_out
