#!/bin/bash
# 1-GPU: launch list of the HBM-bound kernels (assembly, Omega, scaling, epilogue) at config 5
mkdir -p gpurun_out
python bench.py --config 5 --steps 1 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/plain_c5.log 2>&1 && \
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none \
    -k regex:'k_by_image|k_by_point|k_omega|k_pose|k_scale_system|k_qxx_epilogue|k_point_scatter|k_image_finalize|k_camera|k_precond|k_build_rhs|k_datum|k_update|k_solve' \
    --launch-skip 3000 -c 1100 --csv --log-file gpurun_out/launches_c5_hbm.csv \
    python bench.py --config 5 --steps 1 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/ncu_c5.log 2>&1
echo "ncu c5 exit $?"; tail -2 gpurun_out/ncu_c5.log | cut -c1-300
