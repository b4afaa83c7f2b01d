bash tools/gpu_job.sh tests -s > /dev/null; tail -3 gpurun_out/pytest_gpu.log
OZ="JAICOV_GEMM_OZAKI=8 JAICOV_OZAKI_MIN_TILES=1 JAICOV_OZAKI_MIN_K=128"
env $OZ timeout 1500 python -m pytest tests/test_gpu_parity.py tests/test_gpu_fullsize.py -m gpu -q -s > gpurun_out/ozaki_parity.log 2>&1; echo "ozaki parity exit $?"; tail -4 gpurun_out/ozaki_parity.log | cut -c1-300; grep -E "full size" gpurun_out/ozaki_parity.log | cut -c1-300
for cfg in 4 5; do for sv in dense structured; do
  echo "FP64  $(timeout 600 python tools/one_pass.py $cfg $sv 3 2>&1 | tail -1)"
  echo "OZ8   $(JAICOV_GEMM_OZAKI=8 timeout 600 python tools/one_pass.py $cfg $sv 3 2>&1 | tail -1)"
  echo "OZ7   $(JAICOV_GEMM_OZAKI=7 timeout 600 python tools/one_pass.py $cfg $sv 3 2>&1 | tail -1)"
done; done 2>&1 | tee gpurun_out/ozaki_ab.log
timeout 900 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --kernel-name-base demangled -k regex:'k_gemm<1, 1' --csv --log-file gpurun_out/lauum_c5_traffic.csv python tools/one_pass.py 5 dense 1 > gpurun_out/lauum_c5_traffic.log 2>&1; echo "ncu lauum exit $?"; grep -c k_gemm gpurun_out/lauum_c5_traffic.csv; tail -4 gpurun_out/lauum_c5_traffic.csv | cut -c1-400
JAICOV_TILE_BAND=8 timeout 900 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --kernel-name-base demangled -k regex:'k_gemm<1, 1' --csv --log-file gpurun_out/lauum_c5_traffic_band8.csv python tools/one_pass.py 5 dense 1 > gpurun_out/lauum_c5_traffic_band8.log 2>&1; tail -3 gpurun_out/lauum_c5_traffic_band8.csv | cut -c1-400
JAICOV_GEMM_OZAKI=8 timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_gemm_oz -c 1 -f -o gpurun_out/ncu_full_k_gemm_oz python tools/ozaki_gpu_check.py --worker time > gpurun_out/ncu_oz.log 2>&1; echo "ncu oz exit $?"
bash tools/gpu_job.sh bench 1 c5_n1 --steps 3 --warmup 3 | cut -c1-600
