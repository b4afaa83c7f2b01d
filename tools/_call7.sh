timeout 900 python -m pytest tests/test_dense_image_sigma.py -m gpu -q -s > gpurun_out/dense_sigma.log 2>&1; echo "dense sigma exit $?"; grep -E "dense image|passed|failed" gpurun_out/dense_sigma.log | cut -c1-300
bash tools/gpu_job.sh ozaki > /dev/null; grep -E '"stage"' gpurun_out/ozaki_check.log | cut -c1-700
for sv in dense structured; do
  echo "OZ8 band8 $(JAICOV_GEMM_OZAKI=8 timeout 600 python tools/one_pass.py 5 $sv 3 2>&1 | tail -1)"
  echo "OZ8 band0 $(JAICOV_GEMM_OZAKI=8 JAICOV_OZAKI_BAND=0 timeout 600 python tools/one_pass.py 5 $sv 3 2>&1 | tail -1)"
  echo "OZ8 band16 $(JAICOV_GEMM_OZAKI=8 JAICOV_OZAKI_BAND=16 timeout 600 python tools/one_pass.py 5 $sv 3 2>&1 | tail -1)"
done 2>&1 | tee gpurun_out/ozaki_band_ab.log
echo "OZ8 band8 config4 dense $(JAICOV_GEMM_OZAKI=8 timeout 600 python tools/one_pass.py 4 dense 3 2>&1 | tail -1)" | tee -a gpurun_out/ozaki_band_ab.log
JAICOV_GEMM_OZAKI=8 timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_gemm_oz -c 1 -f -o gpurun_out/ncu_full_k_gemm_oz_band8 python tools/ozaki_gpu_check.py --worker time > gpurun_out/ncu_oz.log 2>&1; echo "ncu oz exit $?"
