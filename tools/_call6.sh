timeout 900 python -m pytest tests/test_dense_image_sigma.py -m gpu -q -s > gpurun_out/dense_sigma.log 2>&1; echo "dense sigma exit $?"; tail -12 gpurun_out/dense_sigma.log | cut -c1-400
OZ="JAICOV_GEMM_OZAKI=8 JAICOV_OZAKI_MIN_TILES=1 JAICOV_OZAKI_MIN_K=128"
env $OZ timeout 1500 python -m pytest tests/test_gpu_multi.py -m gpu -q -x -s > gpurun_out/ozaki_multi2.log 2>&1; echo "ozaki multi exit $?"; tail -5 gpurun_out/ozaki_multi2.log | cut -c1-400
JAICOV_GEMM_OZAKI=8 bash tools/gpu_job.sh bench 2 c5_n2_oz8 --steps 2 --warmup 3 --no-cpu-baseline | cut -c1-400
python - <<'PY'
import json
j=json.load(open('gpurun_out/bench_c5_n2_oz8.json'))
print('N=2 OZ8: ms', j['ms_per_step'], 'check', {k:j['check'][k] for k in ('solve_residual','datum_residual','cofactor_residual','omega_rel_diff','ok')}, 'structured ms', j['structured']['ms_per_step'], j['structured']['check']['ok'], 'e2e', j['e2e']['value'])
PY
timeout 900 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --kernel-name-base demangled -k regex:'k_gemm<.int.1, .int.1' --csv --log-file gpurun_out/lauum_c5_traffic.csv python tools/one_pass.py 5 dense 1 > gpurun_out/lauum_c5_traffic.log 2>&1; echo "ncu lauum exit $?"; python tools/summarize_launches.py gpurun_out/lauum_c5_traffic.csv
JAICOV_TILE_BAND=8 timeout 900 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --kernel-name-base demangled -k regex:'k_gemm<.int.1, .int.1' --csv --log-file gpurun_out/lauum_c5_traffic_band8.csv python tools/one_pass.py 5 dense 1 > gpurun_out/lauum_c5_traffic_band8.log 2>&1; python tools/summarize_launches.py gpurun_out/lauum_c5_traffic_band8.csv
