bash tools/gpu_job.sh tests -s > /dev/null; tail -3 gpurun_out/pytest_gpu.log | cut -c1-200; grep -E "full size" gpurun_out/pytest_gpu.log | cut -c1-260
for sv in dense structured; do
  echo "OZ8 cluster2 $(JAICOV_OZAKI_CLUSTER=2 timeout 600 python tools/one_pass.py 5 $sv 3 2>&1 | tail -1)"
  echo "OZ8 cluster1 $(timeout 600 python tools/one_pass.py 5 $sv 3 2>&1 | tail -1)"
done 2>&1 | tee gpurun_out/ozaki_cluster_ab.log
bash tools/gpu_job.sh bench 1 c5_n1 --steps 3 --warmup 3 | cut -c1-300
python - <<'PY'
import json
j=json.load(open('gpurun_out/bench_c5_n1.json'))
print('N=1: ms', j['ms_per_step'], 'frac', j['roofline']['frac'], 'check', j['check']['ok'], j['check']['cofactor_residual'], 'dmma', j['fp64_dmma']['ms_per_step'], j['fp64_dmma']['frac_of_fp64_peak'], j['fp64_dmma']['check']['ok'])
print('structured', j['structured']['ms_per_step'], j['structured']['check']['ok'], 'e2e', j['e2e'], 'launches', j['gpu_launches'])
print(j['other_configs'])
PY
