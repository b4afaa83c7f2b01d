bash tools/gpu_job.sh bench 8 c5_n8 --steps 3 --warmup 3 --no-cpu-baseline | cut -c1-200
python - <<'PY'
import json
j=json.load(open('gpurun_out/bench_c5_n8.json'))
print('N=8 owner-only: ms', j['ms_per_step'], 'frac', j['roofline']['frac'], 'check', j['check']['ok'], j['check']['cofactor_residual'], j['check']['solve_residual'], 'stages', {k: round(v,1) for k,v in j['config']['stage_ms'].items()})
print('dmma', j['fp64_dmma']['ms_per_step'], j['fp64_dmma']['frac_of_fp64_peak'], j['fp64_dmma']['check']['ok'])
print('structured', j['structured']['ms_per_step'], j['structured']['check']['ok'])
print('e2e', j['e2e'], 'clocks', j['clocks'])
PY
nvidia-smi --query-gpu=memory.used --format=csv,noheader | head -8 | tr '\n' ' '; echo
JAICOV_DIST_STORAGE=replica bash tools/gpu_job.sh bench 8 c5_n8_replica --steps 3 --warmup 3 --no-cpu-baseline --no-e2e --no-structured --no-dmma | cut -c1-100
python - <<'PY'
import json
j=json.load(open('gpurun_out/bench_c5_n8_replica.json'))
print('N=8 replica: ms', j['ms_per_step'], 'check', j['check']['ok'], 'stages', {k: round(v,1) for k,v in j['config']['stage_ms'].items()})
PY
