export JAICOV_TEST_WORLD=8
timeout 900 python -m pytest tests/test_gpu_multi.py -m gpu -q -x -s -k "cfg4mid-dense-1 or single_process and cfg2-dense-1 or single_process and example" > gpurun_out/multi8.log 2>&1; echo "multi8 exit $?"; grep -E "^\.?\{|passed|failed" gpurun_out/multi8.log | cut -c1-700
bash tools/gpu_job.sh bench 8 c5_n8 --steps 3 --warmup 3 --no-cpu-baseline | cut -c1-200
python - <<'PY'
import json
j=json.load(open('gpurun_out/bench_c5_n8.json'))
print('N=8: ms', j['ms_per_step'], 'frac', j['roofline']['frac'], 'check', j['check']['ok'], j['check']['cofactor_residual'], j['check']['solve_residual'], 'stages', {k: round(v,1) for k,v in j['config']['stage_ms'].items()})
print('dmma', j['fp64_dmma']['ms_per_step'], j['fp64_dmma']['frac_of_fp64_peak'], j['fp64_dmma']['check']['ok'])
print('structured', j['structured']['ms_per_step'], j['structured']['check']['ok'], j['structured']['stage_ms'])
print('e2e', j['e2e'], 'clocks', j['clocks'])
PY
