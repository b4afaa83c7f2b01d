timeout 900 python -m pytest tests/test_gpu_multi.py -m gpu -q -x -s -k "cfg4mid-dense-1 or cfg2-dense-1 or cfg4-dense-8 or example-auto-8 or cfg3 or single_process" > gpurun_out/multi2_oo2.log 2>&1; echo "multi exit $?"; grep -E "passed|failed|Error" gpurun_out/multi2_oo2.log | tail -3 | cut -c1-400
bash tools/gpu_job.sh bench 2 c5_n2 --steps 2 --warmup 3 --no-cpu-baseline | cut -c1-200
python - <<'PY'
import json
j=json.load(open('gpurun_out/bench_c5_n2.json'))
print('N=2: ms', j['ms_per_step'], 'frac', j['roofline']['frac'], 'check', j['check']['ok'], j['check']['cofactor_residual'], j['check']['solve_residual'], 'stages', {k: round(v,1) for k,v in j['config']['stage_ms'].items()})
print('dmma', j['fp64_dmma']['ms_per_step'], j['fp64_dmma']['check']['ok'], 'structured', j['structured']['ms_per_step'], j['structured']['check']['ok'], 'e2e', j['e2e'])
PY
