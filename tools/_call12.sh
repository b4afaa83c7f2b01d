timeout 1800 python -m pytest tests/test_gpu_multi.py -m gpu -q -x -s > gpurun_out/multi2_oo.log 2>&1; echo "multi exit $?"; grep -E "passed|failed|Error|error" gpurun_out/multi2_oo.log | tail -5 | cut -c1-500
grep -E "^\.?\{" gpurun_out/multi2_oo.log | python -c "
import sys, ast
for l in sys.stdin:
    d = ast.literal_eval(l.lstrip('.').strip())
    print(d['scene'], d['solver_used'], 'pt', d['panel_tiles'], 'q', '%.2e' % d['qxx_scaled_err'], 'x', '%.2e' % d['param_rel_err'], 'verify', {k: '%.1e' % v for k, v in d['verify'].items()}, 'ms', round(d['ms_last_pass'], 2))
"
bash tools/gpu_job.sh bench 2 c5_n2 --steps 2 --warmup 3 --no-cpu-baseline | cut -c1-200
python - <<'PY'
import json
j=json.load(open('gpurun_out/bench_c5_n2.json'))
print('N=2: ms', j['ms_per_step'], 'frac', j['roofline']['frac'], 'check', j['check']['ok'], j['check']['cofactor_residual'], j['check']['solve_residual'], 'stages', {k: round(v,1) for k,v in j['config']['stage_ms'].items()})
print('dmma', j['fp64_dmma']['ms_per_step'], j['fp64_dmma']['check']['ok'], 'structured', j['structured']['ms_per_step'], j['structured']['check']['ok'], 'e2e', j['e2e'])
PY
nvidia-smi --query-gpu=memory.used --format=csv | head -3
