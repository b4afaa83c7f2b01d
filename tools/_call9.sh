python tools/gemm_k_sweep.py 2>&1 | tee gpurun_out/gemm_k_sweep.log
for pw in 8 16 32; do
  JAICOV_PANEL_TILES=$pw bash tools/gpu_job.sh bench 4 c5_n4_pw$pw --steps 2 --warmup 3 --no-e2e --no-structured --no-dmma --no-cpu-baseline --no-check > /dev/null
  python - <<PY
import json
j=json.load(open('gpurun_out/bench_c5_n4_pw$pw.json'))
print('N=4 panel_tiles=$pw: ms', round(j['ms_per_step'],1), {k: round(v,1) for k,v in j['config']['stage_ms'].items()})
PY
done
