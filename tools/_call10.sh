python tools/gemm_k_sweep.py 2>&1 | tee gpurun_out/gemm_k_sweep.log
for mk in 1024 4096; do
JAICOV_OZAKI_MIN_K=$mk JAICOV_PANEL_TILES=16 bash tools/gpu_job.sh bench 2 c5_n2_mk$mk --steps 2 --warmup 3 --no-e2e --no-structured --no-dmma --no-cpu-baseline --no-check > /dev/null
python - <<PY
import json
j=json.load(open('gpurun_out/bench_c5_n2_mk$mk.json'))
print('N=2 pw16 min_k=$mk: ms', round(j['ms_per_step'],1), {k: round(v,1) for k,v in j['config']['stage_ms'].items()})
PY
done
JAICOV_GEMM_OZAKI=0 JAICOV_PANEL_TILES=16 bash tools/gpu_job.sh bench 2 c5_n2_dmma16 --steps 2 --warmup 3 --no-e2e --no-structured --no-dmma --no-cpu-baseline --no-check > /dev/null
python - <<PY
import json
j=json.load(open('gpurun_out/bench_c5_n2_dmma16.json'))
print('N=2 pw16 DMMA: ms', round(j['ms_per_step'],1), {k: round(v,1) for k,v in j['config']['stage_ms'].items()})
PY
