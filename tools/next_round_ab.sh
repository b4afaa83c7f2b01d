#!/bin/bash
# A/B experiments prepared for the next round; run under gpurun, results into gpurun_out/.
#   bash tools/next_round_ab.sh tile      1 GPU : band-swizzled tile order of the lower-triangular GEMM launches (JAICOV_TILE_BAND)
#   bash tools/next_round_ab.sh panel N   N GPUs: panel width of the distributed Cholesky (JAICOV_PANEL_TILES)
mkdir -p gpurun_out
case "$1" in
tile)
  for band in 0 4 8 16; do
    for solver in structured dense; do
      JAICOV_TILE_BAND=$band JAICOV_SOLVER=$solver python bench.py --config 5 --steps 2 --warmup 3 --no-e2e --no-cpu-baseline \
        > gpurun_out/ab_tile_${solver}_band$band.log 2>&1
      echo "band=$band $solver: $(grep -o '"ms_per_step": [0-9.]*' gpurun_out/ab_tile_${solver}_band$band.log) $(grep -o '"frac": [0-9.]*' gpurun_out/ab_tile_${solver}_band$band.log)"
    done
  done ;;
panel)
  N=${2:-8}
  for pw in 4 8 16; do
    JAICOV_PANEL_TILES=$pw python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29561 \
      bench.py --gpus $N --steps 2 --warmup 3 --no-e2e --no-structured > gpurun_out/ab_panel_n${N}_pw$pw.log 2>&1
    echo "N=$N panel_tiles=$pw: $(grep -o '"ms_per_step": [0-9.]*' gpurun_out/ab_panel_n${N}_pw$pw.log) $(grep -o '"stage_ms": {[^}]*}' gpurun_out/ab_panel_n${N}_pw$pw.log)"
  done ;;
*) echo "usage: $0 tile | panel N" ;;
esac
