#!/bin/bash
# A/B experiments prepared for the next round; run under gpurun, results into gpurun_out/.
#   bash tools/next_round_ab.sh tile      1 GPU : band-swizzled tile order of the lower-triangular GEMM launches (JAICOV_TILE_BAND)
#   bash tools/next_round_ab.sh panel N   N GPUs: panel width of the distributed Cholesky (JAICOV_PANEL_TILES)
#   bash tools/next_round_ab.sh ozaki     1 GPU : first run of the int8-digit tcgen05 GEMM (csrc/ozaki.cu, never executed so far):
#                                                 correctness on small tile grids and SPD systems under timeouts, then timings;
#                                                 only if all of that is green: config 4 / 5 passes with JAICOV_GEMM_OZAKI=8
#   bash tools/next_round_ab.sh ozaki_multi N  N GPUs: the same switch through the multi-GPU parity tests and the config-5 bench
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
ozaki)
  timeout 1500 python tools/ozaki_gpu_check.py > gpurun_out/ozaki_check.log 2>&1
  cat gpurun_out/ozaki_check.log
  if grep -q '"ok": false' gpurun_out/ozaki_check.log; then echo "ozaki: checks not green, skipping the bench runs"; exit 0; fi
  for cfg in 4 5; do
    for solver in dense structured; do
      JAICOV_GEMM_OZAKI=8 JAICOV_SOLVER=$solver timeout 900 python bench.py --config $cfg --steps 2 --warmup 3 --no-e2e --no-cpu-baseline \
        > gpurun_out/ab_ozaki_c${cfg}_$solver.log 2>&1
      echo "config $cfg $solver, 8 digits: $(grep -o '"ms_per_step": [0-9.]*' gpurun_out/ab_ozaki_c${cfg}_$solver.log) $(grep -o '"frac": [0-9.]*' gpurun_out/ab_ozaki_c${cfg}_$solver.log)"
    done
  done
  # the parity networks are small (launches of a few tiles): send EVERY launch through the digit path for this run
  JAICOV_GEMM_OZAKI=8 JAICOV_OZAKI_MIN_TILES=1 JAICOV_OZAKI_MIN_K=128 timeout 1200 python -m pytest tests/test_gpu_parity.py -x -q -m gpu > gpurun_out/ozaki_parity.log 2>&1; tail -3 gpurun_out/ozaki_parity.log ;;
ozaki_ncu)
  # after `ozaki` is green: launch list and one full capture of the digit-product tile kernel on the timing worker
  # (the plain run directly before, as the profiling recipe asks)
  JAICOV_GEMM_OZAKI=8 python tools/ozaki_gpu_check.py --worker time > gpurun_out/ozaki_time_plain.log 2>&1 &&
  JAICOV_GEMM_OZAKI=8 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 200 --csv \
    --log-file gpurun_out/launches_ozaki_time.csv python tools/ozaki_gpu_check.py --worker time > gpurun_out/ozaki_time_ncu.log 2>&1
  JAICOV_GEMM_OZAKI=8 ncu --set full --clock-control none --import-source on -k regex:k_gemm_oz -c 1 \
    -o gpurun_out/ncu_full_k_gemm_oz python tools/ozaki_gpu_check.py --worker time >> gpurun_out/ozaki_time_ncu.log 2>&1
  tail -5 gpurun_out/ozaki_time_ncu.log ;;
ozaki_multi)
  # N GPUs (after `ozaki` is green on one): multi-GPU parity with every launch on the digit path, then the config-5 bench
  N=${2:-2}
  JAICOV_GEMM_OZAKI=8 JAICOV_OZAKI_MIN_TILES=1 JAICOV_OZAKI_MIN_K=128 timeout 900 python -m pytest tests/test_gpu_multi.py -x -q -m gpu \
    > gpurun_out/ozaki_multi_parity_n$N.log 2>&1; tail -3 gpurun_out/ozaki_multi_parity_n$N.log
  JAICOV_GEMM_OZAKI=8 timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29562 \
    bench.py --gpus $N --steps 2 --warmup 3 --no-e2e > gpurun_out/ab_ozaki_c5_n$N.log 2>&1
  echo "config 5, $N GPUs, 8 digits: $(grep -o '"ms_per_step": [0-9.]*' gpurun_out/ab_ozaki_c5_n$N.log) $(grep -o '"frac": [0-9.]*' gpurun_out/ab_ozaki_c5_n$N.log)" ;;
*) echo "usage: $0 tile | panel N | ozaki | ozaki_ncu | ozaki_multi N" ;;
esac
