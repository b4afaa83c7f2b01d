#!/bin/bash
# The one parametrised GPU runner (replaces round 1's scratch scripts): `gpurun -- bash tools/gpu_job.sh <job> [args]`.
# Every job writes into gpurun_out/ (merged back by gpurun); risky steps run under their own timeout.
mkdir -p gpurun_out
TR() { local n=$1; shift; python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port ${PORT:-29551} "$@"; }
job=$1; shift
case "$job" in
tests)      # the whole -m gpu suite + smoke
  timeout 2400 python -m pytest tests -m gpu -q -x "$@" > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?"; tail -5 gpurun_out/pytest_gpu.log
  python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3 ;;
pytest)     # selected tests: bash tools/gpu_job.sh pytest <log name> <pytest args...>
  name=$1; shift
  timeout 2400 python -m pytest "$@" > gpurun_out/$name.log 2>&1; echo "pytest exit $?"; tail -15 gpurun_out/$name.log ;;
bench)      # bash tools/gpu_job.sh bench <N> <tag> [bench args]: one bench line at N GPUs
  N=$1; tag=$2; shift 2
  if [ "$N" = "1" ]; then timeout 1500 python bench.py "$@" > gpurun_out/bench_$tag.log 2>&1
  else timeout 1500 bash -c "$(declare -f TR); TR $N bench.py --gpus $N $*" > gpurun_out/bench_$tag.log 2>&1; fi
  echo "bench $tag exit $?"; grep '^{' gpurun_out/bench_$tag.log | tail -1 > gpurun_out/bench_$tag.json; cut -c1-1800 gpurun_out/bench_$tag.json
  grep -v '^{' gpurun_out/bench_$tag.log | grep -iE "error|assert|Traceback" | tail -5 | cut -c1-300 ;;
multi)      # bash tools/gpu_job.sh multi <N> <scene> [env...]: multi-GPU parity worker (tests/multi_worker.py)
  N=$1; scene=$2; shift 2
  env "$@" timeout 900 bash -c "$(declare -f TR); TR $N tests/multi_worker.py $scene" > gpurun_out/multi_${scene}_n$N.log 2>&1
  echo "multi $scene N=$N exit $?"; grep '^{' gpurun_out/multi_${scene}_n$N.log | tail -1 | cut -c1-900 ;;
ozaki)      # int8-digit tcgen05 GEMM: correctness under timeouts, then timings (tools/ozaki_gpu_check.py)
  timeout 1500 python tools/ozaki_gpu_check.py "$@" > gpurun_out/ozaki_check.log 2>&1; cat gpurun_out/ozaki_check.log | cut -c1-1200 ;;
tileband)   # band-swizzled tile order of the lower-triangular launches: bash tools/gpu_job.sh tileband <config> <solver> <bands...>
  cfg=$1; solver=$2; shift 2
  for band in "$@"; do
    JAICOV_TILE_BAND=$band JAICOV_SOLVER=$solver timeout 900 python bench.py --config $cfg --steps 2 --warmup 3 --no-e2e --no-cpu-baseline \
      > gpurun_out/ab_tile_c${cfg}_${solver}_band$band.log 2>&1
    echo "config $cfg $solver band=$band: $(grep -o '"ms_per_step": [0-9.]*' gpurun_out/ab_tile_c${cfg}_${solver}_band$band.log | head -1) $(grep -o '"stage_ms": {[^}]*}' gpurun_out/ab_tile_c${cfg}_${solver}_band$band.log | head -1)"
  done ;;
ncu_full)   # bash tools/gpu_job.sh ncu_full <tag> <kernel regex> <count> <command...>: plain run first, then the full capture
  tag=$1; rx=$2; cnt=$3; shift 3
  timeout 900 "$@" > gpurun_out/ncu_${tag}_plain.log 2>&1 || { echo "plain run failed"; tail -5 gpurun_out/ncu_${tag}_plain.log; exit 0; }
  timeout 1500 ncu --set full --clock-control none --import-source on -k regex:"$rx" -c $cnt -f -o gpurun_out/ncu_full_$tag "$@" > gpurun_out/ncu_${tag}.log 2>&1
  echo "ncu $tag exit $?"; tail -3 gpurun_out/ncu_${tag}.log ;;
ncu_list)   # bash tools/gpu_job.sh ncu_list <tag> <count> <command...>: launch list with durations and DRAM bytes
  tag=$1; cnt=$2; shift 2
  timeout 900 "$@" > gpurun_out/list_${tag}_plain.log 2>&1 || { echo "plain run failed"; tail -5 gpurun_out/list_${tag}_plain.log; exit 0; }
  timeout 1500 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c $cnt --csv \
    --log-file gpurun_out/launches_$tag.csv "$@" > gpurun_out/list_${tag}.log 2>&1
  echo "ncu list $tag exit $?"; python tools/summarize_launches.py gpurun_out/launches_$tag.csv 2>/dev/null | head -40 ;;
*) echo "usage: $0 tests | pytest | bench | multi | ozaki | tileband | ncu_full | ncu_list" ;;
esac
