#!/bin/bash
# 2 GPUs: multi-GPU parity (both routes), then config 5 on 2 GPUs, dense and structured
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_multi.py -m gpu -q -s > gpurun_out/pytest_r.log 2>&1; echo "pytest exit $?"; grep -E "passed|failed|error" gpurun_out/pytest_r.log | tail -3
grep -E "^\{'scene|FAILED|Error|error" gpurun_out/pytest_r.log | cut -c1-400 | head -20
for sv in structured dense; do
  python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus 2 --steps 2 --warmup 3 --solver $sv > gpurun_out/r_c5_n2_$sv.log 2>&1
  echo "c5 N=2 $sv: exit $? $(grep -o '"ms_per_step": [0-9.]*' gpurun_out/r_c5_n2_$sv.log) $(grep -o '"stage_ms": {[^}]*}' gpurun_out/r_c5_n2_$sv.log) $(grep -o '"e2e": {"value": [0-9.]*' gpurun_out/r_c5_n2_$sv.log)"
  grep -v '^{' gpurun_out/r_c5_n2_$sv.log | tail -3 | cut -c1-300
done
