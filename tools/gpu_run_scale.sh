#!/bin/bash
# bench.py at N GPUs (argument), default workload (config 5): dense headline + structured route + e2e of both
N=${1:-8}
mkdir -p gpurun_out
if [ "$N" = "1" ]; then
  python bench.py --steps 3 --warmup 3 > gpurun_out/scale_n$N.log 2>&1
else
  python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29551 bench.py --gpus $N --steps 3 --warmup 3 > gpurun_out/scale_n$N.log 2>&1
fi
echo "N=$N exit $?"
grep '^{' gpurun_out/scale_n$N.log | tail -1 > gpurun_out/scale_n$N.json
python - <<PY
import json
j=json.load(open('gpurun_out/scale_n$N.json'))
print('dense ms', round(j['ms_per_step'],1), 'frac', round(j['roofline']['frac'],3), 'TF', round(j['roofline']['achieved'],1), 'e2e', round(j['e2e']['value'],4), j['e2e']['phase_ms'])
print({k: round(v,1) for k,v in j['config']['stage_ms'].items()})
s=j.get('structured')
if s: print('structured', {k:(round(v,3) if isinstance(v,float) else v) for k,v in s.items() if k not in ('note',)})
print(j['clocks'])
PY
grep -v '^{' gpurun_out/scale_n$N.log | grep -iv "OMP_NUM\|\*\*\*\*\|NCCL version" | tail -5 | cut -c1-300
