#!/bin/bash
# Strong-scaling run of the headline workload on the GPUs of one box: bench.py at N = 1,2,4,...,$1
# plus the multi-GPU parity check.  Usage (under gpurun --gpus G): bash tools/scale_run.sh G [workload]
G=${1:-2}; WL=${2:-c3}; OUT=gpurun_out; mkdir -p $OUT
for N in 1 2 4 8; do
  [ $N -gt $G ] && break
  if [ $N -eq 1 ]; then
    timeout 300 python bench.py --gpus 1 --workload $WL --steps 5 --warmup 3 --no-cpu-baseline > $OUT/scale_${WL}_g1.json 2> $OUT/scale_${WL}_g1.err
  else
    timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((29600+N)) \
      bench.py --gpus $N --workload $WL --steps 5 --warmup 3 > $OUT/scale_${WL}_g$N.json 2> $OUT/scale_${WL}_g$N.err
  fi
  python - <<PY
import json
try:
    d=json.load(open("$OUT/scale_${WL}_g$N.json")); print("N=$N", d["value"], d["ms_per_step"], d["e2e"]["value"], d["clocks"])
except Exception as ex:
    print("N=$N failed", ex)
PY
done
