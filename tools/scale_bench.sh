#!/bin/bash
# usage (on an N-GPU box): bash tools/scale_bench.sh "1 2 4 8" C2 [extra args]  -> gpurun_out/scale_<cfg>_<n>.json
ns=${1:-"1 2"}; cfg=${2:-C2}; shift; shift
mkdir -p gpurun_out
for n in $ns; do
  if [ "$n" = "1" ]; then python bench.py --gpus 1 --steps 20 --warmup 3 --config $cfg --no-cpu-baseline "$@" > gpurun_out/scale_${cfg}_$n.json 2> gpurun_out/scale_${cfg}_$n.err
  else python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29500+n)) bench.py --gpus $n --steps 20 --warmup 3 --config $cfg --no-cpu-baseline "$@" > gpurun_out/scale_${cfg}_$n.json 2> gpurun_out/scale_${cfg}_$n.err; fi
  python -c "
import json
l=[x for x in open('gpurun_out/scale_${cfg}_$n.json') if x.startswith('{')]
d=json.loads(l[-1]); print('$cfg', 'n=$n', 'ms/frame %.3f'%d['ms_per_frame'], 'Mpix/s %.1f'%d['value'], 'e2e ms %.3f'%d['e2e']['ms_per_frame'])" || tail -5 gpurun_out/scale_${cfg}_$n.err
done
