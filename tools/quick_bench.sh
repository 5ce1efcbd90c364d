#!/bin/bash
# usage (on the GPU box): bash tools/quick_bench.sh [config] [extra bench args]   -> one line per filter mode
cfg=${1:-C2}; shift
mkdir -p gpurun_out
for f in exact hw hybrid; do
  python bench.py --steps 10 --warmup 3 --filter $f --config $cfg --no-cpu-baseline "$@" > gpurun_out/bench_$f.json 2> gpurun_out/bench_$f.err
  python -c "
import json;d=json.load(open('gpurun_out/bench_$f.json'));print('$f', '$cfg', 'ms/frame %.3f'%d['ms_per_frame'], 'Mpix/s %.1f'%d['value'], 'e2e ms %.3f'%d['e2e']['ms_per_frame'], d['clocks'])"
  tail -3 gpurun_out/bench_$f.err
done
