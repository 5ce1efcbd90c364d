#!/bin/bash
# usage (on the GPU box): bash tools/variant_bench.sh "C2 C3" -> one line per library variant under project-marshmallow_b200/variants/
cfgs=${1:-"C2"}; shift
mkdir -p gpurun_out
for lib in project-marshmallow_b200/variants/*.so; do
  for cfg in $cfgs; do
    MM_LIBRARY=$PWD/$lib python bench.py --steps 10 --warmup 3 --config $cfg --no-cpu-baseline "$@" > gpurun_out/variant.json 2> gpurun_out/variant.err
    python -c "
import json;d=json.load(open('gpurun_out/variant.json'));print('$(basename $lib)', '$cfg', 'ms/frame %.3f'%d['ms_per_frame'], 'Mpix/s %.1f'%d['value'], 'cadence16 ms %.3f (stream %.3f)'%(d.get('reference_cadence',{}).get('ms_per_frame',-1), d.get('reference_cadence',{}).get('stream_ms_per_frame_incl_host_launch_gaps',-1)), d['clocks']['sm_mhz'])" || tail -3 gpurun_out/variant.err
  done
done
