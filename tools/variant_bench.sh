#!/bin/bash
# usage (on the GPU box): bash tools/variant_bench.sh "C2 C3" [bench args] -> one line per (library variant, trips-in-flight, config)
# library variants: project-marshmallow_b200/variants/*.so (same source, different -D build options), else the in-tree build
cfgs=${1:-"C2"}; shift
mkdir -p gpurun_out
libs=$(ls project-marshmallow_b200/variants/*.so 2>/dev/null); [ -z "$libs" ] && libs=project-marshmallow_b200/libmarshmallow_b200.so
for lib in $libs; do
  for trips in 1 2 4 8; do
    for cfg in $cfgs; do
      MM_LIBRARY=$PWD/$lib python bench.py --steps 10 --warmup 3 --config $cfg --no-cpu-baseline --lanes $trips "$@" > gpurun_out/variant.json 2> gpurun_out/variant.err
      python -c "
import json;d=json.load(open('gpurun_out/variant.json'));c=d.get('reference_cadence',{});print('$(basename $lib)', 'lanes=$trips', '$cfg', 'ms/frame %.3f'%d['ms_per_frame'], 'Mpix/s %.1f'%d['value'], 'cadence16 ms %.3f (reproject %.3f + march %.3f)'%(c.get('ms_per_frame',-1), c.get('reproject_kernel_ms',-1), c.get('phase16_march_kernel_ms',-1)), d['clocks']['sm_mhz'])" || tail -3 gpurun_out/variant.err
    done
  done
done
