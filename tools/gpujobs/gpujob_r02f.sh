set -x
mkdir -p gpurun_out
M=gpu__time_duration.sum,smsp__inst_executed.sum,smsp__issue_active.avg.pct_of_peak_sustained_active,sm__cycles_active.avg,sm__cycles_active.min,sm__cycles_active.max,sm__cycles_elapsed.max,sm__warps_active.avg.pct_of_peak_sustained_active,smsp__thread_inst_executed_per_inst_executed.ratio
for cfg in C3 C2; do
timeout 300 ncu --metrics $M --clock-control none -k regex:cloud_march --csv --log-file gpurun_out/r02f_shard_${cfg}.csv python tools/ab_bench.py --config $cfg --variants static,lanes2 --shard 2/8 --frames 1 > /dev/null 2>&1
done
python tools/ab_bench.py --config C3 --variants static --shard 2/8 --frames 6 > gpurun_out/r02f_order.log 2>&1
MM_DEBUG_ROW_ORDER=identity python tools/ab_bench.py --config C3 --variants static --shard 2/8 --frames 6 >> gpurun_out/r02f_order.log 2>&1
MM_DEBUG_ROW_ORDER=reverse python tools/ab_bench.py --config C3 --variants static --shard 2/8 --frames 6 >> gpurun_out/r02f_order.log 2>&1
python tools/ab_bench.py --config C2 --variants static,lanes2 --shard 2/8 --frames 6 >> gpurun_out/r02f_order.log 2>&1
MM_DEBUG_ROW_ORDER=identity python tools/ab_bench.py --config C2 --variants static,lanes2 --shard 2/8 --frames 6 >> gpurun_out/r02f_order.log 2>&1
cat gpurun_out/r02f_order.log
