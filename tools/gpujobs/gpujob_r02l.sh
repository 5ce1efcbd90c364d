set -x
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_packed_gpu.py -x -q > gpurun_out/r02l_pytest.log 2>&1; echo "pytest exit=$?"; tail -15 gpurun_out/r02l_pytest.log
for a in ieee fma; do
python tools/ab_bench.py --config C3 --arith $a --variants static,packed >> gpurun_out/r02l_ab.log 2>&1
python tools/ab_bench.py --config C2 --arith $a --variants static,packed >> gpurun_out/r02l_ab.log 2>&1
done
python tools/ab_bench.py --config C3 --variants static,packed --shard 0/8 --all-ranks --frames 4 >> gpurun_out/r02l_ab.log 2>&1
python tools/ab_bench.py --config C2 --variants static,packed,lanes2 --shard 0/8 --all-ranks --frames 4 >> gpurun_out/r02l_ab.log 2>&1
cat gpurun_out/r02l_ab.log
timeout 600 ncu --metrics gpu__time_duration.sum,smsp__inst_executed.sum,smsp__thread_inst_executed_per_inst_executed.ratio,smsp__issue_active.avg.pct_of_peak_sustained_active,sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active,sm__warps_active.avg.pct_of_peak_sustained_active --clock-control none -k regex:cloud_march --csv --log-file gpurun_out/r02l_ab_ncu.csv python tools/ab_bench.py --config C3 --variants static,packed --frames 1 > /dev/null 2>&1
