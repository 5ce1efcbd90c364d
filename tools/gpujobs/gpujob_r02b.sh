# round 2, job B (1 GPU): scheduler A/B -- static grid vs persistent tile / group / refill; whole frame and 8-way shares
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_scheduler_gpu.py -x -q > gpurun_out/r02b_pytest.log 2>&1; echo "pytest exit=$?"; tail -3 gpurun_out/r02b_pytest.log
python tools/ab_bench.py --config C3 --variants static,tile,group,refill16 > gpurun_out/r02b_ab.log 2>&1
python tools/ab_bench.py --config C2 --variants static,tile,group >> gpurun_out/r02b_ab.log 2>&1
python tools/ab_bench.py --config C3 --variants static,tile,group,lanes2 --shard 0/8 --all-ranks --frames 4 >> gpurun_out/r02b_ab.log 2>&1
python tools/ab_bench.py --config C2 --variants static,tile,group,lanes2 --shard 0/8 --all-ranks --frames 4 >> gpurun_out/r02b_ab.log 2>&1
python tools/ab_bench.py --config C3 --variants static,tile,group --shard 0/2 --all-ranks --frames 4 >> gpurun_out/r02b_ab.log 2>&1
cat gpurun_out/r02b_ab.log
timeout 600 ncu --metrics gpu__time_duration.sum,smsp__inst_executed.sum,smsp__thread_inst_executed_per_inst_executed.ratio,l1tex__t_sector_hit_rate.pct,smsp__issue_active.avg.pct_of_peak_sustained_active --clock-control none -k regex:cloud_march --csv --log-file gpurun_out/r02b_ab_ncu.csv python tools/ab_bench.py --config C3 --variants static,tile,group --frames 1 > /dev/null 2>&1
grep -c cloud_march gpurun_out/r02b_ab_ncu.csv
