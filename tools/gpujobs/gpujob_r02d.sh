# round 2, job D (1 GPU): the contracted arithmetic mode -- parity tests, IEEE regression tests, A/B timings, ncu instruction counts
set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_arith_fma_gpu.py -x -q > gpurun_out/r02d_pytest_fma.log 2>&1; echo "fma pytest exit=$?"; tail -5 gpurun_out/r02d_pytest_fma.log
timeout 2400 python -m pytest tests -m gpu -x -q --deselect tests/test_arith_fma_gpu.py > gpurun_out/r02d_pytest.log 2>&1; echo "pytest exit=$?"; tail -5 gpurun_out/r02d_pytest.log
for a in ieee fma; do
  python tools/ab_bench.py --config C3 --arith $a --variants static >> gpurun_out/r02d_ab.log 2>&1
  python tools/ab_bench.py --config C2 --arith $a --variants static >> gpurun_out/r02d_ab.log 2>&1
  python tools/ab_bench.py --config C3 --arith $a --filter hybrid --variants static >> gpurun_out/r02d_ab.log 2>&1
  python tools/ab_bench.py --config C3 --arith $a --filter exact --variants static >> gpurun_out/r02d_ab.log 2>&1
  python tools/ab_bench.py --config C3 --arith $a --variants static,lanes2 --shard 0/8 --all-ranks --frames 4 >> gpurun_out/r02d_ab.log 2>&1
done
cat gpurun_out/r02d_ab.log
timeout 600 ncu --metrics gpu__time_duration.sum,smsp__inst_executed.sum,smsp__thread_inst_executed_per_inst_executed.ratio,smsp__issue_active.avg.pct_of_peak_sustained_active --clock-control none -k regex:cloud_march --csv --log-file gpurun_out/r02d_ab_ncu.csv python tools/ab_bench.py --config C3 --arith fma --variants static --frames 1 > /dev/null 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:cloud_march -s 2 -c 1 -f -o gpurun_out/prof_r02d_C3_hw_fma python tools/ab_bench.py --config C3 --arith fma --variants static --frames 1 > /dev/null 2> gpurun_out/r02d_ncu.err
timeout 900 ncu --set full --clock-control none --import-source on -k regex:cloud_march -s 2 -c 1 -f -o gpurun_out/prof_r02d_C3_hw python tools/ab_bench.py --config C3 --arith ieee --variants static --frames 1 > /dev/null 2>> gpurun_out/r02d_ncu.err
