set -x
mkdir -p gpurun_out
timeout 2400 python -m pytest tests -x -q -m gpu > gpurun_out/r02t_pytest.log 2>&1; echo "pytest exit=$?"; tail -3 gpurun_out/r02t_pytest.log
for a in ieee fma; do
python tools/ab_bench.py --config C3 --arith $a --variants static --frames 6 >> gpurun_out/r02t_ab.log 2>&1
python tools/ab_bench.py --config C2 --arith $a --variants static --frames 6 >> gpurun_out/r02t_ab.log 2>&1
done
python tools/ab_bench.py --config C3 --variants static --shard 0/8 --all-ranks --frames 4 >> gpurun_out/r02t_ab.log 2>&1
cat gpurun_out/r02t_ab.log
