set -x
mkdir -p gpurun_out
timeout 2400 python -m pytest tests -m gpu -x -q > gpurun_out/r02h_pytest.log 2>&1; echo "pytest exit=$?"; tail -4 gpurun_out/r02h_pytest.log
for a in ieee fma; do
python tools/ab_bench.py --config C3 --arith $a --variants static >> gpurun_out/r02h_ab.log 2>&1
python tools/ab_bench.py --config C2 --arith $a --variants static >> gpurun_out/r02h_ab.log 2>&1
done
python tools/ab_bench.py --config C3 --variants static --shard 0/8 --all-ranks --frames 4 >> gpurun_out/r02h_ab.log 2>&1
cat gpurun_out/r02h_ab.log
