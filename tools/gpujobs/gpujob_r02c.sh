set -x
mkdir -p gpurun_out
python tools/ab_bench.py --config C3 --variants static,tile > gpurun_out/r02c_ab.log 2>&1
python tools/ab_bench.py --config C2 --variants static,tile >> gpurun_out/r02c_ab.log 2>&1
python tools/ab_bench.py --config C3 --variants static,tile,lanes2 --shard 0/8 --all-ranks --frames 4 >> gpurun_out/r02c_ab.log 2>&1
python tools/ab_bench.py --config C2 --variants static,tile,lanes2 --shard 0/8 --all-ranks --frames 4 >> gpurun_out/r02c_ab.log 2>&1
python tools/ab_bench.py --config C3 --variants static,tile --shard 0/2 --all-ranks --frames 4 >> gpurun_out/r02c_ab.log 2>&1
python tools/ab_bench.py --config C3 --variants static,tile --shard 0/4 --all-ranks --frames 4 >> gpurun_out/r02c_ab.log 2>&1
cat gpurun_out/r02c_ab.log
