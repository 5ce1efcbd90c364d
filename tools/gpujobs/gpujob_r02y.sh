set -x
mkdir -p gpurun_out
python tools/ab_bench.py --config C3 --variants static --shard 0/8 --all-ranks --frames 5 > gpurun_out/r02y_ab.log 2>&1
python tools/ab_bench.py --config C3 --variants static --shard 0/8 --all-ranks --frames 5 --snake 2>&1 | sed 's/^/snake /' >> gpurun_out/r02y_ab.log
python tools/ab_bench.py --config C3 --variants static --shard 0/8 --all-ranks --frames 5 --row-block 16 2>&1 | sed 's/^/rb16 /' >> gpurun_out/r02y_ab.log
python tools/ab_bench.py --config C3 --variants static --shard 0/8 --all-ranks --frames 5 --row-block 16 --snake 2>&1 | sed 's/^/rb16 snake /' >> gpurun_out/r02y_ab.log
python tools/ab_bench.py --config C3 --variants static --shard 0/8 --all-ranks --frames 5 --row-block 24 2>&1 | sed 's/^/rb24 /' >> gpurun_out/r02y_ab.log
cat gpurun_out/r02y_ab.log
