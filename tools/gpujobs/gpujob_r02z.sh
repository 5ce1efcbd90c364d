set -x
mkdir -p gpurun_out
L=gpurun_out/r02z_k1s.log
timeout 600 python -m pytest tests/test_march_parity_gpu.py tests/test_arith_fma_gpu.py -m gpu -x -q -k "lanes or split or phase or fma" > gpurun_out/r02z_tests.log 2>&1; tail -3 gpurun_out/r02z_tests.log
: > $L
for lib in variants/nofast.so project-marshmallow_b200/libmarshmallow_b200.so; do
  echo "== $lib" >> $L
  MM_LIBRARY=$PWD/$lib python tools/ab_bench.py --config C2 --variants static,lanes2,lanes4 --frames 8 >> $L 2>&1
  MM_LIBRARY=$PWD/$lib python tools/ab_bench.py --config C2 --variants lanes2,lanes4,lanes8,auto --shard 0/8 --all-ranks --frames 5 >> $L 2>&1
  MM_LIBRARY=$PWD/$lib python tools/ab_bench.py --config C3 --variants static,lanes2 --shard 0/8 --all-ranks --frames 5 >> $L 2>&1
  MM_LIBRARY=$PWD/$lib python tools/ab_bench.py --config C3 --variants static,lanes2,lanes4 --shard 0/16 --frames 5 >> $L 2>&1
done
cat $L
