set -x
mkdir -p gpurun_out
L=gpurun_out/r02_phase16_ab.log
python tools/ab_bench.py --config C3 --phase16 --variants static,lanes2,lanes4,tile,refill16,auto --frames 8 > $L 2>&1
python tools/ab_bench.py --config C2 --phase16 --variants static,lanes2,lanes4,lanes8,tile,auto --frames 8 >> $L 2>&1
MM_LIBRARY=$PWD/variants/nofast.so python tools/ab_bench.py --config C3 --phase16 --variants lanes2,lanes4 --frames 8 2>&1 | sed 's/^/nofast /' >> $L
cat $L
