set -x
mkdir -p gpurun_out
L=gpurun_out/r02_phase16_ab2.log
: > $L
for sz in 1280x720 1600x900 1920x1080 2560x1440 3200x1800; do
  echo "== $sz" >> $L
  python tools/ab_bench.py --config C3 --size $sz --phase16 --variants static,lanes2,lanes4,lanes8 --frames 8 >> $L 2>&1
done
echo "== C2 1440p" >> $L
python tools/ab_bench.py --config C2 --size 2560x1440 --phase16 --variants static,lanes2,lanes4 --frames 8 >> $L 2>&1
echo "== row shards C3 1080p" >> $L
python tools/ab_bench.py --config C3 --size 1920x1080 --shard 0/8 --variants static,lanes2,lanes4 --frames 8 >> $L 2>&1
python tools/ab_bench.py --config C3 --size 1920x1080 --shard 0/4 --variants static,lanes2,lanes4 --frames 8 >> $L 2>&1
python tools/ab_bench.py --config C2 --shard 0/4 --variants static,lanes2,lanes4 --frames 8 >> $L 2>&1
python tools/ab_bench.py --config C2 --shard 0/16 --variants static,lanes2,lanes4,lanes8 --frames 8 >> $L 2>&1
cat $L
