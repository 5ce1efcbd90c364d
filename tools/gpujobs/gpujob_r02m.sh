set -x
mkdir -p gpurun_out
for nb in 4 5 6; do MM_X2_BLOCKS=$nb python tools/ab_bench.py --config C3 --variants packed --frames 4 2>&1 | sed "s/^/blocks=$nb /" >> gpurun_out/r02m_ab.log; done
python tools/ab_bench.py --config C3 --variants static --frames 4 >> gpurun_out/r02m_ab.log 2>&1
cat gpurun_out/r02m_ab.log
timeout 600 ncu --section SourceCounters --section LaunchStats --section Occupancy --section WarpStateStats --section SchedulerStats --clock-control none --import-source on -k regex:x2 -s 2 -c 1 -f -o gpurun_out/prof_r02m_x2 python tools/ab_bench.py --config C3 --variants packed --frames 1 > /dev/null 2> gpurun_out/r02m.err
ls -la gpurun_out/prof_r02m_x2.ncu-rep
