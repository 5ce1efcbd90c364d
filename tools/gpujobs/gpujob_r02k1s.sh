# round 2 (1 GPU): ncu --set full of K1s on the launches the reference's cadence makes
set -x
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:cloud_march_split -s 3 -c 1 -f -o gpurun_out/prof_r02_phase16_C2_lanes4 python tools/ab_bench.py --config C2 --phase16 --variants auto --frames 4 > gpurun_out/r02_k1s_ncu1.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:cloud_march_split -s 3 -c 1 -f -o gpurun_out/prof_r02_phase16_C3_lanes2 python tools/ab_bench.py --config C3 --phase16 --variants auto --frames 4 > gpurun_out/r02_k1s_ncu2.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:cloud_march -s 3 -c 1 -f -o gpurun_out/prof_r02_phase16_C3_lanes1 python tools/ab_bench.py --config C3 --phase16 --variants static --frames 4 > gpurun_out/r02_k1s_ncu3.log 2>&1
ls -la gpurun_out/*.ncu-rep
