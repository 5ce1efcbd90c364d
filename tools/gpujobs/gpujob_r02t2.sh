# round 2, job T2 (1 GPU): every GPU test on the build with the new automatic lane thresholds, the automatic choice timed, a C2 bench line
set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -x -q -m gpu > gpurun_out/r02t_pytest.log 2>&1; tail -2 gpurun_out/r02t_pytest.log
python tools/ab_bench.py --config C2 --phase16 --variants auto --frames 8
python tools/ab_bench.py --config C3 --size 1280x720 --phase16 --variants auto --frames 8
python tools/ab_bench.py --config C2 --shard 0/8 --all-ranks --variants auto --frames 5
timeout 600 python bench.py --config C2 --no-cpu-baseline > gpurun_out/r02t_bench_C2_n1.json 2> gpurun_out/r02t_bench.err; tail -c 900 gpurun_out/r02t_bench_C2_n1.json
