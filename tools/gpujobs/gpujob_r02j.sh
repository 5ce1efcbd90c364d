# round 2, job J (1 GPU): full GPU tests, default bench, reference arm, launch list + full ncu capture of the final build, sanitizer (both arithmetic builds)
set -x
mkdir -p gpurun_out
timeout 2400 python -m pytest tests -m gpu -x -q > gpurun_out/r02j_pytest.log 2>&1; echo "pytest exit=$?"; tail -3 gpurun_out/r02j_pytest.log
timeout 900 python bench.py > gpurun_out/r02j_bench_C3.json 2> gpurun_out/r02j_bench_C3.err; echo "bench exit=$?"; tail -3 gpurun_out/r02j_bench_C3.err
timeout 600 python bench.py --impl reference --steps 5 --warmup 2 > gpurun_out/r02j_bench_reference.json 2> gpurun_out/r02j_bench_reference.err; echo "ref exit=$?"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02j_launches_C3.csv python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-extras > gpurun_out/r02j_launch_bench.json 2> gpurun_out/r02j_launch.err
timeout 900 ncu --set full --clock-control none --import-source on -k regex:cloud_march -s 3 -c 1 -f -o gpurun_out/prof_r02j_C3_hw python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-extras > /dev/null 2> gpurun_out/r02j_ncu_hw.err
timeout 900 ncu --set full --clock-control none --import-source on -k regex:cloud_march -s 3 -c 1 -f -o gpurun_out/prof_r02j_C3_hw_fma python bench.py --arith fma --steps 2 --warmup 3 --no-cpu-baseline --no-extras > /dev/null 2> gpurun_out/r02j_ncu_fma.err
timeout 1500 bash tools/sanitize.sh 2>&1 | tail -12
