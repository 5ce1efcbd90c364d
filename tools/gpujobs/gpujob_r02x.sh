# round 2, job X (1 GPU): final validation -- smoke, every GPU test, sanitizer, ncu captures of the final build
set -x
mkdir -p gpurun_out
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r02x_smoke.log 2>&1; echo "smoke exit=$?"; tail -1 gpurun_out/r02x_smoke.log
timeout 2400 python -m pytest tests -x -q -m gpu > gpurun_out/r02x_pytest.log 2>&1; echo "pytest exit=$?"; tail -3 gpurun_out/r02x_pytest.log
timeout 1500 bash tools/sanitize.sh 2>&1 | tail -10
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02x_launches_C3.csv python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-extras > gpurun_out/r02x_launch_bench.json 2> gpurun_out/r02x_launch.err
timeout 900 ncu --set full --clock-control none --import-source on -k regex:cloud_march -s 3 -c 1 -f -o gpurun_out/prof_r02x_C3_hw python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-extras > /dev/null 2> gpurun_out/r02x_ncu_hw.err
timeout 900 ncu --set full --clock-control none --import-source on -k regex:cloud_march -s 3 -c 1 -f -o gpurun_out/prof_r02x_C3_hw_fma python bench.py --arith fma --steps 2 --warmup 3 --no-cpu-baseline --no-extras > /dev/null 2> gpurun_out/r02x_ncu_fma.err
