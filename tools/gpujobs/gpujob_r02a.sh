# round 2, job A (1 GPU): GPU tests, default bench (C3 4K), launch list and one full ncu capture of the march kernel
set -x
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv
nproc
timeout 2400 python -m pytest tests -m gpu -x -q > gpurun_out/r02a_pytest.log 2>&1; echo "pytest exit=$?"; tail -5 gpurun_out/r02a_pytest.log
timeout 900 python bench.py > gpurun_out/r02a_bench_C3.json 2> gpurun_out/r02a_bench_C3.err; echo "bench exit=$?"; tail -3 gpurun_out/r02a_bench_C3.err
timeout 600 python bench.py --impl reference --steps 5 --warmup 2 > gpurun_out/r02a_bench_reference.json 2> gpurun_out/r02a_bench_reference.err; echo "ref exit=$?"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02a_launches_C3.csv python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-extras > gpurun_out/r02a_launch_bench.json 2> gpurun_out/r02a_launch.err
timeout 900 ncu --set full --clock-control none --import-source on -k regex:cloud_march -s 3 -c 1 -f -o gpurun_out/prof_r02a_C3_hw python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-extras > /dev/null 2> gpurun_out/r02a_ncu_hw.err
ls -la gpurun_out | tail -8
