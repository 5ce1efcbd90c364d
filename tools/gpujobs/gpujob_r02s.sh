# round 2, job S (1 GPU): what the driver runs at round end -- smoke, the GPU tests, the bench arms
set -x
mkdir -p gpurun_out
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r02s_smoke.log 2>&1; echo "smoke exit=$?"; tail -4 gpurun_out/r02s_smoke.log | cut -c1-400
timeout 2400 python -m pytest tests -x -q -m gpu > gpurun_out/r02s_pytest.log 2>&1; echo "pytest exit=$?"; tail -3 gpurun_out/r02s_pytest.log
timeout 600 python bench.py --impl reference --gpus 1 --steps 20 --warmup 3 > gpurun_out/r02s_bench_reference.json 2> gpurun_out/r02s_bench_reference.err; echo "ref exit=$?"
timeout 900 python bench.py --gpus 1 --steps 20 --warmup 3 > gpurun_out/r02s_bench_C3.json 2> gpurun_out/r02s_bench_C3.err; echo "bench exit=$?"; tail -2 gpurun_out/r02s_bench_C3.err
