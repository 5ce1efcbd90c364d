set -x
mkdir -p gpurun_out
python tools/e2e_probe.py --config C3 > gpurun_out/r02u_e2e.log 2>&1; python tools/e2e_probe.py --config C2 >> gpurun_out/r02u_e2e.log 2>&1; cat gpurun_out/r02u_e2e.log
timeout 1200 python -m pytest tests/test_march_parity_gpu.py tests/test_packed_gpu.py tests/test_arith_fma_gpu.py -x -q > gpurun_out/r02u_pytest.log 2>&1; echo "pytest exit=$?"; tail -2 gpurun_out/r02u_pytest.log
python tools/ab_bench.py --config C3 --variants static --frames 6
