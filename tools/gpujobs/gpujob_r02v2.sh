# round 2, final validation after the K1s fast path (1 GPU): smoke, every GPU test, sanitizer, the default bench line
set -x
mkdir -p gpurun_out
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r02v2_smoke.log 2>&1; echo "smoke exit=$?"; tail -1 gpurun_out/r02v2_smoke.log
timeout 2400 python -m pytest tests -x -q -m gpu > gpurun_out/r02v2_pytest.log 2>&1; echo "pytest exit=$?"; tail -3 gpurun_out/r02v2_pytest.log
timeout 1500 bash tools/sanitize.sh 2>&1 | tail -10
for f in gpurun_out/sanitizer_*.log; do cp $f gpurun_out/r02v2_$(basename $f); done
timeout 900 python bench.py > gpurun_out/r02v2_bench_C3_n1.json 2> gpurun_out/r02v2_bench.err; tail -c 600 gpurun_out/r02v2_bench_C3_n1.json
