# round 2, last GPU call (1 GPU, ~11 GPU-minutes were left): validation of the final tree -- every GPU test, smoke, the default bench line,
# and the launch list of a short bench run.  Most important first: the call's limit is whatever budget remains once the box is up.
set -x
mkdir -p gpurun_out
timeout 400 python -m pytest tests -x -q -m gpu > gpurun_out/r02final_pytest.log 2>&1; echo "pytest exit=$?"; tail -3 gpurun_out/r02final_pytest.log
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r02final_smoke.log 2>&1; echo "smoke exit=$?"; tail -1 gpurun_out/r02final_smoke.log
timeout 400 python bench.py > gpurun_out/r02final_bench_C3_n1.json 2> gpurun_out/r02final_bench.err; echo "bench exit=$?"; tail -c 400 gpurun_out/r02final_bench_C3_n1.json
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02final_launches.raw.csv python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-extras > /dev/null 2>&1; echo "ncu exit=$?"
