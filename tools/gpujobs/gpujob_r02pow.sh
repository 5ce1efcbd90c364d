# round 2 (1 GPU): the exact pow filter of cloudTest (MM_POW_FILTER) -- every GPU test on the new build, A/B against the same source built
# with -DMM_POW_FILTER=0 (whole C3 frame, the FP32-sampler march, one 8-way share of C2), ncu --set full of the new K1 (its summary feeds
# bench.py's roofline_issue, so it is regenerated on the box before the bench line is taken), the default bench line
set -x
mkdir -p gpurun_out
timeout 400 python -m pytest tests -x -q -m gpu > gpurun_out/r02pow_pytest.log 2>&1; echo "pytest exit=$?"; tail -3 gpurun_out/r02pow_pytest.log
for v in old new; do
  lib=$PWD/project-marshmallow_b200/libmarshmallow_b200.so; [ $v = old ] && lib=$PWD/project-marshmallow_b200/variants/nopowfilter.so
  for args in "--config C3" "--config C3 --filter hybrid" "--config C2 --shard 0/8 --variants auto"; do
    echo "$v: $(MM_LIBRARY=$lib timeout 120 python tools/ab_bench.py --frames 6 --variants static $args 2>&1 | tail -1)" | tee -a gpurun_out/r02pow_ab.txt
  done
done
timeout 240 ncu --set full --clock-control none --import-source on -k regex:cloud_march_kernel -s 3 -c 1 -f -o gpurun_out/r02pow_k1 python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-extras > /dev/null 2> gpurun_out/r02pow_ncu.err; echo "ncu exit=$?"
python tools/ncu_summary.py gpurun_out/r02pow_k1.ncu-rep > gpurun_out/r02pow_C3_hw.summary.csv && [ -s gpurun_out/r02pow_C3_hw.summary.csv ] && cp gpurun_out/r02pow_C3_hw.summary.csv profiles/r02_C3_hw.summary.csv
timeout 400 python bench.py > gpurun_out/r02pow_bench_C3_n1.json 2> gpurun_out/r02pow_bench.err; echo "bench exit=$?"; tail -c 300 gpurun_out/r02pow_bench_C3_n1.json
