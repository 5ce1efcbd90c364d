# round 2, last call (1 GPU, 3 GPU-minutes left): the pow-filter build under the contracted arithmetic (ncu summary that bench.py --arith fma reads)
# and every rank's share of the 8-way C3 / C2 frames timed alone (a rank's kernel does not depend on the others)
set -x
mkdir -p gpurun_out
timeout 100 ncu --set full --clock-control none -k regex:cloud_march_kernel -s 3 -c 1 -f -o gpurun_out/r02pow_k1_fma python bench.py --arith fma --steps 2 --warmup 3 --no-cpu-baseline --no-extras > /dev/null 2> gpurun_out/r02pow_ncu_fma.err; echo "ncu exit=$?"
python tools/ncu_summary.py gpurun_out/r02pow_k1_fma.ncu-rep > gpurun_out/r02pow_C3_hw_fma.summary.csv
timeout 60 python tools/ab_bench.py --frames 4 --config C3 --shard 0/8 --all-ranks --variants auto 2>&1 | tail -1 | tee gpurun_out/r02pow_shares.txt
timeout 60 python tools/ab_bench.py --frames 4 --config C2 --shard 0/8 --all-ranks --variants auto 2>&1 | tail -1 | tee -a gpurun_out/r02pow_shares.txt
