# round 2 (1 GPU): ncu --set full with source of the FINAL build -- K1 on the whole C3 frame, and K5 / K6 at 1080p -- for per-line budgets
set -x
mkdir -p gpurun_out
timeout 240 ncu --set full --clock-control none --import-source on -k regex:cloud_march_kernel -s 3 -c 1 -f -o gpurun_out/r02final_k1 python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-extras > /dev/null 2> gpurun_out/r02final_ncu_k1.err; echo "ncu k1 exit=$?"
timeout 200 ncu --set full --clock-control none --import-source on -k regex:'god_ray|radial_blur|reproject' -c 10 -f -o gpurun_out/r02final_aux python tools/aux_kernels_driver.py > gpurun_out/r02final_aux.log 2>&1; echo "ncu aux exit=$?"
ls -la gpurun_out
