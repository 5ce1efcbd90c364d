set -x
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -6
bash tools/sanitize.sh 2>&1 | tail -12
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_bench_default.csv python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/launch_bench.json 2> gpurun_out/launch.err
ncu --set full --clock-control none --import-source on -k regex:cloud_march -s 3 -c 1 -f -o gpurun_out/prof_r01b_hw python bench.py --steps 2 --warmup 3 --no-cpu-baseline > /dev/null 2> gpurun_out/ncu_hw.err
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:'god_ray|radial_blur|present|shadow|reproject|tonemap' --csv --log-file gpurun_out/aux_launches.csv python tools/aux_kernels_driver.py > gpurun_out/aux.log 2>&1
tail -2 gpurun_out/aux.log
ls -la gpurun_out | tail -8
