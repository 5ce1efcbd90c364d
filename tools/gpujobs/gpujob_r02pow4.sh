# round 2, what is left of the GPU budget: the full sanitizer pass (memcheck + racecheck, every variant) over the pow-filter build, as far as it gets
mkdir -p gpurun_out
timeout 80 bash tools/sanitize.sh 2>&1 | tail -8
for t in memcheck racecheck; do [ -f gpurun_out/sanitizer_$t.log ] && cp gpurun_out/sanitizer_$t.log gpurun_out/r02pow_full_sanitizer_$t.log; done
