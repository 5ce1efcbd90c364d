# round 2, the last two GPU-minutes: compute-sanitizer memcheck over the pow-filter build (K1 and K1s, three sampler modes, both arithmetic builds)
mkdir -p gpurun_out
SAN_QUICK=1 timeout 105 bash tools/sanitize.sh 2>&1 | tail -6
cp gpurun_out/sanitizer_memcheck.log gpurun_out/r02pow_sanitizer_memcheck.log
