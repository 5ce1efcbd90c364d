set -x
mkdir -p gpurun_out
timeout 3000 bash tools/sanitize.sh 2>&1 | tail -14
