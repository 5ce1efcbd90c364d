# round 2 (4 GPUs): the three host-delivery paths of a mirrored sharded dispatch, end to end
set -x
mkdir -p gpurun_out
L=gpurun_out/r02_e2e_rest_modes.txt
: > $L
timeout 600 python -m pytest tests/test_march_parity_gpu.py -m gpu -x -q -k "host_mirror or render_to_host" > gpurun_out/r02_e2e_tests.log 2>&1; tail -2 gpurun_out/r02_e2e_tests.log
run() {  # N mode
  MM_E2E_REST=$2 timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $1 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus $1 --steps 10 --warmup 3 --no-extras --no-cpu-baseline > gpurun_out/e2e_tmp.json 2> gpurun_out/e2e_tmp.err
  python - "$1" "$2" <<'PY' >> gpurun_out/r02_e2e_rest_modes.txt
import json, sys
try:
    d = json.loads(open('gpurun_out/e2e_tmp.json').read().strip().splitlines()[-1])
    print(f"N={sys.argv[1]} rest={sys.argv[2]:6s} kernel {d['ms_per_step']:.3f} ms  e2e {d['e2e']['ms_per_frame']:.3f} ms  host hash {d['e2e']['host_frame_sha256'][:8]} frame hash {d['frame_sha256'][:8]}")
except Exception as e:
    print(f"N={sys.argv[1]} rest={sys.argv[2]} FAILED {e}"); print(open('gpurun_out/e2e_tmp.err').read()[-1500:])
PY
}
run 4 fused
run 4 copy
run 4 copy2
run 2 copy
run 2 copy2
run 2 fused
cat $L
