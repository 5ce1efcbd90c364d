# round 2 (8 GPUs): the default bench line of the final build at N = 8, and the copy2 delivery path beside it
set -x
mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus 8 > gpurun_out/r02z_bench_C3_n8.json 2> gpurun_out/r02z_bench_n8.err
tail -c 1500 gpurun_out/r02z_bench_C3_n8.json
MM_E2E_REST=copy2 timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29542 bench.py --gpus 8 --steps 10 --warmup 3 --no-extras --no-cpu-baseline > gpurun_out/r02z_bench_C3_n8_copy2.json 2> gpurun_out/r02z_bench_n8_copy2.err
MM_E2E_REST=fused timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29543 bench.py --gpus 8 --steps 10 --warmup 3 --no-extras --no-cpu-baseline > gpurun_out/r02z_bench_C3_n8_fused.json 2> gpurun_out/r02z_bench_n8_fused.err
python - <<'PY'
import json
for m in ("copy2", "fused"):
    d = json.loads(open(f"gpurun_out/r02z_bench_C3_n8_{m}.json").read().strip().splitlines()[-1])
    print(f"N=8 rest={m:6s} kernel {d['ms_per_step']:.3f} ms  e2e {d['e2e']['ms_per_frame']:.3f} ms  host hash {d['e2e']['host_frame_sha256'][:8]}")
PY
