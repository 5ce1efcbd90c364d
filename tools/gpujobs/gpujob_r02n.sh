# round 2, job N (2 GPUs): reprojection fast path, single-process two-device dispatch, N = 2 bench with the host barrier
set -x
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_reproject.py tests/test_multigpu_gpu.py tests/test_post_chain.py -x -q -m gpu -k "not sharded_frame or 2" > gpurun_out/r02n_pytest.log 2>&1; echo "pytest exit=$?"; tail -3 gpurun_out/r02n_pytest.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29502 bench.py --gpus 2 > gpurun_out/r02n_bench_C3_n2.json 2> gpurun_out/r02n_bench_C3_n2.err; echo "bench2 exit=$?"; grep -v "^\[W\|^$\|OMP_NUM\|\*\*\*" gpurun_out/r02n_bench_C3_n2.err | tail -5
timeout 600 python bench.py --steps 10 --no-cpu-baseline --sustained-seconds 0 > gpurun_out/r02n_bench_C3_n1_quick.json 2> gpurun_out/r02n_bench_C3_n1_quick.err; echo "bench1 exit=$?"
python - <<'PY'
import json
def last(p):
    l=[x for x in open(p) if x.startswith('{')]
    return json.loads(l[-1]) if l else None
d=last('gpurun_out/r02n_bench_C3_n2.json')
if d: print('N=2', d['ms_per_frame'], 'e2e', d['e2e']['ms_per_frame'], d['sharded_equals_single_gpu'], d['frame_sha256'][:12], 'nccl', d['nccl_gather_comparison']['ms_per_frame'])
d=last('gpurun_out/r02n_bench_C3_n1_quick.json')
if d: print('N=1', d['ms_per_frame'], 'e2e', d['e2e']['ms_per_frame'], 'cadence', d['reference_cadence'], 'sched', d['scheduler_ms_per_frame'])
PY
