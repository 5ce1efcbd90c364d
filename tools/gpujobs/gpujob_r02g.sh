# round 2, job G (8 GPUs): the headline frame sharded over 8 B200s (+ C2, NCCL comparison, sustained), C5 / C5b 64-frame animation, byte-identity test
set -x
mkdir -p gpurun_out
nvidia-smi -L | wc -l
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1"
timeout 600 $TR --master-port 29508 bench.py --gpus 8 > gpurun_out/r02g_bench_C3_n8.json 2> gpurun_out/r02g_bench_C3_n8.err; echo "bench8 exit=$?"
timeout 300 $TR --master-port 29509 bench.py --gpus 8 --config C5 --animation 64 --steps 20 --warmup 3 > gpurun_out/r02g_anim_C5_n8.json 2> gpurun_out/r02g_anim_C5_n8.err; echo "anim C5 exit=$?"
timeout 300 $TR --master-port 29510 bench.py --gpus 8 --config C5b --animation 64 --steps 20 --warmup 3 > gpurun_out/r02g_anim_C5b_n8.json 2> gpurun_out/r02g_anim_C5b_n8.err; echo "anim C5b exit=$?"
timeout 300 $TR --master-port 29511 bench.py --gpus 8 --arith fma --no-extras --no-cpu-baseline --steps 10 > gpurun_out/r02g_bench_C3_n8_fma.json 2> gpurun_out/r02g_bench_C3_n8_fma.err; echo "fma8 exit=$?"
timeout 400 python -m pytest tests/test_multigpu_gpu.py -x -q -k "8" > gpurun_out/r02g_pytest_mgpu.log 2>&1; echo "pytest exit=$?"; tail -2 gpurun_out/r02g_pytest_mgpu.log
python - <<'PY'
import json
def last(p):
    l=[x for x in open(p) if x.startswith('{')]
    return json.loads(l[-1]) if l else None
d=last('gpurun_out/r02g_bench_C3_n8.json')
if d:
    print({k:d.get(k) for k in ('value','ms_per_frame','frame_sha256','sharded_equals_single_gpu','per_rank_kernel_ms')})
    print('e2e', d['e2e']['ms_per_frame'], d['e2e']['value']); print('roofline', d.get('roofline',{}).get('frac')); print('nccl', d.get('nccl_gather_comparison')); print('extra', d.get('extra')); print('sustained', d.get('sustained'))
for f in ('r02g_anim_C5_n8','r02g_anim_C5b_n8','r02g_bench_C3_n8_fma'):
    a=last('gpurun_out/%s.json'%f)
    print(f, a and {k:a.get(k) for k in ('value','ms_per_step','ms_per_frame','n_gpus','frame_sha256','sharded_equals_single_gpu')})
PY
