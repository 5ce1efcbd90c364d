# round 2, job E (2 GPUs): new GPU tests (interop, night sky, 8K rows, div selftest), multi-GPU test, bench at N = 2 with the NCCL comparison
set -x
mkdir -p gpurun_out
nvidia-smi -L
timeout 1500 python -m pytest tests/test_interop_gpu.py tests/test_multigpu_gpu.py tests/test_march_parity_gpu.py -x -q -s -k "interop or external or sharded or night_frame or native_8k or exact_divide" > gpurun_out/r02e_pytest.log 2>&1; echo "pytest exit=$?"; grep -E "DRIVER ANSWER|pipe fd|passed|failed|Error" gpurun_out/r02e_pytest.log | head -20; tail -3 gpurun_out/r02e_pytest.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29502 bench.py --gpus 2 > gpurun_out/r02e_bench_C3_n2.json 2> gpurun_out/r02e_bench_C3_n2.err; echo "bench2 exit=$?"; tail -5 gpurun_out/r02e_bench_C3_n2.err
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29503 bench.py --gpus 2 --impl reference --steps 3 --warmup 1 > gpurun_out/r02e_bench_ref_n2.json 2> gpurun_out/r02e_bench_ref_n2.err; echo "ref2 exit=$?"
python - <<'PY'
import json
d=json.loads([l for l in open('gpurun_out/r02e_bench_C3_n2.json') if l.startswith('{')][-1])
print({k:d[k] for k in ('value','ms_per_frame','frame_sha256','sharded_equals_single_gpu','per_rank_kernel_ms')})
print('e2e', d['e2e']); print('roofline', d.get('roofline',{}).get('frac')); print('nccl', d.get('nccl_gather_comparison')); print('extra', d.get('extra')); print('sustained', d.get('sustained'))
r=json.loads([l for l in open('gpurun_out/r02e_bench_ref_n2.json') if l.startswith('{')][-1]); print('ref', r['value'], r['config']['parallelism'])
PY
