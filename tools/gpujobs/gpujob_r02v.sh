set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_march_parity_gpu.py tests/test_multigpu_gpu.py tests/test_scheduler_gpu.py -x -q -k "render_to_host or mirror or sharded_frame_is_byte_identical or dispatch_multi or persistent_scheduler_changes or lanes_per_ray" > gpurun_out/r02v_pytest.log 2>&1; echo "pytest exit=$?"; tail -4 gpurun_out/r02v_pytest.log
python tools/e2e_probe.py --config C3 > gpurun_out/r02v_e2e.log 2>&1; MM_E2E_SPLIT=0 python tools/e2e_probe.py --config C3 2>&1 | sed 's/^/nosplit /' >> gpurun_out/r02v_e2e.log; python tools/e2e_probe.py --config C2 >> gpurun_out/r02v_e2e.log 2>&1; cat gpurun_out/r02v_e2e.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29502 bench.py --gpus 2 --no-extras --no-cpu-baseline --steps 10 > gpurun_out/r02v_bench_n2.json 2> gpurun_out/r02v_bench_n2.err; echo "bench2 exit=$?"; grep -v "^\[W\|^$\|OMP_NUM\|\*\*\*" gpurun_out/r02v_bench_n2.err | tail -5
python -c "
import json
d=json.loads([l for l in open('gpurun_out/r02v_bench_n2.json') if l.startswith('{')][-1]); print('N=2', d['ms_per_frame'], 'e2e', d['e2e']['ms_per_frame'], d['sharded_equals_single_gpu'])"
