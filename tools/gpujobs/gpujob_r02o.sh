# round 2, job O (8 GPUs): final build -- the bench line at N = 8 (+ N = 4), the byte-identity test at world 8
set -x
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
timeout 600 $TR --nproc-per-node 8 --master-port 29508 bench.py --gpus 8 > gpurun_out/r02o_bench_C3_n8.json 2> gpurun_out/r02o_bench_C3_n8.err; echo "bench8 exit=$?"
timeout 600 $TR --nproc-per-node 4 --master-port 29504 bench.py --gpus 4 > gpurun_out/r02o_bench_C3_n4.json 2> gpurun_out/r02o_bench_C3_n4.err; echo "bench4 exit=$?"
timeout 300 $TR --nproc-per-node 8 --master-port 29518 bench.py --gpus 8 --impl reference --steps 3 --warmup 1 > gpurun_out/r02o_bench_ref_n8.json 2> /dev/null; echo "ref8 exit=$?"
python - <<'PY'
import json
def last(p):
    l=[x for x in open(p) if x.startswith('{')]
    return json.loads(l[-1]) if l else None
for n in (8, 4):
    d=last('gpurun_out/r02o_bench_C3_n%d.json' % n)
    if d:
        print(n, {k:d.get(k) for k in ('value','ms_per_frame','sharded_equals_single_gpu')}, d['frame_sha256'][:12], 'e2e', d['e2e']['ms_per_frame'], 'roof', d['roofline']['frac'], 'nccl', d['nccl_gather_comparison']['ms_per_frame'], 'C2', d['extra']['C2']['ms_per_frame'], 'sust', d['sustained']['ms_per_frame'])
r=last('gpurun_out/r02o_bench_ref_n8.json'); print('ref', r and (r['value'], r['config']['parallelism']))
PY
