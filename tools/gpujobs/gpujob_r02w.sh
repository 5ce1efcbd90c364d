# round 2, job W (8 GPUs): final build at N = 8 and N = 4
set -x
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
timeout 600 $TR --nproc-per-node 8 --master-port 29508 bench.py --gpus 8 > gpurun_out/r02w_bench_C3_n8.json 2> gpurun_out/r02w_bench_C3_n8.err; echo "bench8 exit=$?"
timeout 600 $TR --nproc-per-node 4 --master-port 29504 bench.py --gpus 4 > gpurun_out/r02w_bench_C3_n4.json 2> gpurun_out/r02w_bench_C3_n4.err; echo "bench4 exit=$?"
timeout 600 $TR --nproc-per-node 2 --master-port 29502 bench.py --gpus 2 > gpurun_out/r02w_bench_C3_n2.json 2> gpurun_out/r02w_bench_C3_n2.err; echo "bench2 exit=$?"
timeout 600 python bench.py > gpurun_out/r02w_bench_C3_n1.json 2> gpurun_out/r02w_bench_C3_n1.err; echo "bench1 exit=$?"
python - <<'PY'
import json
def last(p):
    l=[x for x in open(p) if x.startswith('{')]
    return json.loads(l[-1]) if l else None
for n in (1, 2, 4, 8):
    d=last('gpurun_out/r02w_bench_C3_n%d.json' % n)
    if d:
        print(n, 'value', round(d['value'],1), 'ms', round(d['ms_per_frame'],4), d.get('sharded_equals_single_gpu'), d['frame_sha256'][:12], 'e2e', round(d['e2e']['ms_per_frame'],4), round(d['e2e']['value'],1), 'roof', round(d['roofline']['frac'],3), 'sust', round(d['sustained']['ms_per_frame'],4), 'nccl', d.get('nccl_gather_comparison',{}).get('ms_per_frame'), 'C2', d['extra']['C2']['ms_per_frame'])
PY
