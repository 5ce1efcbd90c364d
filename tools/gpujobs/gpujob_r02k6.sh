# round 2 (1 GPU): K6 with the special-function-unit pow in the present pass -- tests, and per-launch times old / new (ncu, cold cache)
set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_post_chain.py -m gpu -x -q > gpurun_out/r02_k6_tests.log 2>&1; tail -2 gpurun_out/r02_k6_tests.log
for v in old new; do
  lib=project-marshmallow_b200/libmarshmallow_b200.so; [ $v = old ] && lib=variants/k6old.so
  MM_LIBRARY=$PWD/$lib timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r02_k6_$v.csv python tools/aux_kernels_driver.py > /dev/null 2>&1
  python - $v <<'PY'
import csv, sys, collections
rows = [r for r in csv.reader(open(f"gpurun_out/r02_k6_{sys.argv[1]}.csv")) if len(r) > 10]
hdr = rows[0]; k = hdr.index("Kernel Name"); v = hdr.index("Metric Value"); u = hdr.index("Metric Unit")
t = collections.defaultdict(list)
for r in rows[1:]:
    val = float(r[v].replace(",", "")); val = val / 1000 if r[u] == "ns" else val
    t[r[k][:70]].append(val)
for name, vals in t.items():
    if any(s in name for s in ("god_ray", "radial_blur", "present", "tonemap", "reproject")):
        print(sys.argv[1], name, "us: last", round(vals[-1], 2), "min", round(min(vals), 2))
PY
done
