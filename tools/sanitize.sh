#!/bin/bash
# compute-sanitizer pass over the kernels (run on the GPU box): memcheck + racecheck on a small frame of every mode,
# the noise builds, the reprojection pass, the tonemap, the post chain and the cloud-shadow pass.  Output -> gpurun_out/sanitizer_*.log
mkdir -p gpurun_out
cat > /tmp/san_driver.py <<'PY'
import sys, os
sys.path.insert(0, os.getcwd()); sys.path.insert(0, os.path.join(os.getcwd(), "tests"))
import numpy as np, _pkg, scenes
mm = _pkg.load_package()
assets = scenes.load_assets()
W, H = 96, 54
sc = scenes.make_scene(mm, "C1", assets, W=W, H=H)
cs = mm.ComputeShader(0, (W, H), placement=sc["textures"]["placement"], curl=sc["textures"]["curl"], lowRes=sc["textures"]["lowres"], hiRes=sc["textures"]["hires"])
cs.allocOutput()
# every march kernel variant: K1 (static grid), K1p (persistent warps; a tile at a time, refill at 16 / 8 dead lanes) and
# K1s (2, 4, 8 lanes per ray), with and without counters, in the three sampler modes
variants = [(1, mm.MM_SCHED_STATIC, 0), (1, mm.MM_SCHED_PACKED, 0), (1, mm.MM_SCHED_PERSISTENT, 32), (1, mm.MM_SCHED_PERSISTENT, 16), (1, mm.MM_SCHED_PERSISTENT, 8),
            (2, mm.MM_SCHED_AUTO, 0), (4, mm.MM_SCHED_AUTO, 0), (8, mm.MM_SCHED_AUTO, 0)]
quick = os.environ.get("SAN_QUICK") == "1"                 # SAN_QUICK=1: memcheck only, K1 + K1s (2 lanes), both arithmetic builds, no counters
if quick:
    variants = [variants[0], variants[5]]
for arith in (mm.MM_ARITH_IEEE, mm.MM_ARITH_FMA):          # both builds of the march (cloud_march.cu, cloud_march_fma.cu)
    cs.setArithmetic(arith)
    for counters in ((False,) if quick else (False, True)):
        cs.enableCounters(counters)
        for mode in (mm.MM_FILTER_EXACT, mm.MM_FILTER_HW, mm.MM_FILTER_HYBRID):
            cs.setFilterMode(mode)
            for lanes, sched, refill in variants:
                cs.setLanesPerRay(lanes)
                cs.setScheduler(sched, refill)
                cs.renderToHost(sc["cam"], sc["sky"], sc["sun"])
                cs.updateUniformBuffers(sc["cam"], sc["cam"], sc["sky"], sc["sun"])
                cs.dispatch(mm.MM_PHASE16)
                cs.dispatch(mm.MM_FULL, 1, 3, 2)
                cs.synchronize()
cs.setArithmetic(mm.MM_ARITH_IEEE)
cs.setLanesPerRay(0)
cs.setScheduler(mm.MM_SCHED_AUTO, 0)
cs.tonemapRGBA8()
import torch
src = torch.rand((H, W, 4), device="cuda")
fb1, fb2 = torch.empty_like(src), torch.empty_like(src)
out8 = torch.empty((H, W, 4), dtype=torch.uint8, device="cuda")
cs.godRay(sc["cam"], sc["sun"], src.data_ptr(), fb1.data_ptr())
cs.radialBlur(sc["cam"], sc["sun"], fb1.data_ptr(), fb2.data_ptr())
cs.tonemapPresent(fb2.data_ptr(), out8.data_ptr())
cs.postChain(sc["cam"], sc["sun"], src.data_ptr(), out8.data_ptr(), bgra=True)
cs.synchronize()
pos = np.random.default_rng(0).uniform(-20000, 20000, (5000, 3)).astype(np.float32)
for mode in (mm.MM_FILTER_EXACT, mm.MM_FILTER_HW):
    cs.setFilterMode(mode)
    cs.cloudShadow(pos, want_fetches=True)
cs.buildCurlNoise()
if "--volumes" in sys.argv:
    cs.buildNoiseVolumes(1)
cs.close()
print("sanitizer driver done")
PY
for tool in memcheck $([ "$SAN_QUICK" = 1 ] || echo racecheck); do
  compute-sanitizer --tool $tool --error-exitcode 9 python /tmp/san_driver.py > gpurun_out/sanitizer_$tool.log 2>&1
  echo "$tool exit=$?"; tail -4 gpurun_out/sanitizer_$tool.log
done
