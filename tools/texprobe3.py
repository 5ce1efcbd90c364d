"""Third probe: non-power-of-two extents with random (non-dyadic) coordinates, 2D and 3D (writes gpurun_out/texprobe3.npz)."""
import os, sys
sys.path.insert(0, os.getcwd()); sys.path.insert(0, os.path.join(os.getcwd(), "tests"))
import numpy as np, _pkg
mm = _pkg.load_package()
rng = np.random.default_rng(77)
out = {}
for name, shape in (("n50x27", (27, 50, 4)), ("n5x3", (3, 5, 4)), ("n1920x1080", (1080, 1920, 4)), ("n3x7", (7, 3, 4))):
    t = rng.integers(0, 256, shape, dtype=np.uint8)
    cs = mm.ComputeShader(0, (8, 8), nightSky=t)
    for tag, lo, hi in (("unit", 0.0, 1.0), ("wide", -3.0, 3.0), ("far", -300.0, 300.0)):
        uvw = rng.uniform(lo, hi, (150000, 3)).astype(np.float32)
        out[f"{name}_{tag}_uvw"] = uvw
        out[f"{name}_{tag}_out"] = cs.sample(mm.MM_TEX_NIGHTSKY, mm.MM_FILTER_HW, uvw)
    out[f"{name}_tex"] = t
    cs.close()
for name, shape in (("v5x6x7", (7, 6, 5, 4)), ("v3x3x3", (3, 3, 3, 4))):
    t = rng.integers(0, 256, shape, dtype=np.uint8)
    cs = mm.ComputeShader(0, (8, 8), lowRes=t)
    for tag, lo, hi in (("unit", 0.0, 1.0), ("wide", -3.0, 3.0)):
        uvw = rng.uniform(lo, hi, (150000, 3)).astype(np.float32)
        out[f"{name}_{tag}_uvw"] = uvw
        out[f"{name}_{tag}_out"] = cs.sample(mm.MM_TEX_LOWRES, mm.MM_FILTER_HW, uvw)
    out[f"{name}_tex"] = t
    cs.close()
os.makedirs("gpurun_out", exist_ok=True)
np.savez_compressed("gpurun_out/texprobe3.npz", **out)
print("texprobe3 done")
