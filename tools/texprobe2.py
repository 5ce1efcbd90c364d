"""Second probe: recover the eight 8-bit corner weights of the 3D linear filter directly (one-hot corner textures)."""
import os, sys
sys.path.insert(0, os.getcwd()); sys.path.insert(0, os.path.join(os.getcwd(), "tests"))
import numpy as np, _pkg
mm = _pkg.load_package()
rng = np.random.default_rng(5)
out = {}
a = (np.arange(0, 256, 3) + 0.5) / 256.0          # 86 weight values per axis (k = 1, 4, 7, ...)
A, B, G = np.meshgrid(a, a, a, indexing="ij")
grid = np.stack([0.25 + A.ravel() / 2, 0.25 + B.ravel() / 2, 0.25 + G.ravel() / 2], 1)
r = np.stack([0.25 + rng.random(400000) / 2, 0.25 + rng.random(400000) / 2, 0.25 + rng.random(400000) / 2], 1)
for half in (0, 1):
    t = np.zeros((2, 2, 2, 4), np.uint8)
    for c in range(4):
        x, y = c & 1, (c >> 1) & 1
        t[half, y, x, c] = 255
    cs = mm.ComputeShader(0, (8, 8), lowRes=t)
    out[f"grid_z{half}"] = cs.sample(mm.MM_TEX_LOWRES, mm.MM_FILTER_HW, grid.astype(np.float32))
    out[f"rand_z{half}"] = cs.sample(mm.MM_TEX_LOWRES, mm.MM_FILTER_HW, r.astype(np.float32))
    cs.close()
out["grid_uvw"] = grid.astype(np.float32); out["rand_uvw"] = r.astype(np.float32)
np.savez_compressed("gpurun_out/texprobe2.npz", **out)
print("texprobe2 done")
