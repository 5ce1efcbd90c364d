"""Probe the B200 texture unit's linear filtering arithmetic (run on the GPU box; writes gpurun_out/texprobe.npz).
Goal: a CPU model of tex2D/tex3D (RGBA8_UNORM, normalized coords, wrap, linear) that is bit-exact, so that the
hardware-sampler march can be checked against an oracle that filters the same way."""
import os, sys
sys.path.insert(0, os.getcwd()); sys.path.insert(0, os.path.join(os.getcwd(), "tests"))
import numpy as np, _pkg
mm = _pkg.load_package()
rng = np.random.default_rng(123)
out = {}

def tex2d(arr):
    cs = mm.ComputeShader(0, (8, 8), placement=arr)
    return cs
def sample(cs, slot, uvw):
    return cs.sample(slot, mm.MM_FILTER_HW, uvw.astype(np.float32))

# A: 1D ramp 0,255 : weight quantisation and the coordinate -> fixed point rounding
t = np.zeros((1, 4, 4), np.uint8); t[0, :, 0] = [0, 255, 0, 255]; t[0, :, 1] = [10, 11, 12, 13]; t[0, :, 2] = [37, 201, 90, 3]; t[0, :, 3] = 255
cs = tex2d(t)
n = 1 << 18
u = (0.125 + (np.arange(n, dtype=np.float64) / n) * 0.25)            # from centre of texel 0 to centre of texel 1
uvw = np.stack([u, np.full(n, 0.5), np.zeros(n)], 1)
out["A_u"] = u.astype(np.float32); out["A_tex"] = t; out["A_out"] = sample(cs, mm.MM_TEX_PLACEMENT, uvw)
# A2: same sweep shifted by whole periods (wrap, large coordinates)
for k, shift in enumerate((3.0, -5.0, 61.0, 1000.0)):
    uvw2 = uvw.copy(); uvw2[:, 0] += shift
    out[f"A_shift{k}_u"] = uvw2[:, 0].astype(np.float32); out[f"A_shift{k}_out"] = sample(cs, mm.MM_TEX_PLACEMENT, uvw2)
cs.close()

# B: non power of two width (5) ramp
t5 = np.zeros((1, 5, 4), np.uint8); t5[0, :, 0] = [0, 255, 0, 255, 0]; t5[0, :, 1] = [1, 50, 100, 150, 200]; t5[0, :, 3] = 255
cs = tex2d(t5)
u = (np.arange(n, dtype=np.float64) / n) * 2.0 - 0.5
uvw = np.stack([u, np.full(n, 0.5), np.zeros(n)], 1)
out["B_u"] = u.astype(np.float32); out["B_tex"] = t5; out["B_out"] = sample(cs, mm.MM_TEX_PLACEMENT, uvw)
cs.close()

# C: bilinear 2x2 random texels, all 256x256 weight combinations (coordinates at the centres of the 1/256 weight cells) + random coords
t22 = rng.integers(0, 256, (2, 2, 4), dtype=np.uint8)
cs = tex2d(t22)
a = (np.arange(256) + 0.5) / 256.0
U = 0.25 + (a[None, :] / 2.0) * np.ones((256, 1)); V = 0.25 + (a[:, None] / 2.0) * np.ones((1, 256))     # texel centres at 0.25 and 0.75
uvw = np.stack([U.ravel(), V.ravel(), np.zeros(65536)], 1)
out["C_tex"] = t22; out["C_uv"] = uvw.astype(np.float32); out["C_out"] = sample(cs, mm.MM_TEX_PLACEMENT, uvw)
r = rng.uniform(-3, 3, (200000, 3)); out["C_ruv"] = r.astype(np.float32); out["C_rout"] = sample(cs, mm.MM_TEX_PLACEMENT, r)
cs.close()

# D: trilinear 2x2x2 random texels: 64^3 weight grid + random coords; and a 4x4x4 random volume with random coords
t222 = rng.integers(0, 256, (2, 2, 2, 4), dtype=np.uint8)
cs = mm.ComputeShader(0, (8, 8), lowRes=t222)
a = (np.arange(0, 256, 4) + 0.5) / 256.0
A, B, G = np.meshgrid(a, a, a, indexing="ij")
uvw = np.stack([0.25 + A.ravel() / 2, 0.25 + B.ravel() / 2, 0.25 + G.ravel() / 2], 1)
out["D_tex"] = t222; out["D_uvw"] = uvw.astype(np.float32); out["D_out"] = sample(cs, mm.MM_TEX_LOWRES, uvw)
r = rng.uniform(-3, 3, (300000, 3)); out["D_ruvw"] = r.astype(np.float32); out["D_rout"] = sample(cs, mm.MM_TEX_LOWRES, r)
cs.close()
t444 = rng.integers(0, 256, (4, 4, 4, 4), dtype=np.uint8)
cs = mm.ComputeShader(0, (8, 8), lowRes=t444)
r = rng.uniform(-2, 50, (300000, 3)); out["E_tex"] = t444; out["E_ruvw"] = r.astype(np.float32); out["E_rout"] = sample(cs, mm.MM_TEX_LOWRES, r)
cs.close()

# F: the real textures at coordinates like the march's
sys.path.insert(0, "tests")
import scenes
assets = scenes.load_assets()
cs = mm.ComputeShader(0, (8, 8), placement=assets["placement"], curl=assets["curl"], lowRes=assets["lowres"], hiRes=assets["hires"])
for name, slot, lo, hi in (("placement", 0, -1.5, 1.5), ("curl", 2, -15, 15), ("lowres", 3, -3, 3), ("hires", 4, -60, 60)):
    r = rng.uniform(lo, hi, (400000, 3)); out[f"F_{name}_uvw"] = r.astype(np.float32); out[f"F_{name}_out"] = sample(cs, slot, r)
cs.close()
os.makedirs("gpurun_out", exist_ok=True)
np.savez_compressed("gpurun_out/texprobe.npz", **out)
print("texprobe done", {k: v.shape for k, v in out.items() if k.endswith("out")})
