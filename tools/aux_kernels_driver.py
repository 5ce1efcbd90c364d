"""Launch every auxiliary kernel once at its BASELINE size (run under `ncu --metrics gpu__time_duration.sum`):
K2 curl noise 128^2, K3 volumes 128^3 + 32^3, K4 tonemap 1080p, K5 reprojection 1080p, texture packing."""
import os, sys
sys.path.insert(0, os.getcwd()); sys.path.insert(0, os.path.join(os.getcwd(), "tests"))
import numpy as np, torch, _pkg, scenes
mm = _pkg.load_package()
assets = scenes.load_assets()
W, H = 1920, 1080
sc = scenes.make_scene(mm, "C2", assets)
cs = mm.ComputeShader(0, (W, H), placement=sc["textures"]["placement"], curl=sc["textures"]["curl"], lowRes=sc["textures"]["lowres"], hiRes=sc["textures"]["hires"])
a = torch.rand((H, W, 4), device="cuda") * 30
b = torch.zeros((H, W, 4), device="cuda")
cs.bindOutput(b.data_ptr())
cs.bindPrevious(a.data_ptr())
cam = sc["cam"]
prev = mm.host_camera((3.0, 1.0, 2.0), -np.pi / 2 + 0.01, -20 * scenes.DEG2RAD)
cs.updateUniformBuffers(cam, prev, sc["sky"], sc["sun"])
for _ in range(3):
    cs.dispatchReproject()
    cs.synchronize()
    cs.tonemapRGBA8()
    cs.buildCurlNoise(want_copy=False)
    cs.buildNoiseVolumes(0, want_copy=False)
cs.close()
print("aux driver done")
