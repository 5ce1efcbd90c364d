"""Launch every auxiliary kernel once at its BASELINE size (run under `ncu --metrics gpu__time_duration.sum`):
K2 curl noise 128^2, K3 volumes 128^3 + 32^3, K4 tonemap 1080p, K5 reprojection 1080p, K6 post chain 1080p (three passes and the
fused two-kernel chain), K7 cloud shadows for a 1080p G-buffer of ground positions, texture packing."""
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
fb1, fb2 = torch.empty_like(a), torch.empty_like(a)
out8 = torch.empty((H, W, 4), dtype=torch.uint8, device="cuda")
gpos = (torch.rand((H * W, 3), device="cuda") - 0.5) * torch.tensor([40000.0, 800.0, 40000.0], device="cuda")
gout = torch.empty(H * W, device="cuda")
cs.setFilterMode(mm.MM_FILTER_HW)
for _ in range(3):
    cs.godRay(cam, sc["sun"], a.data_ptr(), fb1.data_ptr())
    cs.radialBlur(cam, sc["sun"], fb1.data_ptr(), fb2.data_ptr())
    cs.tonemapPresent(fb2.data_ptr(), out8.data_ptr())
    cs.postChain(cam, sc["sun"], a.data_ptr(), out8.data_ptr())
    cs.cloudShadowDevice(gpos.data_ptr(), H * W, gout.data_ptr())
    cs.synchronize()
    cs.dispatchReproject()
    cs.synchronize()
    cs.tonemapRGBA8()
    cs.buildCurlNoise(want_copy=False)
    cs.buildNoiseVolumes(0, want_copy=False)
cs.close()
print("aux driver done")
