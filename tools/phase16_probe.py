"""One MM_FULL frame then MM_PHASE16 dispatches at 1080p (for ncu: the small-launch regime).  argv[1] = lanes per ray"""
import os, sys
sys.path.insert(0, os.getcwd()); sys.path.insert(0, os.path.join(os.getcwd(), "tests"))
import numpy as np, torch, _pkg, scenes
mm = _pkg.load_package()
assets = scenes.load_assets()
sc = scenes.make_scene(mm, sys.argv[2] if len(sys.argv) > 2 else "C2", assets)
W, H = sc["W"], sc["H"]
cs = mm.ComputeShader(0, (W, H), placement=sc["textures"]["placement"], curl=sc["textures"]["curl"], lowRes=sc["textures"]["lowres"], hiRes=sc["textures"]["hires"])
cs.allocOutput()
cs.setLanesPerRay(int(sys.argv[1]))
cs.updateUniformBuffers(sc["cam"], None, sc["sky"], sc["sun"])
for _ in range(4):
    cs.dispatch(mm.MM_PHASE16)
    cs.synchronize()
    print("phase16 ms", cs.lastKernelMs())
cs.close()
