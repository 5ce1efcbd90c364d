"""Rank 0's share of an 8-way sharded C2 frame, a few times (for ncu: the K1s ray-split kernel chosen for this launch size)."""
import os, sys
sys.path.insert(0, os.getcwd()); sys.path.insert(0, os.path.join(os.getcwd(), "tests"))
import _pkg, scenes
mm = _pkg.load_package()
assets = scenes.load_assets()
sc = scenes.make_scene(mm, "C2", assets)
cs = mm.ComputeShader(0, (sc["W"], sc["H"]), placement=sc["textures"]["placement"], curl=sc["textures"]["curl"], lowRes=sc["textures"]["lowres"], hiRes=sc["textures"]["hires"])
cs.allocOutput()
cs.updateUniformBuffers(sc["cam"], None, sc["sky"], sc["sun"])
for _ in range(4):
    cs.dispatch(mm.MM_FULL, 0, 8, 8)
    cs.synchronize()
    print("rank 0 of 8: ms", cs.lastKernelMs())
cs.close()
