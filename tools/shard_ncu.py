"""One rank's share of an N-way sharded frame, a few times (for ncu).  usage: shard_ncu.py [config] [rank] [N] [lanes]"""
import os, sys
sys.path.insert(0, os.getcwd()); sys.path.insert(0, os.path.join(os.getcwd(), "tests"))
import _pkg, scenes
mm = _pkg.load_package()
assets = scenes.load_assets()
cfg = sys.argv[1] if len(sys.argv) > 1 else "C2"
rank = int(sys.argv[2]) if len(sys.argv) > 2 else 0
n = int(sys.argv[3]) if len(sys.argv) > 3 else 8
sc = scenes.make_scene(mm, cfg, assets)
cs = mm.ComputeShader(0, (sc["W"], sc["H"]), placement=sc["textures"]["placement"], curl=sc["textures"]["curl"], lowRes=sc["textures"]["lowres"], hiRes=sc["textures"]["hires"])
cs.allocOutput()
cs.setLanesPerRay(int(sys.argv[4]) if len(sys.argv) > 4 else 0)
cs.updateUniformBuffers(sc["cam"], None, sc["sky"], sc["sun"])
for _ in range(4):
    cs.dispatch(mm.MM_FULL, rank, n, 8)
    cs.synchronize()
    print(f"{cfg} rank {rank} of {n}: ms", cs.lastKernelMs())
cs.close()
