import os, sys
sys.path.insert(0, os.getcwd()); sys.path.insert(0, os.path.join(os.getcwd(), "tests"))
import numpy as np, torch, _pkg, scenes
mm = _pkg.load_package(); assets = scenes.load_assets()
sc = scenes.make_scene(mm, "C3", assets); W, H = sc["W"], sc["H"]
cs = mm.ComputeShader(0, (W, H), placement=sc["textures"]["placement"], curl=sc["textures"]["curl"], lowRes=sc["textures"]["lowres"], hiRes=sc["textures"]["hires"])
cs.allocOutput(); cs.updateUniformBuffers(sc["cam"], None, sc["sky"], sc["sun"]); cs.setLanesPerRay(1)
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
def t(r, n, rb):
    best = 1e9
    for rep in range(3):
        flush.fill_(1); torch.cuda.synchronize(); cs.dispatch(mm.MM_FULL, r, n, rb); v = cs.lastKernelMs()
        if rep: best = min(best, v)
    return best
print(os.environ.get("MM_DEBUG_ROW_ORDER", "cost-ordered"), "N=8:", " ".join(f"{t(r, 8, 8):.3f}" for r in (0, 1, 2, 7)), "| N=9:", " ".join(f"{t(r, 9, 8):.3f}" for r in (0, 1, 8)), "| N=7:", " ".join(f"{t(r, 7, 8):.3f}" for r in (0, 1, 6)))
cs.close()
