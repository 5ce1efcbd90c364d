"""Is rank 0's slowness at N=8 / 4K an address effect of the output stores?  Same shares with a padded image pitch."""
import os, sys
sys.path.insert(0, os.getcwd()); sys.path.insert(0, os.path.join(os.getcwd(), "tests"))
import numpy as np, torch, _pkg, scenes
mm = _pkg.load_package(); assets = scenes.load_assets()
sc = scenes.make_scene(mm, "C3", assets); W, H = sc["W"], sc["H"]
cs = mm.ComputeShader(0, (W, H), placement=sc["textures"]["placement"], curl=sc["textures"]["curl"], lowRes=sc["textures"]["lowres"], hiRes=sc["textures"]["hires"])
cs.updateUniformBuffers(sc["cam"], None, sc["sky"], sc["sun"]); cs.setLanesPerRay(1)
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
def t(r, n, rb):
    best = 1e9
    for rep in range(3):
        flush.fill_(1); torch.cuda.synchronize(); cs.dispatch(mm.MM_FULL, r, n, rb); v = cs.lastKernelMs()
        if rep: best = min(best, v)
    return best
for pad in (0, 256, 4096, 16 * 67):
    pitch = W * 16 + pad
    buf = torch.empty(pitch * H + 4096, dtype=torch.uint8, device="cuda")
    cs.bindOutput(buf.data_ptr(), pitch)
    print(f"pitch {pitch}: N=8 ranks 0,1,4,7:", " ".join(f"{t(r, 8, 8):.3f}" for r in (0, 1, 4, 7)))
cs.close()
