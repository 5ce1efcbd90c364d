"""Why is a half frame more than half the time?  Per-rank times for N=2 at several row blocks, and the full frame."""
import os, sys
sys.path.insert(0, os.getcwd()); sys.path.insert(0, os.path.join(os.getcwd(), "tests"))
import numpy as np, torch, _pkg, scenes
mm = _pkg.load_package()
assets = scenes.load_assets()
sc = scenes.make_scene(mm, "C2", assets)
W, H = sc["W"], sc["H"]
cs = mm.ComputeShader(0, (W, H), placement=sc["textures"]["placement"], curl=sc["textures"]["curl"], lowRes=sc["textures"]["lowres"], hiRes=sc["textures"]["hires"])
cs.allocOutput()
cs.updateUniformBuffers(sc["cam"], None, sc["sky"], sc["sun"])
cs.setLanesPerRay(1)
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
def t(r, n, rb, fl=True):
    best = 1e9
    for rep in range(4):
        if fl: flush.fill_(1)
        torch.cuda.synchronize()
        cs.dispatch(mm.MM_FULL, r, n, rb)
        v = cs.lastKernelMs()
        if rep: best = min(best, v)
    return best
print("full frame: flushed %.3f  warm L2 %.3f" % (t(0, 1, 1), t(0, 1, 1, False)))
for rb in (1, 2, 4, 8, 16, 64, 540):
    a, b = t(0, 2, rb), t(1, 2, rb)
    print(f"N=2 row_block {rb}: rank0 {a:.3f} rank1 {b:.3f} sum {a+b:.3f}")
for n in (4, 8, 16):
    ts = [t(r, n, 4) for r in range(n)]
    print(f"N={n} row_block 4: " + " ".join(f"{x:.3f}" for x in ts) + f"  sum {sum(ts):.3f}")
cs.close()
