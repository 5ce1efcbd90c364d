"""Per-rank shard times on one GPU: N x row_block x lanes grid.  usage: shard_probe3.py [config] [--snake]"""
import os, sys
sys.path.insert(0, os.getcwd()); sys.path.insert(0, os.path.join(os.getcwd(), "tests"))
import numpy as np, torch, _pkg, scenes
mm = _pkg.load_package()
assets = scenes.load_assets()
cfg = sys.argv[1] if len(sys.argv) > 1 and not sys.argv[1].startswith("-") else "C2"
sc = scenes.make_scene(mm, cfg, assets)
W, H = sc["W"], sc["H"]
cs = mm.ComputeShader(0, (W, H), placement=sc["textures"]["placement"], curl=sc["textures"]["curl"], lowRes=sc["textures"]["lowres"], hiRes=sc["textures"]["hires"])
cs.allocOutput()
cs.updateUniformBuffers(sc["cam"], None, sc["sky"], sc["sun"])
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
SNAKE = mm.MM_ROWS_SNAKE if "--snake" in sys.argv else 0
def t(r, n, rb):
    best = 1e9
    for rep in range(3):
        flush.fill_(1); torch.cuda.synchronize()
        cs.dispatch(mm.MM_FULL | SNAKE, r, n, rb)
        v = cs.lastKernelMs()
        if rep: best = min(best, v)
    return best
cs.setLanesPerRay(1)
full = t(0, 1, 1)
print(cfg, "full frame %.3f ms" % full)
for n in (2, 4, 8):
    for rb in (4, 8, 16, 32):
        out = []
        for lanes in (1, 2, 4):
            cs.setLanesPerRay(lanes)
            ts = [t(r, n, rb) for r in range(n)]
            out.append(f"lanes{lanes}: max {max(ts):.3f} min {min(ts):.3f} sum {sum(ts):.3f} ({full/max(ts):.2f}x)")
        print(f"N={n} rb={rb:2d}  " + " | ".join(out))
cs.close()
