import os, sys
sys.path.insert(0, os.getcwd()); sys.path.insert(0, os.path.join(os.getcwd(), "tests"))
import numpy as np, torch, _pkg, scenes
mm = _pkg.load_package(); assets = scenes.load_assets()
sc = scenes.make_scene(mm, "C3", assets); W, H = sc["W"], sc["H"]
cs = mm.ComputeShader(0, (W, H), placement=sc["textures"]["placement"], curl=sc["textures"]["curl"], lowRes=sc["textures"]["lowres"], hiRes=sc["textures"]["hires"])
cs.allocOutput(); cs.updateUniformBuffers(sc["cam"], None, sc["sky"], sc["sun"]); cs.setLanesPerRay(1)
cs.enableCounters(True)
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
def t(mode, r, n, rb):
    best = 1e9
    for rep in range(3):
        flush.fill_(1); torch.cuda.synchronize(); cs.dispatch(mode, r, n, rb); v = cs.lastKernelMs()
        if rep: best = min(best, v)
    return best
cs.dispatch(mm.MM_FULL); cs.synchronize(); cnt = cs.readCounters(); cs.enableCounters(False)
cost = cnt[..., 0].sum(axis=1).astype(np.float64)      # loop trips per row
for snake in (0, mm.MM_ROWS_SNAKE):
    ts = [t(mm.MM_FULL | snake, r, 8, 8) for r in range(8)]
    work = [cost[mm.multigpu.owned_rows(H, r, 8, 8, bool(snake))].sum() / cost.sum() * 8 for r in range(8)]
    print("snake" if snake else "plain", "ms:", " ".join(f"{x:.3f}" for x in ts), "| relative trips:", " ".join(f"{x:.3f}" for x in work))
cs.close()
