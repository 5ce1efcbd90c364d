"""Per-GPU kernel time of an N-way row-cyclic sharded frame, measured on ONE GPU (each rank's dispatch is timed alone with
CUDA events, L2 flushed before it; the N-GPU frame time is the slowest rank's).  usage: shard_probe.py [config] [row_block]"""
import os, sys
sys.path.insert(0, os.getcwd()); sys.path.insert(0, os.path.join(os.getcwd(), "tests"))
import numpy as np, torch, _pkg, scenes
mm = _pkg.load_package()
assets = scenes.load_assets()
cfg = sys.argv[1] if len(sys.argv) > 1 else "C2"
rb = int(sys.argv[2]) if len(sys.argv) > 2 else 4
sc = scenes.make_scene(mm, cfg, assets)
W, H = sc["W"], sc["H"]
cs = mm.ComputeShader(0, (W, H), placement=sc["textures"]["placement"], curl=sc["textures"]["curl"], lowRes=sc["textures"]["lowres"], hiRes=sc["textures"]["hires"])
cs.allocOutput()
cs.updateUniformBuffers(sc["cam"], None, sc["sky"], sc["sun"])
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
for n in (1, 2, 4, 8):
    row = []
    for lanes in (0, 1, 2, 4, 8):
        cs.setLanesPerRay(lanes)
        worst = 0.0
        for r in range(n):
            best = 1e9
            for rep in range(4):
                flush.fill_(1); torch.cuda.synchronize()
                cs.dispatch(mm.MM_FULL, r, n, rb)
                t = cs.lastKernelMs()
                if rep: best = min(best, t)
            worst = max(worst, best)
        row.append(worst)
    base = row[1] if n == 1 else base1
    if n == 1: base1 = row[1]
    print(f"{cfg} row_block {rb} N={n}: slowest-rank ms  auto {row[0]:.3f} | lanes1 {row[1]:.3f} lanes2 {row[2]:.3f} lanes4 {row[3]:.3f} lanes8 {row[4]:.3f} | speed-up of best vs 1 GPU {base1/min(row[1:]):.2f}x")
cs.close()
