#!/usr/bin/env python
"""Where does the end-to-end frame spend its time?  Kernel time (CUDA events inside mm_dispatch) with and without the host mirror,
wall clock of mm_render_to_host, and a plain device->host copy of the frame for comparison.  python tools/e2e_probe.py [--config C3]"""
import argparse, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, torch, _pkg, scenes
ap = argparse.ArgumentParser(); ap.add_argument("--config", default="C3"); a = ap.parse_args()
mm = _pkg.load_package()
sc = scenes.make_scene(mm, a.config, scenes.load_assets()); W, H = sc["W"], sc["H"]
cs = mm.ComputeShader(0, (W, H), placement=sc["textures"]["placement"], curl=sc["textures"]["curl"], lowRes=sc["textures"]["lowres"], hiRes=sc["textures"]["hires"])
cs.allocOutput()
host = torch.empty((H, W, 4), dtype=torch.float32, pin_memory=True); hnp = host.numpy()
cs.updateUniformBuffers(sc["cam"], None, sc["sky"], sc["sun"])
def avg(f, n=6):
    f(); f(); t = []
    for _ in range(n): t.append(f())
    return sum(t) / n
def plain():
    cs.dispatch(mm.MM_FULL); cs.synchronize(); return cs.lastKernelMs()
def mirrored_kernel():
    cs.renderToHost(sc["cam"], sc["sky"], sc["sun"], out=hnp); return cs.lastKernelMs()
def mirrored_wall():
    t0 = time.perf_counter(); cs.renderToHost(sc["cam"], sc["sky"], sc["sun"], out=hnp); return (time.perf_counter() - t0) * 1e3
dev = torch.empty((H, W, 4), dtype=torch.float32, device="cuda")
def d2h():
    torch.cuda.synchronize(); t0 = time.perf_counter(); host.copy_(dev, non_blocking=True); torch.cuda.synchronize(); return (time.perf_counter() - t0) * 1e3
print(f"{a.config}: kernel alone {avg(plain):.3f} ms | kernel with host mirror {avg(mirrored_kernel):.3f} ms | mm_render_to_host wall {avg(mirrored_wall):.3f} ms | plain D2H copy of the frame {avg(d2h):.3f} ms ({W*H*16/1e6:.0f} MB)")
cs.close()
