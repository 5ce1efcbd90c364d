#!/usr/bin/env python
"""Generate tests/golden/frames/*.npz: small frames of the CPU oracle for the frozen configs (test infrastructure).

The reference itself cannot be executed for this path (no Vulkan/glslang/lavapipe here), so these are outputs of
the oracle restatement, committed to (a) detect drift of the oracle and (b) give the GPU box a fixed target that
does not depend on its own CPU's libm.  Stored: RGBA32F image, the RGBA8 tonemap of it, per-pixel counters."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import _pkg, scenes, oracle_binding as ob

CASES = [("C1", 96, 54, {}), ("C3", 96, 54, {}), ("C5b", 64, 36, {}), ("C1", 64, 36, {"time": 40.0, "wind": (0.6, 0.05, -1.2)}),
         ("C1", 64, 36, {"elevation": 0.75})]   # the last one is a NIGHT frame (sun below the horizon, CC:365-384)

def main(only=()):
    mm = _pkg.load_package()
    assets = scenes.load_assets()
    # frames/: oracle with the exact binary32 sampler (OM_FILTER_FP32); frames_texunit/: oracle with the bit-exact
    # model of the B200 texture unit (OM_FILTER_TEXUNIT) -- the target of the hardware-sampler march (MM_FILTER_HW)
    for sub, filt in (("frames", ob.OM_FILTER_FP32), ("frames_texunit", ob.OM_FILTER_TEXUNIT)):
        if only and sub not in only:
            continue
        out = os.path.join(ROOT, "tests", "golden", sub)
        os.makedirs(out, exist_ok=True)
        for i, (name, W, H, over) in enumerate(CASES):
            sc = scenes.make_scene(mm, name, assets, W=W, H=H, **over)
            night = scenes.synthetic_night_sky() if sc["sun"][5] < 0 else None
            img, cnt = ob.Scene(sc["textures"], sc["cam"], sc["sun"], sc["sky"], nightsky=night, filter_mode=filt).march(W, H)
            np.savez_compressed(os.path.join(out, f"frame{i}_{name}_{W}x{H}.npz"), rgba32f=img, rgba8=ob.tonemap_rgba8(img), counters=cnt.astype(np.uint16),
                                cam=sc["cam"], sun=sc["sun"], sky=sc["sky"], config=np.array(repr((name, W, H, over))))
            print(sub, i, name, W, H, over, "trips/px %.1f lit/px %.2f" % (cnt[..., 0].mean(), cnt[..., 3].mean()))

if __name__ == "__main__":
    main(tuple(sys.argv[1:]))      # optional: which sub-directories to (re)generate
