#!/usr/bin/env python
"""How much could a coarse-coverage prepass PROVE?  ("measure before building the coarse-coverage prepass".)

For every cloudTest call of a frame (march trips and light-cone samples; oracle trace, uncontracted arithmetic, binary32 sampler) this
computes the most favourable conservative bound a prepass could hold for the call: the sample's own height, cloud type and coverage are
taken EXACTLY (a real prepass would have to bound those over a brick as well, and the wind offset moves the noise lattice against
the placement map every frame), and only the low-res noise volume is bounded per brick of B^3 texels (+ the 1-texel apron of the
trilinear footprint): Rmax = max of channel R, Emin = min of the erosion FBM 0.625 G + 0.25 B + 0.125 A.  Trilinear filtering is a
convex combination, so for any sample in the brick  density <= layerDensity * remapC(Rmax, 0.3, 1)  and  erosion >= remapC(Emin,
coverage, 1).  The call is PROVABLY zero when the density bound is < 1e-4 (CC:243) or <= the erosion bound (CC:248-250).

    python tools/prepass_bound.py [--config C3] [--width 240]
"""
import argparse, ctypes as C, os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import oracle_binding as ob, scenes


def remapc(v, lo):
    with np.errstate(divide="ignore", invalid="ignore"):
        q = (v - lo) / (1.0 - lo)
    q = np.where(np.isnan(q), 0.0, q)
    return np.clip(q, 0.0, 1.0)


def brick_bounds(vol, B):
    """vol: [z][y][x][4] uint8 -> (Rmax, Emin) per brick of B^3 texels including a 1-texel apron on the + side (REPEAT wrap; the
    footprint of a sample whose lower texel lies in the brick reaches one texel further)"""
    R = vol[..., 0].astype(np.float32) / 255.0
    E = (0.625 * vol[..., 1].astype(np.float32) + 0.25 * vol[..., 2].astype(np.float32) + 0.125 * vol[..., 3].astype(np.float32)) / 255.0
    n = vol.shape[0] // B
    Rmax = np.full((n, n, n), -1.0, np.float32); Emin = np.full((n, n, n), 2.0, np.float32)
    for dz in range(B + 1):
        for dy in range(B + 1):
            for dx in range(B + 1):
                r = np.roll(R, (-dz, -dy, -dx), (0, 1, 2))[::B, ::B, ::B]
                e = np.roll(E, (-dz, -dy, -dx), (0, 1, 2))[::B, ::B, ::B]
                np.maximum(Rmax, r, out=Rmax); np.minimum(Emin, e, out=Emin)
    return Rmax, Emin


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--config", default="C3")
    ap.add_argument("--width", type=int, default=240)
    a = ap.parse_args()
    assets = scenes.load_assets()
    sc = scenes.scene_from_config(a.config, assets)
    W, H = a.width, a.width * 9 // 16
    S = ob.Scene(sc["textures"], sc["cam"], sc["sun"], sc["sky"])
    lib = ob.lib()
    cap = 40_000_000
    buf = np.zeros((cap, 8), np.float32)
    lib.om_debug_trace.argtypes = [C.c_void_p, C.c_size_t]
    lib.om_debug_trace_count.restype = C.c_size_t
    lib.om_debug_trace(buf.ctypes.data, cap)
    _, cnt = S.march(W, H, nthreads=1)
    n = lib.om_debug_trace_count()
    lib.om_debug_trace(None, 0)
    t = buf[:n]
    u, v, w, layer, cov, gate, res, h = (t[:, i] for i in range(8))
    print(f"{a.config} {W}x{H}: {n} cloudTest calls, {cnt[..., 0].sum()} loop trips, {cnt[..., 3].sum()} lit steps")
    print(f"  actual outcome: zero {np.mean(res == 0):.3f}  (layer density 0: {np.mean(layer == 0):.3f}, low-res gate < 1e-4: {np.mean((layer > 0) & (gate < 1e-4)):.3f}, "
          f"eroded to zero: {np.mean((gate >= 1e-4) & (res == 0)):.3f}), hit {np.mean(res > 0):.3f}")
    low = sc["textures"]["lowres"]
    N = low.shape[0]
    for B in (1, 2, 4, 8):
        Rmax, Emin = brick_bounds(low, B)
        ix = (np.floor(u * N - 0.5).astype(np.int64) % N) // B
        iy = (np.floor(v * N - 0.5).astype(np.int64) % N) // B
        iz = (np.floor(w * N - 0.5).astype(np.int64) % N) // B
        rmax, emin = Rmax[iz, iy, ix], Emin[iz, iy, ix]
        d_ub = layer * remapc(rmax, 0.3)
        e_lb = remapc(emin, cov)
        proven = (d_ub < 1e-4) | (d_ub <= e_lb)
        assert not (proven & (res > 0)).any(), "bound is not conservative"
        beyond_layer = proven & (layer > 0)
        print(f"  brick {B}^3 texels ({B * 391} world units): provably zero {proven.mean():.3f} of all calls "
              f"(of which the height/type gradient alone: {np.mean(layer == 0):.3f}; added by the noise bound: {beyond_layer.mean():.3f}); "
              f"share of the actual zeros proven: {proven.sum() / max(1, (res == 0).sum()):.3f}")


if __name__ == "__main__":
    main()
